"""Autoregressive decoding with a KV cache (new capability; the reference re-runs the whole prefix
every step, generate.py:99-122).  The result of `prefill` / `step` equals the reference's
`model(prefix)[:, -1, :]` while the window does not slide (len <= max_input_len <= max_seq = 2048;
SURVEY.md 0.3): the relative logits depend only on i - j and the sinusoid on the absolute position.

    dec = KVCacheDecoder(model, batch_size=256, max_len=2048)            # allocates the caches once
    logits = dec.prefill(primer_tokens[B, t0], condition[B, 2])          # [B, V] for the next token
    logits = dec.step(next_tokens[B])                                    # one position per call

The per-step kernel sequence (input stage, NL x decode layer, output head, position increment) is
captured in a CUDA graph after the first step; the position lives in device memory.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib
from ._lib import ME_BF16, ME_F32, ptr
from .autograd import _Acts, _layer_args, _stream, _tdtype, run_forward


class KVCacheDecoder:
    def __init__(self, model, batch_size: int, max_len: Optional[int] = None, precision: str = "bf16",
                 use_cuda_graph: bool = True):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.model = model
        self.B = int(batch_size)
        self.T_max = int(max_len or model.max_seq)
        if self.T_max > model.max_seq:
            raise ValueError(f"max_len {self.T_max} exceeds max_seq {model.max_seq}: the cache is only valid "
                             "while the window does not slide")
        self.dtype = ME_BF16 if precision == "bf16" else ME_F32
        self.use_graph = bool(use_cuda_graph)
        dev = next(model.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("midi_emotion_b200: KVCacheDecoder needs the model on a CUDA device")
        self.dev = dev
        tdt = _tdtype(self.dtype)
        B, T, H = self.B, self.T_max, model.num_head
        d, di, V = model.embedding_dim, model.d_inner, model.vocab_size
        dh = d // H
        f32 = dict(device=dev, dtype=torch.float32)
        tt = dict(device=dev, dtype=tdt)
        self.k_cache = [torch.zeros(B, H, T, dh, **tt) for _ in range(model.num_layer)]
        self.v_cache = [torch.zeros(B, H, T, dh, **tt) for _ in range(model.num_layer)]
        self.keypad = torch.zeros(B, T, device=dev, dtype=torch.uint8)
        self.t_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.t_host = 0
        self.tokens_buf = torch.zeros(B, device=dev, dtype=torch.int64)
        # static condition buffer: the captured graph holds its address, prefill() copies into it
        self.cond_buf = torch.zeros(B, 2, **f32)
        self.cond = None
        self.Vp = (V + 7) // 8 * 8
        self.logits = torch.zeros(B, self.Vp, **tt)

        def pair():
            a = torch.empty(B, d, **f32)
            return a, (a if self.dtype == ME_F32 else torch.empty(B, d, **tt))

        self.x = [pair(), pair()]  # ping-pong residual stream
        self.scratch = {
            "qkv": torch.empty(B, 3 * d, **tt), "attn_o": torch.empty(B, d, **tt), "proj": torch.empty(B, d, **tt),
            "h": torch.empty(B, di, **tt),
        }
        self.scratch["out1_f32"], self.scratch["out1_T"] = pair()
        self.graph = None
        self._graph_wc = None
        self._eager_steps = 0

    # ------------------------------------------------------------------
    def reset(self):
        self.keypad.zero_()
        self.t_dev.zero_()
        self.t_host = 0

    def prefill(self, tokens: torch.Tensor, condition: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Run the primer through the full-sequence kernels, fill the caches, return the logits of
        the last position [B, V]."""
        model = self.model
        if tokens.shape[0] != self.B or tokens.dim() != 2:
            raise RuntimeError(f"prefill expects tokens [{self.B}, t0]")
        tokens = tokens.to(self.dev).contiguous()
        if model.mode != 0:
            if condition is None:
                raise RuntimeError("this model needs a (valence, arousal) condition")
            if tuple(condition.shape) != (self.B, 2):
                raise RuntimeError(f"condition must be [{self.B}, 2]")
            self.cond_buf.copy_(condition.to(device=self.dev, dtype=torch.float32))
            self.cond = self.cond_buf
        else:
            self.cond = None
        self.reset()
        Ls = tokens.shape[1] + (2 if model.continuous_token else 0)
        if Ls > self.T_max:
            raise RuntimeError(f"primer length {Ls} exceeds the cache length {self.T_max}")
        H, dh = model.num_head, model.embedding_dim // model.num_head
        stream = _stream()

        def sink(l, qkv):
            _lib.call("me_kv_cache_write", ptr(qkv), self.dtype, self.B, Ls, H, dh, ptr(self.k_cache[l]),
                      ptr(self.v_cache[l]), self.T_max, 0, stream)

        was_training = model.training
        model.eval()
        try:
            with torch.no_grad():
                logits, acts = run_forward(model, tokens, self.cond, False, dtype=self.dtype, kv_sink=sink,
                                           last_only=True)
        finally:
            model.train(was_training)
        self.keypad[:, :Ls].copy_(acts.keypad)
        self.t_host = Ls
        self.t_dev.fill_(Ls)
        return logits[:, :model.vocab_size]

    # ------------------------------------------------------------------
    def _enqueue_step(self, wc=None):
        model, dtype = self.model, self.dtype
        B, d, V = self.B, model.embedding_dim, model.vocab_size
        stream = _stream()
        wc = wc if wc is not None else model._weights(dtype)
        cw0, cb0, _, _ = model._cond_params()
        x_f32, x_T = self.x[0]
        _lib.call("me_embed_decode", ptr(self.tokens_buf), ptr(self.cond), ptr(model.embedding.weight), ptr(cw0),
                  ptr(cb0), ptr(model._pe(self.dev)), B, d, model.d_condition, V, model.mode, model.pad_token,
                  ptr(self.t_dev), dtype, ptr(x_f32), ptr(x_T), ptr(self.keypad), self.T_max, stream)
        a = _Acts()
        a.dtype, a.B, a.Ls, a.training, a.p, a.seed, a.attn_impl = dtype, B, 1, False, 0.0, 0, _lib.ATTN_SIMT
        for l, lay in enumerate(model.enc_layers):
            o_f32, o_T = self.x[(l + 1) & 1]
            act = dict(self.scratch)
            act["out2_f32"], act["out2_T"] = o_f32, o_T
            da = _lib.DecodeLayerArgs()
            da.f = _layer_args(model, wc["layers"][l], lay, act, x_f32, x_T, self.keypad, a, l)
            da.k_cache, da.v_cache = ptr(self.k_cache[l]), ptr(self.v_cache[l])
            da.t_dev, da.T_max = ptr(self.t_dev), self.T_max
            _lib.call("me_decode_layer_step", C.byref(da))
            x_f32, x_T = o_f32, o_T
        if dtype == ME_BF16:
            _lib.call("me_gemm_bf16", ptr(x_T), ptr(wc["Wfc"]), ptr(self.logits), B, V, d, d, d, self.Vp, 0, 0, ME_BF16,
                      _lib.EPI_BIAS, ptr(model.fc.bias), None, None, 0, stream)
        else:
            _lib.call("me_gemm_f32", ptr(x_T), ptr(wc["Wfc"]), ptr(self.logits), B, V, d, d, d, self.Vp, 0, 0,
                      _lib.EPI_BIAS, ptr(model.fc.bias), None, None, 0, stream)
        self.t_dev.add_(1)

    def step(self, tokens: torch.Tensor) -> torch.Tensor:
        """Append one token per sequence (int64 [B]) and return the next-token logits [B, V].
        The returned tensor is a view of a static buffer that the next call overwrites."""
        if self.t_host >= self.T_max:
            raise RuntimeError("KV cache is full: the reference would start sliding its window here, which "
                               "invalidates cached positions (SURVEY.md 0.3)")
        if self.model.mode != 0 and self.cond is None:
            raise RuntimeError("call prefill() first (it stores the condition)")
        self.tokens_buf.copy_(tokens.reshape(-1), non_blocking=True)
        with torch.no_grad():
            # the packed compute-type weight copies (re-derived here if a parameter moved); a captured graph holds
            # their addresses, so it is dropped when they were re-allocated (.to(), load_state_dict into new storage)
            wc = self.model._weights(self.dtype)
            if self.graph is not None and wc is not self._graph_wc:
                self.invalidate_graph()
            if not self.use_graph:
                self._enqueue_step(wc)
            elif self.graph is None:
                if self._eager_steps < 1:
                    self._enqueue_step(wc)        # first step eagerly: one-time kernel attribute setup
                    self._eager_steps += 1
                else:
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._enqueue_step(wc)
                    self.graph, self._graph_wc = g, wc
                    g.replay()
            else:
                self.graph.replay()
        self.t_host += 1
        return self.logits[:, :self.model.vocab_size]

    def invalidate_graph(self):
        """Drop the captured step (done automatically when the packed weight copies were re-allocated)."""
        self.graph = None
        self._graph_wc = None
        self._eager_steps = 0
