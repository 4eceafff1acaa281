"""Emotion-conditioned MIDI-token transformer: the Python surface of the reference's model classes
(models/music_multi.py:41-108 MusicTransformerMulti, models/music_continuous_token.py:32-105
MusicTransformerContinuousToken) with the arithmetic done by the sm_100a kernels behind the C-ABI.

* Same constructor keywords, attributes, `forward(x, condition)` signature and `state_dict` keys /
  shapes as the reference, so checkpoints load unchanged.
* No arithmetic happens in PyTorch: `forward` enqueues the C-ABI calls on the current CUDA stream;
  `backward` is hand-derived (one `torch.autograd.Function` around the whole model) so that
  `loss.backward()`, `clip_grad_norm_`, `optim.Adam` and DDP-style hooks see ordinary parameters.
* Precision follows the caller the way the reference does: under autocast the bf16 tensor-core path runs,
  otherwise the exact fp32 path.  The reference's own call sites (train.py:281, generate.py:116) use
  `torch.cuda.amp.autocast`, i.e. float16 + GradScaler: that is accepted and computed in bf16 (logits are
  returned as bfloat16; the scaler's loss scale passes through the linear backward unchanged).  `model.precision = "bf16" | "fp32"` overrides.
* CUDA only.  A CPU tensor, a missing library or an unsupported shape raises; nothing falls back.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import ME_BF16, ME_F32, ptr

MAX_SEQ = 2048
LN_EPS = 1e-6

_PE_CACHE: Dict[int, torch.Tensor] = {}


def positional_table(d: int, max_seq: int = MAX_SEQ) -> torch.Tensor:
    """[max_seq, d] fp32 sinusoid table, evaluated in float64 like models/music_multi.py:137-147
    (angle = pos * e^{-ln(1e4) i/d} * e^{ln(1e4)/d (i%2)} + (pi/2)(i%2)) and cast to fp32 (:157-158).
    It is not a parameter or buffer (absent from state_dict), exactly as in the reference."""
    key = (d, max_seq)
    if key not in _PE_CACHE:
        ln = math.log(10000)
        f = [math.exp(-ln * i / d) for i in range(d)]
        g = [math.exp(ln / d * (i % 2)) for i in range(d)]
        ph = [0.5 * math.pi * (i % 2) for i in range(d)]
        sin = math.sin
        rows = [[sin(pos * f[i] * g[i] + ph[i]) for i in range(d)] for pos in range(max_seq)]
        _PE_CACHE[key] = torch.tensor(rows, dtype=torch.float64).to(torch.float32)
    return _PE_CACHE[key]


class _RelativeGlobalAttention(nn.Module):
    """Parameter holder for models/music_multi.py:167-188 (Wq, Wk, Wv, fc, E)."""

    def __init__(self, h: int, d: int, max_seq: int):
        super().__init__()
        self.h, self.d, self.dh, self.max_seq = h, d, d // h, max_seq
        self.Wq = nn.Linear(d, d)
        self.Wk = nn.Linear(d, d)
        self.Wv = nn.Linear(d, d)
        self.fc = nn.Linear(d, d)
        self.E = nn.Parameter(torch.randn(max_seq, d // h))


class _EncoderLayer(nn.Module):
    """Parameter holder for models/music_multi.py:110-124."""

    def __init__(self, d_model: int, d_inner: int, h: int, max_seq: int):
        super().__init__()
        self.rga = _RelativeGlobalAttention(h, d_model, max_seq)
        self.FFN_pre = nn.Linear(d_model, d_inner)
        self.FFN_suf = nn.Linear(d_inner, d_model)
        self.layernorm1 = nn.LayerNorm(d_model, eps=LN_EPS)
        self.layernorm2 = nn.LayerNorm(d_model, eps=LN_EPS)


class MusicTransformer(nn.Module):
    """One class for the four conditioning modes.

    `continuous_token=True` reproduces MusicTransformerContinuousToken (two Linear(1, d) prefix
    vectors, output length L+2); otherwise MusicTransformerMulti (d_condition > 0 = concat)."""

    def __init__(self, embedding_dim=None, d_inner=None, d_condition=-1, vocab_size=None, num_layer=None,
                 num_head=None, max_seq=MAX_SEQ, dropout=0.0, pad_token=0, continuous_token=False):
        super().__init__()
        assert embedding_dim % num_head == 0, "d_model must be divisible by n_head"
        self.max_seq = max_seq
        self.num_layer = num_layer
        self.num_head = num_head
        self.embedding_dim = embedding_dim
        self.d_inner = d_inner
        self.vocab_size = vocab_size
        self.pad_token = pad_token
        self.continuous_token = bool(continuous_token)
        d_condition = 0 if (d_condition is None or d_condition < 0 or continuous_token) else d_condition
        self.d_condition = d_condition
        self.dropout_p = float(dropout)
        self.precision = "auto"        # "auto" (follow torch autocast) | "fp32" | "bf16"
        self.attn_impl = "auto"        # "auto" | "simt" | "tensor"
        self.causal = True             # False: models/music_regression.py (no mask at all)
        self.use_keypad = True         # key-pad part of generate_mask (music_multi.py:25-38)
        self._vocab_head = True        # fc = Linear(d, V) applied to every position
        # True: the tensor-core attention rounds QK^T, Srel, their sum and the scaled logits to bf16 exactly where
        # the reference does under autocast (music_multi.py:215-222); default keeps them in fp32
        self.reference_rounding = False
        # training, tensor-core attention: keep the probability tiles of the forward pass for the backward kernels
        # (0.44 GB per layer at 12 heads x 32 sequences x 1024 tokens) instead of recomputing them
        self.save_attention_probs = True

        self.embedding = nn.Embedding(vocab_size, embedding_dim - d_condition, padding_idx=pad_token)
        if self.continuous_token:
            self.fc_condition = nn.ModuleList([nn.Linear(1, embedding_dim) for _ in range(2)])
        elif d_condition > 0:
            self.fc_condition = nn.Linear(2, d_condition)
        self.enc_layers = nn.ModuleList(
            [_EncoderLayer(embedding_dim, d_inner, num_head, max_seq) for _ in range(num_layer)])
        self.fc = nn.Linear(embedding_dim, vocab_size)
        self._init_weights()
        self._wcache: Dict[int, dict] = {}
        self._pe_dev: Optional[torch.Tensor] = None
        self._step = 0

    # ------------------------------------------------------------------ parameters
    def _init_weights(self):
        """models/music_multi.py:74-82, music_continuous_token.py:67-75: U(-0.1, 0.1) for the
        embedding (including the pad row), head and condition weights; zero head/condition biases."""
        r = 0.1
        with torch.no_grad():
            self.embedding.weight.uniform_(-r, r)
            self.fc.bias.zero_()
            self.fc.weight.uniform_(-r, r)
            if self.continuous_token:
                for lin in self.fc_condition:
                    lin.weight.uniform_(-r, r)
                    lin.bias.zero_()
            elif self.d_condition > 0:
                self.fc_condition.bias.zero_()
                self.fc_condition.weight.uniform_(-r, r)

    @property
    def mode(self) -> int:
        if self.continuous_token:
            return _lib.COND_MODES["continuous_token"]
        if self.d_condition > 0:
            return _lib.COND_MODES["continuous_concat"]
        return _lib.COND_MODES["none"]  # none / discrete_token are the same arithmetic

    def _param_list(self) -> List[nn.Parameter]:
        # (walking named_parameters() costs ~0.6 ms at 209 tensors and this is called several times per step; the
        # Parameter objects of this model are never replaced -- .to() / load_state_dict keep them)
        pl = self.__dict__.get("_param_list_cache")
        if pl is None:
            pl = [p for _, p in self.named_parameters()]
            self.__dict__["_param_list_cache"] = pl
        return pl

    def _resolve_dtype(self) -> int:
        if self.precision == "bf16":
            return ME_BF16
        if self.precision == "fp32":
            return ME_F32
        if torch.is_autocast_enabled("cuda"):
            adt = torch.get_autocast_dtype("cuda")
            if adt == torch.bfloat16:
                return ME_BF16
            if adt == torch.float16:
                # train.py:281 / generate.py:116 call `torch.cuda.amp.autocast(enabled=amp)`, i.e. fp16 autocast with
                # a GradScaler (train.py:108).  The B200 path computes in bf16 (same tensor-core rate, fp32 exponent
                # range): it runs here unchanged (logits come back as bf16); the GradScaler's loss scale passes
                # through the backward linearly, the fp32 parameter gradients are unscaled by scaler.unscale_().
                global _WARNED_FP16
                if not _WARNED_FP16:
                    _WARNED_FP16 = True
                    import warnings
                    warnings.warn("midi_emotion_b200: float16 autocast requested; the B200 kernels compute in bfloat16 "
                                  "(logits are returned as bfloat16)", stacklevel=3)
                return ME_BF16
            raise RuntimeError(f"midi_emotion_b200: unsupported autocast dtype {adt}")
        return ME_F32

    def _attn_flags(self) -> int:
        return (0 if self.causal else _lib.ATTN_NONCAUSAL) | (_lib.ATTN_REF_ROUNDING if self.reference_rounding else 0)

    def _resolve_attn(self, dtype: int) -> int:
        if self.attn_impl == "simt":
            return _lib.ATTN_SIMT
        if self.attn_impl == "tensor":
            return _lib.ATTN_TENSOR
        dh = self.embedding_dim // self.num_head
        ok = dtype == ME_BF16 and _TENSOR_ATTENTION_AVAILABLE and dh in (32, 48, 64)
        return _lib.ATTN_TENSOR if ok else _lib.ATTN_SIMT

    def invalidate_weight_cache(self):
        self._wcache.clear()

    def _weights(self, dtype: int, refresh: bool = False) -> dict:
        """Packed device copies of the parameters in the compute type (QKV concatenated, bf16 casts).

        `refresh=True` (every full forward pass) re-derives them from the fp32 parameters in one batched
        launch -- the reference re-casts its parameters under autocast on every call as well, and parameter
        version counters cannot be trusted to see an optimiser step (torch's fused Adam does not bump them).
        Without `refresh` (backward of the same step, KV-cache decode steps) the copies are reused unless a
        version counter or a storage pointer moved (load_state_dict, .to())."""
        params = self._param_list()
        versions = tuple(p._version for p in params) + tuple(p.data_ptr() for p in params)
        wc = self._wcache.get(dtype)
        if wc is not None and not refresh and wc["versions"] == versions:
            return wc
        dev = params[0].device
        tdt = torch.bfloat16 if dtype == ME_BF16 else torch.float32
        stream = torch.cuda.current_stream().cuda_stream
        d, di, V = self.embedding_dim, self.d_inner, self.vocab_size
        ptrs = tuple(p.data_ptr() for p in params)
        if wc is None or wc["ptrs"] != ptrs:
            wc = {"layers": [], "ptrs": ptrs}
            for _ in range(self.num_layer):
                wc["layers"].append({
                    "Wqkv": torch.empty(3 * d, d, device=dev, dtype=tdt),
                    "bqkv": torch.empty(3 * d, device=dev, dtype=torch.float32),
                    "E": torch.empty(self.max_seq, d // self.num_head, device=dev, dtype=tdt),
                    "Wo": torch.empty(d, d, device=dev, dtype=tdt),
                    "W1": torch.empty(di, d, device=dev, dtype=tdt),
                    "W2": torch.empty(d, di, device=dev, dtype=tdt),
                })
            if self._vocab_head:
                wc["Wfc"] = torch.empty(V, d, device=dev, dtype=tdt)
            # table of (fp32 parameter -> compute-type copy) pairs for me_convert_batched
            entries = []

            def add(dst, src, dst_dtype):
                src = src.detach()
                rows, cols = (src.shape if src.dim() == 2 else (1, src.shape[0]))
                ld_dst = dst.shape[1] if dst.dim() == 2 else cols
                entries.append((src.data_ptr(), dst.data_ptr(), rows, cols, cols, ld_dst, ME_F32, dst_dtype))

            for l, lay in enumerate(self.enc_layers):
                w = wc["layers"][l]
                for i, lin in enumerate((lay.rga.Wq, lay.rga.Wk, lay.rga.Wv)):
                    add(w["Wqkv"][i * d:(i + 1) * d], lin.weight, dtype)
                    add(w["bqkv"][i * d:(i + 1) * d], lin.bias, ME_F32)
                add(w["E"], lay.rga.E, dtype)
                add(w["Wo"], lay.rga.fc.weight, dtype)
                add(w["W1"], lay.FFN_pre.weight, dtype)
                add(w["W2"], lay.FFN_suf.weight, dtype)
            if self._vocab_head:
                add(wc["Wfc"], self.fc.weight, dtype)
            table = (_lib.ConvertDesc * len(entries))(*[_lib.ConvertDesc(*e) for e in entries])
            raw = torch.frombuffer(bytearray(bytes(table)), dtype=torch.uint8)
            wc["table"] = raw.to(dev)
            wc["n"] = len(entries)
        _lib.call("me_convert_batched", ptr(wc["table"]), wc["n"], stream)
        wc["versions"] = versions
        self._wcache[dtype] = wc
        return wc

    def _pe(self, device) -> torch.Tensor:
        if self._pe_dev is None or self._pe_dev.device != device:
            self._pe_dev = positional_table(self.embedding_dim, self.max_seq).to(device)
        return self._pe_dev

    def _cond_params(self):
        """(cw0, cb0, cw1, cb1) pointers' tensors for the input stage."""
        if self.continuous_token:
            a, b = self.fc_condition[0], self.fc_condition[1]
            return a.weight, a.bias, b.weight, b.bias
        if self.d_condition > 0:
            return self.fc_condition.weight, self.fc_condition.bias, None, None
        return None, None, None, None

    # ------------------------------------------------------------------ forward
    def forward(self, x: torch.Tensor, condition: Optional[torch.Tensor] = None) -> torch.Tensor:
        """x: int64 [B, L] token ids (pad = 0); condition: float [B, 2] (valence, arousal), ignored
        (may be NaN) for none/discrete_token.  Returns logits [B, Ls, V] (Ls = L+2 for
        continuous_token): fp32 on the fp32 path, bf16 under bf16 autocast (as the reference)."""
        if not x.is_cuda:
            raise RuntimeError("midi_emotion_b200: CUDA tensors required (there is no CPU fallback)")
        if x.dtype != torch.int64 or x.dim() != 2:
            raise RuntimeError("midi_emotion_b200: x must be int64 [batch, sequence]")
        if self.mode != 0:
            if condition is None or condition.shape != (x.shape[0], 2):
                raise RuntimeError("midi_emotion_b200: condition must be float [batch, 2]")
            condition = condition.to(device=x.device, dtype=torch.float32).contiguous()
        else:
            condition = None
        from .autograd import model_apply
        return model_apply(self, x.contiguous(), condition)

    def loss(self, x: torch.Tensor, condition: Optional[torch.Tensor], target: torch.Tensor, ignore_index: int = 0,
             return_stats: bool = False):
        """The training step's `output = model(input_, condition); loss = ce_loss(output_flat, target)` (train.py:283-290)
        as ONE call: on the bf16 tensor-core path the output head is fused with the cross-entropy, so the
        [batch, sequence, vocabulary] logits are never materialised.  Same value and gradients as
        `cross_entropy(model(x, condition), target, ignore_index)`; `return_stats` adds the top-1 / top-5 counts of
        utils.accuracy (train.py:256).  On the fp32 path it is exactly that composition."""
        from .autograd import model_loss
        from .loss import cross_entropy
        if not x.is_cuda:
            raise RuntimeError("midi_emotion_b200: CUDA tensors required (there is no CPU fallback)")
        if self._resolve_dtype() != ME_BF16 or not self._vocab_head:
            return cross_entropy(self(x, condition), target, ignore_index, return_stats)
        if self.mode != 0:
            if condition is None or condition.shape != (x.shape[0], 2):
                raise RuntimeError("midi_emotion_b200: condition must be float [batch, 2]")
            condition = condition.to(device=x.device, dtype=torch.float32).contiguous()
        else:
            condition = None
        loss, stats = model_loss(self, x.contiguous(), condition, target, ignore_index)
        if return_stats:
            return loss, {"count": stats[1], "top1": stats[2], "top5": stats[3]}
        return loss

    def extra_repr(self) -> str:
        return (f"d_model={self.embedding_dim}, d_inner={self.d_inner}, d_condition={self.d_condition}, "
                f"layers={self.num_layer}, heads={self.num_head}, vocab={self.vocab_size}, "
                f"continuous_token={self.continuous_token}, dropout={self.dropout_p}")


_TENSOR_ATTENTION_AVAILABLE = True
_WARNED_FP16 = False


def set_dropout(model: MusicTransformer, rate: float) -> MusicTransformer:
    """models/build_model.py:2-7 equivalent: the model keeps a single dropout rate."""
    model.dropout_p = float(rate)
    return model
