"""Emotion regression side model: the Python surface of models/music_regression.py:34-91 (MusicRegression, selected
by models/build_model.py:29-32 with output_size 2) on the same sm_100a kernels as the generative model.

Differences from the generative stack, all restated from the reference:
  * no condition branch (`assert d_condition <= 0`, :42), embedding * sqrt(d) (:81-82);
  * `no_mask=True` is the constructor default (:38,78): NO mask at all -- every position attends to every
    position, pads included (ME_ATTN_NONCAUSAL, no key-pad bytes); the relative term keeps its lower-triangular
    support (`_qe_masking`, :256-262);
  * the first position is pooled through Linear(d, output_size) + tanh (:65-68,89): `fc` is a Sequential, so the
    checkpoint keys are fc.0.weight / fc.0.bias;
  * forward(x) takes the tokens only and returns [B, output_size] (train.py:282-284 feeds it to an L1 loss
    against the (valence, arousal) condition).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib
from ._lib import ME_BF16, ptr
from .transformer import MAX_SEQ, MusicTransformer


class MusicRegression(MusicTransformer):
    def __init__(self, embedding_dim=None, d_inner=None, vocab_size=None, num_layer=None, num_head=None,
                 max_seq=MAX_SEQ, dropout=0.0, pad_token=0, output_size=2, d_condition=-1, no_mask=True):
        assert d_condition is None or d_condition <= 0, "the regression model has no condition branch"
        super().__init__(embedding_dim=embedding_dim, d_inner=d_inner, d_condition=-1, vocab_size=vocab_size,
                         num_layer=num_layer, num_head=num_head, max_seq=max_seq, dropout=dropout,
                         pad_token=pad_token, continuous_token=False)
        if output_size > 8:
            raise ValueError("output_size <= 8")
        self.output_size = int(output_size)
        self.no_mask = bool(no_mask)
        # music_regression.py:65-68; default nn.Linear init (init_weights only touches the embedding, :71-73)
        self.fc = nn.Sequential(nn.Linear(embedding_dim, output_size), nn.Tanh())
        self.__dict__.pop("_param_list_cache", None)   # (the head was replaced after the base constructor ran)
        self.__dict__.pop("_param_names_cache", None)
        self._vocab_head = False
        self.causal = not self.no_mask       # no_mask=False would be generate_mask: causal + key pads
        self.use_keypad = not self.no_mask

    def forward(self, x: torch.Tensor) -> torch.Tensor:   # noqa: D401  (reference signature: tokens only)
        """x: int64 [B, L] token ids.  Returns [B, output_size] in (-1, 1): fp32 on the fp32 path, bf16 under
        autocast (as the reference)."""
        if not x.is_cuda:
            raise RuntimeError("midi_emotion_b200: CUDA tensors required (there is no CPU fallback)")
        if x.dtype != torch.int64 or x.dim() != 2:
            raise RuntimeError("midi_emotion_b200: x must be int64 [batch, sequence]")
        params = self._param_list()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _RegressionFn.apply(self, x.contiguous(), need_grad, *params)


class _RegressionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, tokens, need_grad, *params):
        from .autograd import _stream, run_forward
        _, a = run_forward(model, tokens, None, need_grad, head=False)
        last = a.layers[-1]
        B, Ls, d = a.B, a.Ls, model.embedding_dim
        lin = model.fc[0]
        out = torch.empty(B, model.output_size, device=tokens.device, dtype=torch.float32)
        x_last = last["out2_T"] if a.dtype == ME_BF16 else last["out2_f32"]
        _lib.call("me_pooled_head_forward", ptr(x_last), a.dtype, ptr(lin.weight), ptr(lin.bias), B, Ls, d,
                  model.output_size, ptr(out), _stream())
        ctx.model, ctx.tokens, ctx.acts, ctx.out = model, tokens, a, out
        return out.to(torch.bfloat16) if a.dtype == ME_BF16 else out

    @staticmethod
    def backward(ctx, g_out):
        from .autograd import _stream, backward_stack, _grouper
        model, a, out = ctx.model, ctx.acts, ctx.out
        if a.layers is None or len(a.layers) != model.num_layer or a.layers[0] is None or "z1" not in a.layers[0]:
            raise RuntimeError("midi_emotion_b200: backward called twice or forward ran without grad")
        B, Ls, M, d = a.B, a.Ls, a.M, model.embedding_dim
        dev = ctx.tokens.device
        last = a.layers[-1]
        x_last = last["out2_T"] if a.dtype == ME_BF16 else last["out2_f32"]
        lin = model.fc[0]
        flat, g = _grouper(dev)({"fc.0.weight": (model.output_size, d), "fc.0.bias": (model.output_size,)})
        d_x = torch.zeros(M, d, device=dev, dtype=torch.float32)
        _lib.call("me_pooled_head_backward", ptr(g_out.float().contiguous()), ptr(out), ptr(x_last), a.dtype,
                  ptr(lin.weight), B, Ls, d, model.output_size, ptr(g["fc.0.weight"]), ptr(g["fc.0.bias"]), ptr(d_x),
                  _stream())
        grads = dict(g)
        hook = getattr(model, "_grad_ready_hook", None)
        if hook is not None:
            hook(flat)
        backward_stack(model, ctx.tokens, None, a, d_x, grads)
        ctx.acts = None
        from .autograd import _param_names
        return (None, None, None, *[grads[n] for n in _param_names(model)])
