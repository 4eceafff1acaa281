"""On-device next-token sampling and generation loop (replaces generate.py:99-189).

The reference post-processes the last-position logits with a dozen PyTorch calls and two Python loops
over the batch that read 2*B scalars back per generated token; here the whole step is one kernel
(`me_sample_step`) and the loop `KVCacheDecoder.step -> sample -> next step` never touches the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ME_BF16, ME_F32, ptr


class Sampler:
    """State of generate.py's sampling rules for a batch: excluded symbols, TIMESHIFT table, repeat counts."""

    def __init__(self, batch_size: int, vocab_size: int, exclude: Optional[torch.Tensor] = None,
                 is_timeshift: Optional[torch.Tensor] = None, temperatures: Sequence[float] = (1.2, 1.2),
                 penalty_coeff: float = 0.5, top_k: int = -1, top_p: float = 0.7, device="cuda", seed: int = 0):
        self.B, self.V = batch_size, vocab_size
        dev = torch.device(device)
        u8 = dict(device=dev, dtype=torch.uint8)
        self.exclude = torch.zeros(vocab_size, **u8) if exclude is None else exclude.to(**u8).contiguous()
        self.is_timeshift = (torch.zeros(vocab_size, **u8) if is_timeshift is None
                             else is_timeshift.to(**u8).contiguous())
        temps = list(temperatures)
        self.temp_note, self.temp_rest = (temps[0], temps[0]) if len(temps) == 1 else (temps[0], temps[1])
        self.penalty_coeff, self.top_k, self.top_p = float(penalty_coeff), int(top_k), float(top_p)
        self.repeat_counts = torch.zeros(batch_size, device=dev, dtype=torch.int32)
        self.num_choices = torch.zeros(batch_size, device=dev, dtype=torch.int32)
        self.tokens = torch.zeros(batch_size, device=dev, dtype=torch.int64)
        self.gen = torch.Generator(device=dev).manual_seed(seed)
        self.uniforms = torch.empty(batch_size, device=dev, dtype=torch.float32)

    def sample(self, logits: torch.Tensor, prev_tokens: torch.Tensor, uniforms: Optional[torch.Tensor] = None,
               out_probs: Optional[torch.Tensor] = None) -> torch.Tensor:
        """logits [B, >=V] fp32/bf16 (row pitch = stride(0)); prev_tokens int64 [B].  Returns int64 [B]
        (a buffer owned by the sampler, overwritten by the next call)."""
        if not logits.is_cuda:
            raise RuntimeError("midi_emotion_b200: CUDA tensors required (there is no CPU fallback)")
        if logits.dim() != 2 or logits.shape[0] != self.B or logits.stride(1) != 1:
            raise RuntimeError("midi_emotion_b200: logits must be [batch, vocab] with contiguous rows")
        dt = {torch.float32: ME_F32, torch.bfloat16: ME_BF16}.get(logits.dtype)
        if dt is None:
            raise RuntimeError("midi_emotion_b200: logits must be float32 or bfloat16")
        if uniforms is None:
            self.uniforms.uniform_(0.0, 1.0, generator=self.gen)
            uniforms = self.uniforms
        a = _lib.SampleArgs()
        a.B, a.V, a.ld_logits, a.logits_dtype = self.B, self.V, logits.stride(0), dt
        a.logits, a.exclude, a.is_timeshift = ptr(logits), ptr(self.exclude), ptr(self.is_timeshift)
        a.prev_tokens = ptr(prev_tokens.contiguous())
        a.temp_note, a.temp_rest, a.penalty_coeff = self.temp_note, self.temp_rest, self.penalty_coeff
        a.top_k, a.top_p = self.top_k, self.top_p
        a.repeat_counts, a.uniforms = ptr(self.repeat_counts), ptr(uniforms.contiguous())
        a.out_tokens, a.out_num_choices, a.out_probs = ptr(self.tokens), ptr(self.num_choices), ptr(out_probs)
        a.stream = torch.cuda.current_stream().cuda_stream
        _lib.call("me_sample_step", C.byref(a))
        return self.tokens


def generate(model, primer: torch.Tensor, condition: Optional[torch.Tensor], gen_len: int, sampler: Sampler,
             max_len: Optional[int] = None, precision: str = "bf16", max_input_len: Optional[int] = None,
             discrete_conditions: Optional[torch.Tensor] = None, varying_condition=None) -> torch.Tensor:
    """Autoregressive generation: returns int64 [B, primer_len + gen_len].  The model call and the sampling rules are
    those of generate.py:99-189; no device-to-host transfer happens inside the loop.

    max_input_len       generate.py:101-103: the model only ever sees the last `max_input_len` tokens (the CLI default
                        is 1216).  While the song is shorter, one KV-cache step per token; once the window slides
                        every position moves (absolute sinusoid), cached keys are invalid, and each token costs a
                        re-prefill of the window -- exactly the reference's full-prefix recompute, so the result
                        stays equal to it.  Default: the model's max_seq (less the condition positions).
    discrete_conditions int64 [B, n] emotion tokens kept in front of the window (generate.py:105-107, 78-80).
    varying_condition   (valences [B, gen_len], arousals [B, gen_len]): interpolated conditions, generate.py:110-113;
                        the condition feeds every position, so every step is a re-prefill.
    max_len             cache length (default: what the window needs)."""
    from .decode import KVCacheDecoder

    B, t0 = primer.shape
    dev = primer.device
    n_disc = 0 if discrete_conditions is None else int(discrete_conditions.shape[1])
    extra = 2 if model.continuous_token else 0
    window = int(max_input_len) if max_input_len is not None else model.max_seq
    window = min(window, model.max_seq) - extra - n_disc          # generate.py:75-80
    if window <= 0:
        raise ValueError("max_input_len leaves no room for tokens")
    need = min(t0 + gen_len, window) + extra + n_disc
    dec = KVCacheDecoder(model, B, max_len=max(need, 1) if max_len is None else max_len, precision=precision)
    out = torch.empty(B, t0 + gen_len, device=dev, dtype=torch.int64)
    out[:, :t0] = primer
    cond = condition
    prev = primer[:, -1].contiguous()
    logits = None
    for i in range(gen_len):
        n = t0 + i                                   # tokens generated so far (the model input before the cut)
        if varying_condition is not None:
            cond = torch.stack([varying_condition[0][:, i], varying_condition[1][:, i]], dim=-1).to(dev)
        if i == 0 or n > window or varying_condition is not None:
            inp = out[:, max(0, n - window):n]
            if n_disc:
                inp = torch.cat([discrete_conditions.to(dev), inp], dim=1)
            logits = dec.prefill(inp.contiguous(), cond)
        else:
            logits = dec.step(prev)
        nxt = sampler.sample(logits, prev)
        out[:, n] = nxt
        prev = out[:, n]
    return out
