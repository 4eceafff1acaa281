"""Device-side token pipeline of the training loader (data/loader.py:132-195 + data/data_processing.py:225-247).

The reference prepares every sample in Python: `transpose` loops over the tuples, `tensor_to_ind_tensor` does one
dict lookup per tuple, then crop / prepend / pad.  Here the raw (event, value) tuples of a whole batch go to the
GPU once and one launch (`me_token_pipeline`) produces `input_` and `target` exactly as `Loader.__getitem__` +
`filter_collate` would for the same random decisions, which remain the caller's (bar window, n_transpose, the
bar-start coin and the crop offset are drawn from `random` / `np.random` in the reference, loader.py:108,127,137,151).

    pipe = TokenPipeline(maps, input_len=1024, conditioning="continuous_concat")
    input_, target = pipe(events=[int16 [n_i, 2] ...], n_transpose=[...], start=[... or -1],
                          emotion_tokens=[(v_tok, a_tok) or None ...])
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import TP_MAX_PREFIX, ptr


class TokenPipeline:
    def __init__(self, maps: dict, input_len: int, conditioning: str = "none", regression: bool = False,
                 use_cls_token: bool = False, start_token: Optional[str] = "<START>", pad_token: str = "<PAD>",
                 cls_token: str = "<CLS>", min_pitch: int = 21, max_pitch: int = 108, device="cuda"):
        """maps: the reference's mapping dict (data_processing.get_maps / maps.pt): 'tuple2idx' with (event index,
        value) -> id for tuple tokens and str -> id for symbols, and 'transposable_event_inds'.
        input_len: loader.py's tgt_len; continuous_token shortens it by the two condition positions (:55-57)."""
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("midi_emotion_b200: the token pipeline runs on a CUDA device (there is no CPU fallback)")
        t2i = maps["tuple2idx"]
        tuples = [(k, v) for k, v in t2i.items() if isinstance(k, tuple)]
        n_ev = max(k[0] for k, _ in tuples) + 1
        n_val = max(k[1] for k, _ in tuples) + 1
        lut = torch.full((n_ev, n_val), -1, dtype=torch.int32)
        for (e, v), idx in tuples:
            lut[e, v] = idx
        transposable = torch.zeros(n_ev, dtype=torch.uint8)
        for e in maps["transposable_event_inds"]:
            transposable[e] = 1
        self.lut, self.transposable = lut.to(self.dev), transposable.to(self.dev)
        self.n_ev, self.n_val = n_ev, n_val
        self.conditioning, self.regression = conditioning, bool(regression)
        self.input_len = int(input_len) - (2 if conditioning == "continuous_token" else 0)
        self.target_left_pad = 2 if conditioning == "continuous_token" else 0
        self.pad_id = int(t2i[pad_token])
        self.start_id = int(t2i[start_token]) if start_token is not None else None
        self.cls_id = int(t2i[cls_token]) if (regression and use_cls_token) else None
        self.min_pitch, self.max_pitch = int(min_pitch), int(max_pitch)

    def __call__(self, events: Sequence[torch.Tensor], n_transpose: Optional[Sequence[int]] = None,
                 start: Optional[Sequence[int]] = None, emotion_tokens: Optional[Sequence] = None):
        """events[i]: int16 [n_i, 2] tuples (CPU or CUDA).  start[i] >= 0: the sample does not start at a bar and is
        cropped at that offset (loader.py:150-152); -1 / None: it starts at the bar and <START> is prepended.
        emotion_tokens[i]: (valence id, arousal id) for discrete_token samples that carry them (loader.py:164-170).
        Returns (input_ int64 [B, input_len], target int64 [B, input_len (+2)] or None for regression)."""
        B = len(events)
        if B == 0:
            raise RuntimeError("midi_emotion_b200: empty batch")
        n = [int(e.shape[0]) for e in events]
        max_events = max(1, max(n))
        host = torch.zeros(B, max_events, 2, dtype=torch.int16)
        for i, e in enumerate(events):
            if n[i]:
                host[i, :n[i]] = e.to("cpu", torch.int16)
        prefix = torch.full((B, TP_MAX_PREFIX), -1, dtype=torch.int32)
        n_prefix = torch.zeros(B, dtype=torch.int32)
        st = torch.full((B,), -1, dtype=torch.int32) if start is None else torch.tensor(list(start), dtype=torch.int32)
        for i in range(B):
            pre: List[int] = []
            if emotion_tokens is not None and emotion_tokens[i] is not None:
                pre += [int(emotion_tokens[i][0]), int(emotion_tokens[i][1])]
            if self.cls_id is not None:
                pre.append(self.cls_id)
            if int(st[i]) < 0 and self.start_id is not None:
                pre.append(self.start_id)
            n_prefix[i] = len(pre)
            for k, t in enumerate(pre):
                prefix[i, k] = t
        tr = torch.zeros(B, dtype=torch.int32) if n_transpose is None else torch.tensor(list(n_transpose), dtype=torch.int32)
        dev = self.dev
        d_events = host.to(dev, non_blocking=True)
        d_n = torch.tensor(n, dtype=torch.int32).to(dev)
        d_tr, d_st, d_np, d_pre = tr.to(dev), st.to(dev), n_prefix.to(dev), prefix.to(dev)
        inp = torch.empty(B, self.input_len, device=dev, dtype=torch.int64)
        tgt = None if self.regression else torch.empty(B, self.input_len + self.target_left_pad, device=dev,
                                                        dtype=torch.int64)
        status = torch.empty(B, device=dev, dtype=torch.int32)
        a = _lib.TokenPipelineArgs()
        a.B, a.max_events, a.input_len, a.target_left_pad = B, max_events, self.input_len, self.target_left_pad
        a.pad_token, a.n_event_types, a.n_values = self.pad_id, self.n_ev, self.n_val
        a.min_pitch, a.max_pitch = self.min_pitch, self.max_pitch
        a.events, a.n_events, a.n_transpose, a.start = ptr(d_events), ptr(d_n), ptr(d_tr), ptr(d_st)
        a.n_prefix, a.prefix, a.transposable, a.lut = ptr(d_np), ptr(d_pre), ptr(self.transposable), ptr(self.lut)
        a.input, a.target, a.status = ptr(inp), ptr(tgt), ptr(status)
        a.stream = torch.cuda.current_stream().cuda_stream
        _lib.call("me_token_pipeline", C.byref(a))
        self.last_status = status
        return inp, tgt
