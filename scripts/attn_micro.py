"""Micro-benchmark of the tensor-core relative-attention kernels at the cfg2 layer shape
(B=32, H=12, L=1024, dh=64).  Used for ncu captures and quick timing between kernel edits."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from midi_emotion_b200 import _lib  # noqa: E402

B, H, L, dh = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (32, 12, 1024, 64)))
iters = int(os.environ.get("ITERS", "10"))
MS, d = 2048, H * dh
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B, L, 3, H, dh, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
E = (torch.randn(MS, dh, device="cuda", generator=g) * 0.2).to(torch.bfloat16)
out = torch.empty(B, L, d, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, L, device="cuda")
dout = torch.randn(B, L, d, device="cuda", generator=g).to(torch.bfloat16)
g_qkv = torch.empty_like(qkv)
dE = torch.zeros(MS, dh, device="cuda")
dsum = torch.empty(B, H, L, device="cuda")
dq_acc = torch.empty(_lib.load().me_attention_backward_workspace_floats(B, H, L, dh, 2048), device="cuda")
st = torch.cuda.current_stream().cuda_stream
a = _lib.AttnArgs()
a.dtype, a.impl = _lib.ME_BF16, _lib.ATTN_TENSOR
a.B, a.H, a.Lq, a.Lk, a.dh, a.max_seq, a.q_pos0 = B, H, L, L, dh, MS, 0
a.q, a.k, a.v, a.E = qkv.data_ptr(), qkv.data_ptr() + d * 2, qkv.data_ptr() + 2 * d * 2, E.data_ptr()
for n in "qkv":
    setattr(a, f"{n}_sb", L * 3 * d)
    setattr(a, f"{n}_sh", dh)
a.q_si = a.k_sj = a.v_sj = 3 * d
a.keypad, a.keypad_ld = None, L
a.out, a.o_sb, a.o_si = out.data_ptr(), L * d, d
a.lse, a.pos_dev, a.stream = lse.data_ptr(), None, st
if os.environ.get("SAVED", "1") == "1":
    tiles = B * H * _lib.load().me_attention_saved_tiles(L, 0)
    p_tiles = torch.empty(tiles * 128 * 64, device="cuda", dtype=torch.bfloat16)
    m_tiles = torch.empty(tiles * 128, device="cuda")
    a.p_tiles, a.m_tiles = p_tiles.data_ptr(), m_tiles.data_ptr()
ba = _lib.AttnBwdArgs()
ba.f = a
ba.dout = dout.data_ptr()
ba.dq, ba.dk, ba.dv = g_qkv.data_ptr(), g_qkv.data_ptr() + d * 2, g_qkv.data_ptr() + 2 * d * 2
ba.dE, ba.dsum, ba.dq_acc = dE.data_ptr(), dsum.data_ptr(), dq_acc.data_ptr()


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


fwd_flops = 3 * 2 * B * H * (L * L / 2) * dh     # causal-minimum: QK^T, QE^T, PV
t_f = timeit(lambda: _lib.call("me_attention_forward", C.byref(a)))
t_b = timeit(lambda: _lib.call("me_attention_backward", C.byref(ba)))
print(f"attention B={B} H={H} L={L} dh={dh}: fwd {t_f * 1e3:.1f} us ({fwd_flops / t_f / 1e9:.1f} TFLOP/s useful), "
      f"bwd {t_b * 1e3:.1f} us ({2.5 * fwd_flops / t_b / 1e9:.1f} TFLOP/s useful, incl. prep/convert)")

if os.environ.get("TRACE_FWD"):
    # clock stamps of the heaviest CTA of the forward kernel (ME_TRACE=1 build)
    buf = torch.zeros(4 * 20 * 8, dtype=torch.int64, device="cuda")
    _lib.call("me_debug_trace_set", buf.data_ptr())
    _lib.call("me_attention_forward", C.byref(a))
    torch.cuda.synchronize()
    _lib.call("me_debug_trace_set", None)
    t = buf.cpu().view(4, 20, 8)
    t0 = int(t[t > 0].min())
    for role, name in enumerate(("teamA", "teamB", "mma  ", "cta  ")):
        for st in range(20):
            row = t[role, st]
            if int(row.max()) == 0:
                continue
            print("fwd", name, "step", st, " ".join(f"{(int(v) - t0) if v > 0 else -1:7d}" for v in row[:8]))

if os.environ.get("TRACE"):
    # clock stamps of one CTA of the query-side backward kernel (ME_TRACE=1 build, me_debug_trace_set)
    buf = torch.zeros(2 * 3 * 20 * 8, dtype=torch.int64, device="cuda")
    _lib.call("me_debug_trace_set", buf.data_ptr())
    _lib.call("me_attention_backward", C.byref(ba))
    torch.cuda.synchronize()
    _lib.call("me_debug_trace_set", None)
    for kern, t in zip(("key-side", "query-side"), buf.cpu().view(2, 3, 20, 8)):
        if int(t.max()) == 0:
            continue
        t0 = int(t[t > 0].min())
        names = {0: "thread", 1: "mma   ", 2: "loader"}
        for role in range(3):
            for st in range(20):
                row = t[role, st]
                if int(row.max()) == 0:
                    continue
                print(kern, names[role], "step", st, " ".join(f"{(int(v) - t0) if v > 0 else -1:7d}" for v in row[:8]))
