"""Sweep (tile_n, splits) for the weight-gradient GEMM shapes; prints the best configuration per shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from midi_emotion_b200 import _lib  # noqa: E402
from midi_emotion_b200._lib import ME_F32, ptr  # noqa: E402

st = torch.cuda.current_stream().cuda_stream
K = 32768
for (m, n) in [(768, 3072), (3072, 768), (768, 768), (2304, 768), (1007, 768), (1024, 4096), (3072, 1024)]:
    A = torch.randn(K, (m + 7) // 8 * 8, device="cuda").to(torch.bfloat16)[:, :m]
    B = torch.randn(K, n, device="cuda").to(torch.bfloat16)
    D = torch.empty(m, n, device="cuda")
    res = []
    for bn in (128, 256):
        for sp in (1, 2, 3, 4, 6, 8):
            def f():
                _lib.call("me_gemm_bf16_ex", ptr(A), ptr(B), ptr(D), m, n, K, A.stride(0), B.stride(0), n, 1, 1, ME_F32,
                          0, None, None, None, 0, bn, sp, st)
            for _ in range(2):
                f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                f()
            e1.record()
            torch.cuda.synchronize()
            res.append((e0.elapsed_time(e1) / 10 * 1e3, bn, sp))
    res.sort()
    mt = (m + 127) // 128
    print(f"wgrad {m}x{n}: " + "  ".join(f"bn{b}/s{s}:{t:.0f}us(ctas {mt * ((n + b - 1) // b) * s})" for t, b, s in res[:5]))
