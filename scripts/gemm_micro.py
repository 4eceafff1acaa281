"""Times the tcgen05 GEMM on the shapes of one cfg2 layer (forward, dgrad, wgrad) -- TFLOP/s per shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from midi_emotion_b200 import _lib  # noqa: E402
from midi_emotion_b200._lib import ME_BF16, ME_F32, ptr  # noqa: E402

M = int(os.environ.get("M", 32768))
d, di, V = 768, 3072, 1007
iters = int(os.environ.get("ITERS", "20"))
st = torch.cuda.current_stream().cuda_stream


def run(name, m, n, k, a_mn, b_mn, out_dtype, flags, tile_n=0, splits=0):
    A = torch.randn((k, m) if a_mn else (m, k), device="cuda").to(torch.bfloat16)
    B = torch.randn((k, n) if b_mn else (n, k), device="cuda").to(torch.bfloat16)
    D = torch.empty(m, n, device="cuda", dtype=torch.float32 if out_dtype == ME_F32 else torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    addend = torch.randn(m, n, device="cuda") if flags & _lib.EPI_ADD_F32 else None
    mask = torch.randn(m, n, device="cuda").to(torch.bfloat16) if flags & _lib.EPI_RELU_MASK else None

    def f():
        _lib.call("me_gemm_bf16_ex", ptr(A), ptr(B), ptr(D), m, n, k, A.stride(0), B.stride(0), n, a_mn, b_mn, out_dtype,
                  flags, ptr(bias), ptr(addend), ptr(mask), n if mask is not None else 0, tile_n, splits, st)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:28s} M={m:6d} N={n:5d} K={k:6d} {ms * 1e3:8.1f} us  {2.0 * m * n * k / ms / 1e9:7.1f} TFLOP/s")
    return ms


tot = 0.0
B_, R_, A_, K_ = _lib.EPI_BIAS, _lib.EPI_RELU, _lib.EPI_ADD_F32, _lib.EPI_RELU_MASK
tot += run("fwd qkv", M, 3 * d, d, 0, 0, ME_BF16, B_)
tot += run("fwd out-proj", M, d, d, 0, 0, ME_BF16, B_)
tot += run("fwd ffn1 (bias+relu)", M, di, d, 0, 0, ME_BF16, B_ | R_)
tot += run("fwd ffn2", M, d, di, 0, 0, ME_BF16, B_)
tot += run("dgrad ffn2 (relu mask)", M, di, d, 0, 1, ME_BF16, K_)
tot += run("dgrad ffn1 (+residual)", M, d, di, 0, 1, ME_F32, A_)
tot += run("dgrad out-proj", M, d, d, 0, 1, ME_BF16, 0)
tot += run("dgrad qkv (+residual)", M, d, 3 * d, 0, 1, ME_F32, A_)
tot += run("wgrad ffn2", d, di, M, 1, 1, ME_F32, 0)
tot += run("wgrad ffn1", di, d, M, 1, 1, ME_F32, 0)
tot += run("wgrad out-proj", d, d, M, 1, 1, ME_F32, 0)
tot += run("wgrad qkv", 3 * d, d, M, 1, 1, ME_F32, 0)
print(f"layer total {tot * 1e3:.1f} us")
run("head fwd", M, V, d, 0, 0, ME_BF16, B_)
if os.environ.get("WHATIF"):
    # what the dgrad GEMMs would cost with a pre-transposed (K-major) weight copy
    run("dgrad ffn2 (mask), K-major W^T", M, di, d, 0, 0, ME_BF16, K_)
    run("dgrad ffn1 (+res), K-major W^T", M, d, di, 0, 0, ME_F32, A_)
    run("dgrad out-proj, K-major W^T", M, d, d, 0, 0, ME_BF16, 0)
    run("dgrad qkv (+res), K-major W^T", M, d, 3 * d, 0, 0, ME_F32, A_)
    run("dgrad ffn1 bf16 out no res, MN", M, d, di, 0, 1, ME_BF16, 0)
