"""Micro-benchmark of the HBM-bound kernels at the cfg2 layer shape (M = 32768 rows, d = 768, d_inner = 3072):
add+LayerNorm forward / backward, column sums, input stage.  Prints algorithmic GB/s per kernel; also the
target of the ncu captures under profiles/."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from midi_emotion_b200 import _lib  # noqa: E402
from midi_emotion_b200._lib import ME_BF16, ptr  # noqa: E402

M, d, di, V, dc = int(os.environ.get("M", 32768)), 768, 3072, 1007, 192
iters = int(os.environ.get("ITERS", "20"))
st = torch.cuda.current_stream().cuda_stream
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)


def timeit(name, fn, nbytes):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:34s} {ms * 1e3:8.1f} us   {nbytes / ms / 1e6:8.1f} GB/s algorithmic")


x = torch.randn(M, d, device=dev, generator=g)
y = torch.randn(M, d, device=dev, generator=g).to(torch.bfloat16)
gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
out32, outT = torch.empty(M, d, device=dev), torch.empty(M, d, device=dev, dtype=torch.bfloat16)
z, mean, rstd = torch.empty(M, d, device=dev), torch.empty(M, device=dev), torch.empty(M, device=dev)
timeit("add+LayerNorm forward (p=0.1)",
       lambda: _lib.call("me_add_layernorm_forward", ptr(x), ptr(y), ME_BF16, ptr(gamma), ptr(beta), 1e-6, M, d, 0.1, 7,
                         ptr(out32), ptr(outT), ptr(z), ptr(mean), ptr(rstd), st), M * d * 16)
dout = torch.randn(M, d, device=dev, generator=g)
dz, dyT = torch.empty(M, d, device=dev), torch.empty(M, d, device=dev, dtype=torch.bfloat16)
dg, db = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
timeit("add+LayerNorm backward (p=0.1)",
       lambda: _lib.call("me_add_layernorm_backward", ptr(dout), None, ptr(z), ptr(mean), ptr(rstd), ptr(gamma), M, d, 0.1,
                         7, ME_BF16, ptr(dz), ptr(dyT), ptr(dg), ptr(db), st), M * d * 14)
gh = torch.randn(M, di, device=dev, generator=g).to(torch.bfloat16)
cs = torch.zeros(di, device=dev)
timeit("column sums [M, d_inner] bf16", lambda: _lib.call("me_colsum", ptr(gh), ME_BF16, M, di, di, ptr(cs), st), M * di * 2)
gq = torch.randn(M, 3 * d, device=dev, generator=g).to(torch.bfloat16)
cq = torch.zeros(3 * d, device=dev)
timeit("column sums [M, 3d] bf16", lambda: _lib.call("me_colsum", ptr(gq), ME_BF16, M, 3 * d, 3 * d, ptr(cq), st), M * 3 * d * 2)
ws = torch.empty(148 * di, device=dev)
timeit("column sums [M, d_inner] bf16, scratch", lambda: _lib.call("me_colsum_ws", ptr(gh), ME_BF16, M, di, di, ptr(cs), ptr(ws), ws.numel(), st), M * di * 2)
timeit("column sums [M, 3d] bf16, scratch", lambda: _lib.call("me_colsum_ws", ptr(gq), ME_BF16, M, 3 * d, 3 * d, ptr(cq), ptr(ws), ws.numel(), st), M * 3 * d * 2)
cs.zero_()
_lib.call("me_colsum_ws", ptr(gh), ME_BF16, M, di, di, ptr(cs), ptr(ws), ws.numel(), st)
print("scratch path max rel err vs torch:", float(((cs - gh.float().sum(0)).abs().max() / gh.float().sum(0).abs().max())))
B, L = M // 1024, 1024
tok = torch.randint(1, V, (B, L), device=dev, generator=g)
cond = torch.rand(B, 2, device=dev, generator=g)
emb = torch.randn(V, d - dc, device=dev, generator=g)
cw, cb = torch.randn(dc, 2, device=dev, generator=g), torch.zeros(dc, device=dev)
pe = torch.randn(2048, d, device=dev, generator=g)
keypad = torch.empty(B, L, device=dev, dtype=torch.uint8)
timeit("input stage (embed+cond+PE+dropout)",
       lambda: _lib.call("me_embed_forward", ptr(tok), ptr(cond), ptr(emb), ptr(cw), ptr(cb), None, None, ptr(pe), B, L, d,
                         dc, V, 3, 0, 0.1, 11, ME_BF16, ptr(out32), ptr(outT), ptr(keypad), st), M * d * 6)
