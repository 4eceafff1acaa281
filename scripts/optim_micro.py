"""Micro-benchmark of the optimiser step (train.py:319-325) over the parameter tensors of the cfg2 model
(12L/768d, 88 M fp32 parameters in 197 tensors): ClipAdam (csrc/optimizer.cu) against clip_grad_norm_ +
torch.optim.Adam(fused=True).  Algorithmic bytes: 4 B/parameter for the norm pass, 28 B/parameter for the update
(p, g, m, v read; p, m, v written).  Also the target of the ncu capture under profiles/."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from midi_emotion_b200 import ClipAdam, build_model  # noqa: E402

iters = int(os.environ.get("ITERS", "20"))
cfg = dict(vocab_size=1007, n_layer=12, n_head=12, d_model=768, d_inner=3072, dropout=0.1, d_condition=192,
           conditioning="continuous_concat")
model, _ = build_model(cfg)
model = model.cuda()
params = list(model.parameters())
n = sum(p.numel() for p in params)
g = torch.Generator(device="cuda").manual_seed(0)
for p in params:
    p.grad = torch.randn(p.shape, device="cuda", generator=g) * 1e-3
print(f"{len(params)} tensors, {n / 1e6:.2f} M parameters")


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
    print(f"{name:52s} {ms * 1e3:8.1f} us   {nbytes / ms / 1e6:8.1f} GB/s algorithmic")


fused = ClipAdam(params, lr=2e-5, max_grad_norm=1.0)
plain = ClipAdam(params, lr=2e-5)
ref = torch.optim.Adam(params, lr=2e-5, fused=True)


def torch_step():
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    ref.step()


if os.environ.get("ONLY", "") != "torch":
    timeit("ClipAdam: norm + prepare + update", lambda: fused.step(), n * 32)
    timeit("ClipAdam: update only (no clipping)", lambda: plain.step(), n * 28)
if os.environ.get("ONLY", "") != "fused":
    timeit("clip_grad_norm_ + torch.optim.Adam(fused=True)", torch_step, n * 32 + n * 8)
