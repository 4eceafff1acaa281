"""cProfile of the host side of one training step (what the Python / ctypes layer costs per step)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from midi_emotion_b200 import ClipAdam, build_model  # noqa: E402

cfg, L, Ls, B, _, _ = bench.workload("cfg2")
torch.manual_seed(0)
model, _ = build_model(dict(cfg))
model = model.cuda().train()
opt = ClipAdam(model.parameters(), lr=2e-5, max_grad_norm=1.0)
tok, cond, tgt = bench.synthetic_batch(cfg, B, L, 1, device="cuda")


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = model.loss(tok, cond, tgt)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
