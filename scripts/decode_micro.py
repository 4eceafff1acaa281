"""Decode-step latency of the KV-cache path at BASELINE configs[3] (B=256, 12L/768d, T=2048)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from midi_emotion_b200 import KVCacheDecoder, build_model  # noqa: E402

B = int(os.environ.get("B", 256))
T = int(os.environ.get("T", 2048))
cfg = dict(vocab_size=1007, n_layer=12, n_head=12, d_model=768, d_inner=3072, dropout=0.1, d_condition=192,
           conditioning="continuous_concat")
t_start = time.time()
torch.manual_seed(0)
model, _ = build_model(dict(cfg))
with torch.no_grad():
    for n, p in model.named_parameters():
        if n.endswith("rga.E"):
            p.mul_(0.2)
model = model.cuda().eval()
print(f"model built in {time.time() - t_start:.1f} s", flush=True)
dec = KVCacheDecoder(model, B, max_len=T, precision="bf16", use_cuda_graph=os.environ.get("GRAPH", "1") == "1")
cond = torch.rand(B, 2, device="cuda") * 2 - 1
tok = torch.randint(1, 1007, (B, 4), device="cuda")
dec.prefill(tok, cond)
for c in dec.k_cache + dec.v_cache:
    c.normal_(0, 0.5)
nxt = torch.randint(1, 1007, (B,), device="cuda")
for _ in range(3):
    dec.step(nxt)     # eager, capture, first replay
torch.cuda.synchronize()
print(f"decoder ready at {time.time() - t_start:.1f} s", flush=True)
d, NL = 768, 12
for t in [int(x) for x in os.environ.get("TS", "128,512,1024,1536,2040").split(",")]:
    n = 5
    dec.t_dev.fill_(t)
    dec.t_host = t
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        dec.step(nxt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    kv_bytes = B * NL * 2 * (t + n / 2) * d * 2
    w_bytes = NL * 12 * d * d * 2 + d * 1007 * 2
    print(f"t={t:5d}: {ms:7.3f} ms/step  {B / ms * 1e3:9.0f} tok/s   HBM-algorithmic {(kv_bytes + w_bytes) / ms / 1e6:7.1f} GB/s")
