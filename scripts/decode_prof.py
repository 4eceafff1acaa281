import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from midi_emotion_b200 import KVCacheDecoder, build_model
from torch.profiler import profile, ProfilerActivity
cfg = dict(vocab_size=1007, n_layer=12, n_head=12, d_model=768, d_inner=3072, dropout=0.1, d_condition=192, conditioning="continuous_concat")
model, _ = build_model(dict(cfg)); model = model.cuda().eval()
B, T = 256, 2048
dec = KVCacheDecoder(model, B, max_len=T, precision="bf16", use_cuda_graph=False)
cond = torch.rand(B, 2, device="cuda"); tok = torch.randint(1, 1007, (B, 4), device="cuda")
dec.prefill(tok, cond)
nxt = torch.randint(1, 1007, (B,), device="cuda")
for _ in range(3): dec.step(nxt)
dec.t_dev.fill_(int(os.environ.get("TPOS", "1024"))); dec.t_host = 1024
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): dec.step(nxt)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
