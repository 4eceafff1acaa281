import sys, faulthandler
faulthandler.enable()
print("start", flush=True)
import torch
print("torch", flush=True)
sys.path.insert(0, "/root/repo")
from midi_emotion_b200 import _lib
print("import _lib", flush=True)
lib = _lib.load()
print("loaded", lib.me_version(), flush=True)
x = torch.zeros(4, device="cuda")
print("cuda ok", flush=True)
print("sm100", lib.me_device_is_sm100(), flush=True)
import ctypes as C
from midi_emotion_b200._lib import ME_BF16, ME_F32, ptr
st = torch.cuda.current_stream().cuda_stream
print("stream", st, flush=True)
for (m, n, k) in [(256, 256, 256), (32768, 768, 768)]:
    A = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    B = torch.randn(n, k, device="cuda").to(torch.bfloat16)
    D = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    print("alloc", m, n, k, flush=True)
    _lib.call("me_gemm_bf16_ex", ptr(A), ptr(B), ptr(D), m, n, k, k, k, n, 0, 0, ME_BF16, 0, None, None, None, 0, 0, 0, st)
    print("launched", flush=True)
    torch.cuda.synchronize()
    print("gemm ok", m, n, k, float(D.float().abs().mean()), flush=True)
