"""Helpers shared by the -m gpu tests (imported only when CUDA is present)."""
import ctypes as C

import torch

from midi_emotion_b200 import _lib
from midi_emotion_b200._lib import ME_BF16, ME_F32, ptr


def stream():
    return torch.cuda.current_stream().cuda_stream


def dev():
    return torch.device("cuda:0")


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def gemm_bf16(A, B, M, N, K, a_mn, b_mn, out_dtype, flags=0, bias=None, addend=None, mask=None, tile_n=0, splits=0,
              ldd=None):
    """A, B: bf16 storage tensors (2-D, row-major, as the kernel sees them)."""
    ldd = ldd or N
    D = torch.full((M, ldd), float("nan"), device=A.device,
                   dtype=torch.float32 if out_dtype == ME_F32 else torch.bfloat16)
    _lib.call("me_gemm_bf16_ex", ptr(A), ptr(B), ptr(D), M, N, K, A.stride(0), B.stride(0), ldd, a_mn, b_mn,
              out_dtype, flags, ptr(bias), ptr(addend), ptr(mask), mask.stride(0) if mask is not None else 0,
              tile_n, splits, stream())
    return D


def record_parity(name, values, fname="parity_small_r02.json"):
    """Append measured parity numbers to gpurun_out/<fname> (copied to profiles/ after a GPU run)."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, fname)
    try:
        data = json.load(open(path))
    except Exception:
        data = {}
    data[name] = values
    with open(path, "w") as fh:
        json.dump(data, fh, indent=1, sort_keys=True)
