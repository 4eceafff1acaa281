"""Builds the C-ABI library (csrc/*.cu -> lib/libmidi_emotion_b200.so) with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the .so stays in-tree (git-ignored) so it travels to the GPU box.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libmidi_emotion_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", INCLUDE,
]
if os.environ.get("ME_EXP"):        # throw-away experiment switches for kernel tuning (-DME_EXP=<n>)
    NVCC_FLAGS.append("-DME_EXP=" + os.environ["ME_EXP"])
if os.environ.get("ME_TRACE") == "1":   # per-phase clock stamps in the attention backward kernel (tuning builds)
    NVCC_FLAGS.append("-DME_ATTN_TRACE")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the midi_emotion_b200 CUDA library cannot be built")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    """Hash of the sources and flags, independent of where the tree lives: the library built in the build
    container must be recognised as current in the copy of the tree the GPU box runs from (a different absolute
    path), or every process there would rebuild it -- all ranks of a torchrun launch at once."""
    h = hashlib.sha256()
    files = sources() + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")]
    files += [os.path.join(INCLUDE, f) for f in sorted(os.listdir(INCLUDE))]
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode() + b"\0" + fh.read())
    h.update(" ".join(x for x in NVCC_FLAGS if x != INCLUDE).encode())
    return h.hexdigest()


def _is_current(stamp: str, dig: str) -> bool:
    try:
        return os.path.exists(LIB) and open(stamp).read().strip() == dig
    except OSError:
        return False


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile when sources changed; returns the library path.  Safe to call from several processes at once (one
    builds under a file lock, the others wait and find the result); the library appears atomically."""
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.sha256")
    dig = _digest()
    if not force and _is_current(stamp, dig):
        return LIB
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _is_current(stamp, dig):   # another process built it while this one waited
                return LIB
            return _build_locked(stamp, dig, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(stamp: str, dig: str, verbose: bool) -> str:
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
           "-Xlinker", "--no-undefined", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if os.path.exists(stamp):
        os.remove(stamp)          # never a current-looking stamp next to a half-replaced library
    os.replace(tmp, LIB)
    with open(stamp + ".tmp", "w") as fh:
        fh.write(dig)
    os.replace(stamp + ".tmp", stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
