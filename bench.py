#!/usr/bin/env python
"""Benchmark of the hot path: one training step (forward + CE + backward + clip + Adam) of the
emotion-conditioned MIDI transformer on synthetic token batches (BASELINE.json configs[1]):
continuous_concat 12L/768d/12h, seq_len 1024, bf16, batch 32 per GPU.

    python bench.py --gpus N --steps K --warmup W          # this repo (CUDA kernels via the C-ABI)
    python bench.py --impl reference ...                   # the reference algorithm on the host CPU

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG2 = dict(vocab_size=1007, n_layer=12, n_head=12, d_model=768, d_inner=3072, dropout=0.1, d_condition=192,
            conditioning="continuous_concat")
SEQ_LEN = 1024
BATCH_PER_GPU = 32
METRIC = "MIDI tokens/sec train fwd+bwd @ seq1024"
DEFAULT_OPTIM = "fused"
CFG3 = dict(vocab_size=1007, n_layer=24, n_head=16, d_model=1024, d_inner=4096, dropout=0.1, d_condition=192,
            conditioning="continuous_concat")


def workload(name):
    """BASELINE.json configs -> (model config, token length L, stack length Ls, sequences per GPU, metric, label).
    `cfg2` is the configuration the metric is quoted on; the others are extra lines."""
    if name == "cfg2":
        return dict(CFG2), SEQ_LEN, SEQ_LEN, BATCH_PER_GPU, METRIC, "continuous_concat 12L/768d/12h seq1024 " \
            "batch32/GPU bf16 (BASELINE configs[1])"
    if name == "cfg3":
        return dict(CFG3), 2048, 2048, 16, "MIDI tokens/sec train fwd+bwd @ seq2048", \
            "continuous_concat 24L/1024d/16h seq2048 batch16/GPU bf16 (BASELINE configs[2])"
    mode = name
    cfg = dict(CFG2, conditioning=mode, d_condition=-1)
    if mode == "discrete_token":       # ten emotion tokens appended to the vocabulary (loader.py:58-75)
        cfg["vocab_size"] = 1017
    # continuous_token prepends two condition positions (music_continuous_token.py:93-100; loader.py:55-57 shortens
    # the token window by two), so the stack still runs at 1024 positions
    L = SEQ_LEN - 2 if mode == "continuous_token" else SEQ_LEN
    return cfg, L, SEQ_LEN, BATCH_PER_GPU, METRIC, f"{mode} 12L/768d/12h seq1024 batch32/GPU bf16 (BASELINE configs[4] sweep)"


def flops_per_token(cfg, Ls):
    """Algorithmic forward FLOPs/token (causal-minimum attention); fwd+bwd = 3x."""
    d, di, NL, V = cfg["d_model"], cfg["d_inner"], cfg["n_layer"], cfg["vocab_size"]
    return NL * (8 * d * d + 4 * d * di + 3 * Ls * d) + 2 * d * V


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm_gbs=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sustained=j["bf16_tflops_sustained"],
                    source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/me_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=self.fh,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median over the samples taken under load (upper half of the observed clocks/power trace)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def synthetic_batch(cfg, B, L, seed, device=None, pin=False):
    """Tokens ~ U{1..1006}, <START>=1 first, no padding; target = next token; VA ~ U(-1,1).
    discrete_token: an emotion token (ids 1007..1016) leads the sequence and the condition is NaN
    (loader.py:58-75,185-187); none: NaN condition; continuous_token: target left-padded by two pads
    (loader.py:55-57: the two prepended condition positions predict nothing)."""
    g = torch.Generator().manual_seed(seed)
    mode = cfg["conditioning"]
    seq = torch.randint(1, 1007, (B, L + 1), generator=g)
    seq[:, 0] = 1
    if mode == "discrete_token":
        seq[:, 0] = torch.randint(1007, 1017, (B,), generator=g)
    tokens, target = seq[:, :-1].contiguous(), seq[:, 1:].contiguous()
    cond = torch.rand(B, 2, generator=g) * 2 - 1
    if mode in ("none", "discrete_token"):
        cond = torch.full_like(cond, float("nan"))
    if mode == "continuous_token":
        target = torch.cat([torch.zeros(B, 2, dtype=target.dtype), target], 1).contiguous()
    if pin:
        tokens, target, cond = tokens.pin_memory(), target.pin_memory(), cond.pin_memory()
    if device is not None:
        tokens, target, cond = tokens.to(device), target.to(device), cond.to(device)
    return tokens, cond, target


# ----------------------------------------------------------------------------------------------
# reference arm: the reference algorithm (oracle port) on the host cores
# ----------------------------------------------------------------------------------------------
def _reference_model(cfg):
    """The UNMODIFIED reference model package when /root/reference is on this machine (the build container), else
    None (the GPU box): the arm then times the oracle port of the same arithmetic."""
    src = "/root/reference/src"
    if not os.path.isdir(os.path.join(src, "models")):
        return None
    try:
        sys.path.insert(0, src)
        from models.build_model import build_model as ref_build   # noqa: E402
        torch.manual_seed(1234)
        model, _ = ref_build(dict(cfg, dropout=0.0))
        return model.train()
    except Exception:
        return None
    finally:
        if src in sys.path:
            sys.path.remove(src)


def cpu_train_tokens_per_s(steps, warmup, B=1, L=SEQ_LEN, threads=None, cfg=None):
    """The reference's training step on the host cores: forward, CE (ignore_index = pad), backward,
    clip_grad_norm_(1.0), torch.optim.Adam(lr 2e-5) -- train.py:276-292,307-325.  Returns (tokens/s, ms/step,
    threads, kind): kind "reference" = the reference's own nn.Module, "port" = the oracle restatement of its forward
    (both under torch autograd + torch.optim.Adam)."""
    from oracle import midi_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = dict(cfg or CFG2, dropout=0.0)
    tokens, cond, target = synthetic_batch(cfg, B, L, 1002)
    ref = _reference_model(cfg)
    if ref is not None:
        kind = "reference"
        params = list(ref.parameters())
        with torch.no_grad():
            for n, p in ref.named_parameters():
                if n.endswith("rga.E"):
                    p.mul_(0.2)

        def loss_of():
            out = ref(tokens, cond)
            return torch.nn.functional.cross_entropy(out.reshape(-1, out.size(-1)), target.reshape(-1), ignore_index=0)
    else:
        kind = "port"
        leaves = {k: v.clone().requires_grad_(True) for k, v in O.init_params(cfg, seed=1234, e_scale=0.2).items()}
        params = list(leaves.values())

        def loss_of():
            return O.loss_fn(O.forward(leaves, cfg, tokens, cond), target)
    opt = torch.optim.Adam(params, lr=2e-5)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        loss = loss_of()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    total = sum(times)
    return B * target.size(1) * len(times) / total, 1e3 * total / len(times), threads, kind


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    cfg, L, Ls, _, metric, label = workload(args.workload)
    # bounded sample: B=1 sequence of the same model/seq_len per step (the full batch of 32 is ~2 min/step)
    tps, ms, threads, kind = cpu_train_tokens_per_s(steps, warmup, B=1, L=L, cfg=cfg)
    sample = f"{steps} steps of batch 1 x seq {Ls} (same model, fp32, fwd+CE+bwd+clip_grad_norm_+torch.optim.Adam), {threads} threads"
    line = {
        "impl": "reference", "metric": metric, "value": tps, "unit": "tokens/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "train step (fwd+CE+bwd+clip+Adam) " + label,
                   "global_batch": 1, "seq_len": Ls,
                   "note": ("the unmodified reference model package on CPU" if kind == "reference" else
                            "oracle port of the reference PyTorch path on CPU (/root/reference is not on this machine)")},
        "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_decode_tokens_per_s(prefix=1024, B=4, steps=2, threads=None):
    """The reference's decode step on the host cores: generate.py:99-122 re-runs the model on the whole prefix for
    every generated token (no KV cache), so one step at prefix length t costs a full forward pass over [B, t].
    Timed at the same mid-sequence prefix as the GPU decode leg, on a reduced batch."""
    from oracle import midi_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = dict(CFG2, dropout=0.0)
    params = O.init_params(cfg, seed=1234, e_scale=0.2)
    g = torch.Generator().manual_seed(1004)
    tok = torch.randint(1, cfg["vocab_size"], (B, prefix), generator=g)
    tok[:, 0] = 1
    cond = torch.rand(B, 2, generator=g) * 2 - 1
    O.decode_last_logits(params, cfg, tok[:, :64], cond)          # warm-up (thread pool, allocator)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.decode_last_logits(params, cfg, tok, cond)
    dt = (time.perf_counter() - t0) / steps
    return {"value": B / dt, "unit": "tokens/s", "cores": threads, "kind": "port", "ms_per_step": 1e3 * dt,
            "sample": f"{steps} steps of the no-cache full-prefix recompute (generate.py:99-122) at prefix {prefix}, "
                      f"batch {B}, same model, fp32, {threads} threads"}


# ----------------------------------------------------------------------------------------------
# decode leg (BASELINE configs[3]): a real generation -- prefill + one KV-cache step per token to 2048
# ----------------------------------------------------------------------------------------------
def decode_leg(model, peaks, B=256, T=2048, t0=2):
    """generate.py:99-189 for 64 primers x 4 (valence, arousal) pairs = 256 sequences: prefill of the primer, then
    T - t0 KV-cache steps, each followed by the on-device sampling step (special symbols excluded, temperatures
    1.2 / 1.2, repeat penalty 0.5, top-p 0.7); nothing touches the host inside the loop.  `value`: generated
    tokens/s over the whole generation, device-timed; `e2e`: the same with the primer coming from pinned host memory
    and the generated tokens copied back inside the timed region.  HBM-algorithmic bytes: per step the weights once
    plus the K / V rows of every sequence up to the current position."""
    from midi_emotion_b200 import KVCacheDecoder, Sampler
    model.eval()
    torch.cuda.empty_cache()
    dev = next(model.parameters()).device
    V = CFG2["vocab_size"]
    g = torch.Generator().manual_seed(5)
    primers = torch.randint(5, V, (B // 4, t0), generator=g)
    primers[:, 0] = 1                                                    # <START>
    primer_host = primers.repeat_interleave(4, dim=0).contiguous().pin_memory()
    va = torch.tensor([[0.8, 0.8], [0.8, -0.8], [-0.8, 0.8], [-0.8, -0.8]])   # train.py:361-366
    cond_host = va.repeat(B // 4, 1).contiguous().pin_memory()
    exclude = torch.zeros(V, dtype=torch.uint8)
    exclude[:5] = 1
    dec = KVCacheDecoder(model, B, max_len=T, precision="bf16")
    out = torch.empty(B, T, device=dev, dtype=torch.int64)
    out_host = torch.empty(B, T, dtype=torch.int64).pin_memory()

    def generate_once(seed):
        sampler = Sampler(B, V, exclude=exclude, seed=seed)
        primer = primer_host.to(dev, non_blocking=True)
        cond = cond_host.to(dev, non_blocking=True)
        out[:, :t0] = primer
        logits = dec.prefill(primer, cond)
        prev = primer[:, -1].contiguous()
        for t in range(t0, T):
            nxt = sampler.sample(logits, prev)
            out[:, t] = nxt
            prev = out[:, t]
            if t + 1 < T:
                logits = dec.step(prev)

    # warm-up: a short generation (kernel attribute setup, CUDA-graph capture of the step)
    dec_T = T
    T = 16
    generate_once(1)
    T = dec_T
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    generate_once(9)
    e1.record()
    out_host.copy_(out, non_blocking=True)
    e2.record()
    torch.cuda.synchronize()
    ms_dev, ms_e2e = e0.elapsed_time(e1), e0.elapsed_time(e2)
    gen = B * (T - t0)
    d, NL = CFG2["d_model"], CFG2["n_layer"]
    w = NL * 12 * d * d * 2 + d * V * 2
    bytes_total = sum(w + B * NL * 2 * t * d * 2 for t in range(t0, T - 1))    # step at position t reads keys 0..t
    gbs = bytes_total / ms_dev / 1e6
    distinct = int(torch.unique(out_host[:, t0:]).numel())
    model.train()
    return {"metric": "decode tokens/sec @ seq2048 (prefill + 2046 KV-cache steps with on-device sampling, B=256)",
            "value": gen / ms_dev * 1e3, "unit": "tokens/s", "ms_total": ms_dev, "ms_per_step": ms_dev / (T - t0),
            "batch": B, "max_len": T, "primer_len": t0, "distinct_tokens_generated": distinct,
            "e2e": {"value": gen / ms_e2e * 1e3, "unit": "tokens/s", "h2d_bytes": primer_host.numel() * 8 + cond_host.numel() * 4,
                    "d2h_bytes": out_host.numel() * 8},
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": gbs / peaks["hbm_gbs"], "algorithmic_bytes_total": bytes_total}}


# ----------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------
KERNEL_CLASSES = {
    0: "gemm_tc_kernel / gemm_tc2_kernel (tcgen05 bf16 GEMMs: every Linear, dgrad and wgrad)",
    1: "attn_fwd_tc_kernel (relative attention forward)",
    2: "attn_bwd_tc_kernel (relative attention backward, key side: dK, dV)",
    3: "attn_bwd_q_tc_kernel (relative attention backward, query side: dQ, dE)",
}


def train_leg(args, CFG, L, Ls, B, K, W, dev, rank, world, lib, with_e2e=True, check_ddp=False):
    """Build the model, run W warm-up and K timed steps (inputs resident), then K steps end to end from pinned host
    buffers.  Returns a dict of raw measurements (rank-local times are reduced with MAX over ranks)."""
    import torch.distributed as dist
    from midi_emotion_b200 import ClipAdam, _lib, build_model, cross_entropy
    from midi_emotion_b200.ddp import DataParallel

    torch.manual_seed(1234)
    model, _ = build_model(dict(CFG))
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("rga.E"):
                p.mul_(0.2)   # N(0,1) E saturates the logits of a random-init deep model
    model = model.to(dev).train()
    model.attn_impl = args.attn
    ddp = DataParallel(model)
    # train.py:182 Adam (lr 2e-5, config.py:43) and the clip at 1.0 of train.py:321-322
    if args.optim == "fused":
        opt = ClipAdam(model.parameters(), lr=2e-5, max_grad_norm=1.0)
    else:
        opt = torch.optim.Adam(model.parameters(), lr=2e-5, fused=True)

    host = [synthetic_batch(CFG, B, L, 1002 + 17 * rank + i, pin=True) for i in range(2)]
    resident = [tuple(t.to(dev) for t in h) for h in host]

    def forward_backward(tokens, cond, target):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if args.loss == "head":    # output head fused with the loss: no [M, V] logits (me_head_cross_entropy)
                loss = model.loss(tokens, cond, target, ignore_index=0)
            else:
                logits = model(tokens, cond)
        if args.loss == "torch":   # the reference's nn.CrossEntropyLoss (train.py:124,288-290) on the logits
            loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.size(-1)).float(), target.reshape(-1),
                                                     ignore_index=0)
        elif args.loss == "fused":   # the same loss, fused with its gradient and the top-k counts (me_cross_entropy)
            loss = cross_entropy(logits, target, ignore_index=0)
        loss.backward()
        return loss

    def train_step(tokens, cond, target):
        loss = forward_backward(tokens, cond, target)
        ddp.sync_gradients()
        if args.optim == "torch":
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ddp_check = None
    if check_ddp and world > 1:
        # one-off: the overlapped, bucketed allreduce must leave on every rank the mean over ranks of the local
        # gradients.  Reference: a hook-less backward, then ONE plain allreduce of the concatenated gradients.
        p_drop = model.dropout_p
        model.dropout_p = 0.0
        hook = model._grad_ready_hook
        model._grad_ready_hook = None
        forward_backward(*resident[0])
        flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        dist.all_reduce(flat)
        flat /= world
        opt.zero_grad(set_to_none=True)
        model._grad_ready_hook = hook
        forward_backward(*resident[0])
        ddp.sync_gradients()
        got = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        err = float((got - flat).norm() / flat.norm().clamp_min(1e-30))
        same = torch.tensor([float(got.double().sum())], device=dev, dtype=torch.float64)
        gathered = [torch.zeros_like(same) for _ in range(world)]
        dist.all_gather(gathered, same)
        ddp_check = {"rel_l2_err_vs_plain_allreduce_mean": err, "ranks_hold_identical_gradients":
                     bool(all(float(x) == float(gathered[0]) for x in gathered)), "world": world,
                     "tolerance": 2e-2, "ok": bool(err < 2e-2),
                     "note": "bf16 split-K / atomic accumulation order differs between the two backward passes"}
        opt.zero_grad(set_to_none=True)
        model.dropout_p = p_drop

    for i in range(W):
        train_step(*resident[i % 2])
    barrier()

    # ---- timed region 1: inputs resident in HBM
    sampler = ClockSampler(dev.index)
    sampler.start()
    lib.me_profile_enable((64 + 8) * CFG["n_layer"] * K + 64)
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    h0 = time.perf_counter()
    for i in range(K):
        loss = train_step(*resident[i % 2])
    host_ms = 1e3 * (time.perf_counter() - h0)     # time the host needed to ENQUEUE the K steps (no sync inside)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    classes = {}
    for cls in sorted(KERNEL_CLASSES, reverse=True):    # class 0 last: me_profile_collect ends the collection
        c_ms, c_fl, c_n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
        if cls == 0:
            lib.me_profile_collect(ctypes.byref(c_ms), ctypes.byref(c_fl), ctypes.byref(c_n))
        else:
            lib.me_profile_collect_class(cls, ctypes.byref(c_ms), ctypes.byref(c_fl), ctypes.byref(c_n))
        classes[cls] = (c_ms.value, c_fl.value, c_n.value)
    clocks = sampler.stop()
    final_loss = float(loss.item())

    # ---- timed region 2: end to end from pinned host buffers, loss read back every step
    ms_e2e = None
    if with_e2e:
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for i in range(K):
            tok, cond, tgt = (t.to(dev, non_blocking=True) for t in host[i % 2])
            train_step(tok, cond, tgt).item()
        e3.record()
        barrier()
        ms_e2e = e2.elapsed_time(e3)

    if world > 1:
        t = torch.tensor([ms, ms_e2e or 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), (float(t[1]) if with_e2e else None)
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    return dict(model=model, ms=ms, ms_e2e=ms_e2e, host_ms=host_ms, launches=launches, classes=classes, clocks=clocks,
                final_loss=final_loss, h2d=h2d, ddp_check=ddp_check, opt=opt)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2",
                    choices=["cfg2", "cfg3", "discrete_token", "continuous_token", "continuous_concat", "none"],
                    help="cfg2 = BASELINE configs[1] (the quoted metric); cfg3 = configs[2]; a conditioning mode = "
                         "that row of the configs[4] sweep")
    ap.add_argument("--batch", type=int, default=None, help="sequences per GPU (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--attn", default="auto", choices=["auto", "simt", "tensor"])
    ap.add_argument("--no-decode", action="store_true", help="skip the KV-cache decode measurement (configs[3])")
    ap.add_argument("--no-cfg3-leg", action="store_true", help="skip the short configs[2] leg (24L/1024d, seq 2048)")
    ap.add_argument("--loss", default="head", choices=["head", "fused", "torch"],
                    help="head: output head fused with the cross-entropy (no logits tensor); fused: head GEMM + fused "
                         "cross-entropy kernel; torch: head GEMM + torch.nn.functional.cross_entropy")
    ap.add_argument("--optim", default=DEFAULT_OPTIM, choices=["torch", "fused"],
                    help="torch: clip_grad_norm_ + torch.optim.Adam(fused=True); fused: ClipAdam (csrc/optimizer.cu)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback)")

    import torch.distributed as dist
    from midi_emotion_b200 import _lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    W, K = max(3, args.warmup), max(1, args.steps)
    if args.workload == "continuous_concat":
        args.workload = "cfg2"
    CFG, L, Ls, B, metric, label = workload(args.workload)
    B = args.batch or B

    r = train_leg(args, CFG, L, Ls, B, K, W, dev, rank, world, lib, with_e2e=True, check_ddp=True)
    model, ms, ms_e2e = r["model"], r["ms"], r["ms_e2e"]

    line = None
    if rank == 0:
        peaks = load_peaks()
        tokens_per_step = world * B * Ls     # positions through the stack (continuous_token: two of them are the condition)
        value = tokens_per_step * K / (ms / 1e3)
        e2e = tokens_per_step * K / (ms_e2e / 1e3)
        fpt = 3 * flops_per_token(CFG, Ls)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tpath):   # dram__bytes_read+write per launch from the committed ncu --set full capture
            traffic = json.load(open(tpath))["mean_dram_bytes_per_launch"]
        by_kernel = []
        for cls, name in KERNEL_CLASSES.items():
            c_ms, c_fl, c_n = r["classes"][cls]
            tf = (c_fl / 1e12) / (c_ms / 1e3) if c_ms > 0 else 0.0
            by_kernel.append({"kernel": name, "launches": c_n, "ms_per_step": c_ms / K, "share_of_step": c_ms / ms,
                              "avg_launch_us": 1e3 * c_ms / c_n if c_n else None, "achieved": tf, "unit": "TFLOP/s",
                              "frac": tf / peaks["tf_sustained"]})
        g_ms, g_fl, g_n = r["classes"][0]
        gemm_tflops = (g_fl / 1e12) / (g_ms / 1e3) if g_ms > 0 else 0.0
        line = {
            "metric": metric, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "train step (fwd+CE+bwd+allreduce+clip+Adam) " + label,
                       "global_batch": world * B, "seq_len": Ls, "parallelism": f"dp{world}",
                       "l2": "per-step working set (activations > 10 GB) far exceeds the 126 MB L2; no flush needed",
                       "attention": args.attn, "loss": args.loss, "optimizer": args.optim,
                       "final_loss": r["final_loss"]},
            "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": int(r["launches"]),
            "host_enqueue_ms_per_step": r["host_ms"] / K,
            "clocks": r["clocks"],
            # the dominant kernel class of the step by time; every timed class is in `by_kernel`, the largest single
            # kernel (by average launch duration) among them in `largest_single_kernel`
            "roofline": {"bound": "tensor", "kernel": KERNEL_CLASSES[0],
                         "achieved": gemm_tflops, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": gemm_tflops / peaks["tf_sustained"], "peak_source": peaks["source"] + " sustained",
                         "frac_of_burst_peak": gemm_tflops / peaks["tf_burst"],
                         "traffic": traffic, "launches": g_n, "share_of_step": g_ms / ms,
                         "by_kernel": by_kernel,
                         "largest_single_kernel": max(by_kernel, key=lambda k: k["avg_launch_us"] or 0.0),
                         "whole_step_tflops": value * fpt / 1e12 / world,
                         "whole_step_frac": value * fpt / 1e12 / peaks["tf_sustained"] / world},
        }
        if r["ddp_check"] is not None:
            line["ddp_check"] = r["ddp_check"]
    # ---- extra legs (not the quoted metric)
    if args.workload == "cfg2" and world == 1 and not args.no_decode:
        dec = decode_leg(model, load_peaks())
        if not args.no_cpu_baseline:
            try:
                dec["cpu_baseline"] = cpu_decode_tokens_per_s()
            except Exception as e:   # a reported baseline must never cost the measured line
                dec["cpu_baseline"] = {"error": repr(e)}
        line["decode"] = dec
    if args.workload == "cfg2" and not args.no_cfg3_leg:
        # BASELINE configs[2] (24L/1024d/16h, seq 2048, DDP) as a short extra leg, so that the N = 1, 2, 4, 8 runs of
        # the driver carry it as well
        del r, model
        torch.cuda.empty_cache()
        c3, L3, Ls3, B3, metric3, label3 = workload("cfg3")
        r3 = train_leg(args, c3, L3, Ls3, B3, 4, 2, dev, rank, world, lib, with_e2e=False)
        if rank == 0:
            v3 = world * B3 * Ls3 * 4 / (r3["ms"] / 1e3)
            f3 = 3 * flops_per_token(c3, Ls3)
            line["cfg3"] = {"metric": metric3, "value": v3, "unit": "tokens/s", "ms_per_step": r3["ms"] / 4, "steps": 4,
                            "warmup": 2, "workload": label3, "global_batch": world * B3, "n_gpus": world,
                            "whole_step_frac": v3 * f3 / 1e12 / load_peaks()["tf_sustained"] / world,
                            "final_loss": r3["final_loss"]}
        del r3
        torch.cuda.empty_cache()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            tps, cms, threads, kind = cpu_train_tokens_per_s(steps=6, warmup=1, B=1, L=L, cfg=CFG)
            line["cpu_baseline"] = {"value": tps, "unit": "tokens/s", "cores": threads, "kind": kind,
                                    "sample": f"6 steps of batch 1 x seq {Ls}, same model, fp32, fwd+CE+bwd+clip+Adam, "
                                              f"{threads} threads", "ms_per_step": cms}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
