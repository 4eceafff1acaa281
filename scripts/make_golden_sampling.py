"""Golden vectors for the sampling step, produced by EXECUTING the reference's own lines.

src/generate.py does not import here (pretty_midi is absent), so the body of its generation loop after the
model call -- from `output = output[-1, :, :]` to the repeat-count update, generate.py:122-189 -- is cut out
of the source text and exec'd unmodified with stub vocabulary maps.  torch.multinomial is the only call
replaced (by an inverse-CDF draw with recorded uniforms: its generator stream cannot be reproduced by any
other implementation).  Output: tests/golden/sampling_*.npz.

    python scripts/make_golden_sampling.py        (needs /root/reference; run in the build container)
"""
import os
import textwrap

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/src/generate.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def reference_segment():
    lines = open(REF).read().split("\n")
    start = next(i for i, l in enumerate(lines) if "output = output[-1, :, :]" in l)
    end = next(i for i, l in enumerate(lines) if "else: repeat_counts[j] = repeat_counts[j] // 2" in l)
    return textwrap.dedent("\n".join(lines[start:end + 1])), (start + 1, end + 1)


def stub_maps(V):
    events = ["ON_PIANO", "ON_DRUMS", "TIMESHIFT", "ON_GUITAR", "ON_BASS", "ON_STRINGS"]
    idx2event = {i: e for i, e in enumerate(events)}
    symbols = ["<PAD>", "<START>", "<END>", "<V-2>", "<A1>"]
    idx2tuple = {}
    for i in range(V):
        if i < len(symbols):
            idx2tuple[i] = symbols[i]
        else:
            idx2tuple[i] = (i % len(events), i // len(events))
    tuple2idx = {v: k for k, v in idx2tuple.items()}
    return {"idx2tuple": idx2tuple, "tuple2idx": tuple2idx, "idx2event": idx2event}


def run_case(name, V, B, seed, temperatures, penalty_coeff, top_k, top_p, scale):
    code, span = reference_segment()
    maps = stub_maps(V)
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, V, generator=g) * scale
    logits[0, 5:9] = float("nan")
    prev = torch.randint(0, V, (B,), generator=g)
    prev[1] = 2 + 5 * 6            # a TIMESHIFT tuple (event index 2)
    repeat_counts = [0, 3, 7, 20, 1, 2][:B] + [0] * max(0, B - 6)
    uniforms = torch.rand(B, generator=g)
    exclude_symbols = [s for s in maps["tuple2idx"].keys() if s[0] == "<"]      # generate.py:57

    def draw(probs, n, replacement=True):
        cdf = torch.cumsum(probs, -1)
        out = torch.empty(probs.shape[0], 1, dtype=torch.int64)
        for j in range(probs.shape[0]):
            hit = torch.nonzero((cdf[j] > uniforms[j]) & (probs[j] > 0))
            out[j, 0] = int(hit[0]) if len(hit) else int(torch.nonzero(probs[j] > 0)[-1])
        return out

    ns = dict(torch=torch, F=F, np=np, output=logits.clone()[None], verbose=False, device="cpu",
              exclude_symbols=exclude_symbols, maps=maps, batch_size=B, gen_inds=prev[None].clone(),
              temperatures=list(temperatures), penalty_coeff=penalty_coeff, repeat_counts=list(repeat_counts),
              top_k=top_k, top_p=top_p)
    real = torch.multinomial
    torch.multinomial = draw
    try:
        exec(code, ns)
    finally:
        torch.multinomial = real
    probs = torch.zeros(B, V)
    probs.scatter_(1, ns["top_inds"], ns["output"])
    exclude = torch.zeros(V, dtype=torch.uint8)
    for s in exclude_symbols:
        exclude[maps["tuple2idx"][s]] = 1
    is_ts = torch.zeros(V, dtype=torch.uint8)
    for i, t in maps["idx2tuple"].items():
        if isinstance(t, tuple) and "TIMESHIFT" in maps["idx2event"][t[0]]:
            is_ts[i] = 1
    np.savez_compressed(
        os.path.join(OUT, f"sampling_{name}.npz"), logits=logits.numpy(), prev=prev.numpy(),
        repeat_counts=np.array(repeat_counts, dtype=np.int32), uniforms=uniforms.numpy(), exclude=exclude.numpy(),
        is_timeshift=is_ts.numpy(), temperatures=np.array(temperatures, dtype=np.float32),
        penalty_coeff=np.float32(penalty_coeff), top_k=np.int32(top_k), top_p=np.float32(top_p),
        tokens=ns["gen_inds"][0].numpy(), probs=probs.numpy(), num_choices=ns["num_choices"].numpy().astype(np.int32),
        new_repeat_counts=np.array(ns["repeat_counts"], dtype=np.int32), ref_lines=np.array(span))
    print(name, "lines", span, "tokens", ns["gen_inds"][0].tolist(), "choices", ns["num_choices"].tolist())


if __name__ == "__main__":
    run_case("default", 1007, 6, 1, (1.2, 1.2), 0.5, -1, 0.7, 3.0)
    run_case("topk_temps", 1017, 6, 2, (0.8, 1.5), 0.5, 40, 0.9, 2.0)
    run_case("no_topp_no_penalty", 1007, 4, 3, (1.0, 1.0), 0.0, -1, 1.0, 1.5)
    run_case("flat", 1007, 5, 4, (1.2, 1.2), 0.5, -1, 0.7, 0.05)
