"""Generate tests/golden/*.npz from the UNMODIFIED reference model package.

Runs only in the build container (needs /root/reference).  It imports
/root/reference/src/models/{build_model,music_multi,music_continuous_token}.py read-only
(no bytecode written) and records, for small seeded configurations of all four
conditioning modes:

  state_dict, tokens, cond, target,
  logits_fp32, logits_bf16 (reference under torch.autocast('cpu', bfloat16)),
  loss_fp32, grads_fp32 (CE ignore_index=0, train.py:124,288-290,317),
  last-position logits for the no-cache decode contract (generate.py:99-122),
  the positional table and generate_mask for known-answer checks.

Usage:  PYTHONDONTWRITEBYTECODE=1 python scripts/make_golden.py
"""
import os
import sys

sys.dont_write_bytecode = True
import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/src"
sys.path.insert(0, REF)
from models.build_model import build_model  # noqa: E402  (the reference, unmodified)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

CASES = {
    # name: (conditioning, V, n_layer, n_head, d_model, d_inner, d_condition, B, L, tail_pad, e_scale)
    "none_tiny":       ("none",              67, 2, 2, 64, 128, -1, 2, 24, 5, 1.0),
    "discrete_tiny":   ("discrete_token",    77, 2, 2, 64, 128, -1, 2, 24, 4, 1.0),
    "ctoken_tiny":     ("continuous_token",  67, 2, 4, 64, 128, -1, 3, 22, 6, 1.0),
    "concat_tiny":     ("continuous_concat", 67, 2, 2, 64, 128, 16, 2, 24, 5, 1.0),
    # head dim 48 (reference default 768/16), vocab 1007 (not a multiple of 8), odd L
    "concat_dh48_v1007": ("continuous_concat", 1007, 1, 2, 96, 192, 24, 2, 37, 7, 0.25),
    # longer sequence spanning several attention tiles, 3 layers
    "concat_L160":     ("continuous_concat", 131, 2, 4, 128, 256, 32, 2, 160, 20, 0.2),
    "ctoken_L130":     ("continuous_token",  131, 2, 2, 64, 128, -1, 2, 130, 9, 0.2),
}


def make_batch(conditioning, V, B, L, tail_pad, seed):
    g = torch.Generator().manual_seed(seed)
    seq = torch.randint(1, V, (B, L + 1), generator=g)
    seq[:, 0] = 1
    npad = torch.randint(0, tail_pad + 1, (B,), generator=g)
    npad[0] = tail_pad                      # make sure padding is exercised
    for b in range(B):
        if npad[b] > 0:
            seq[b, L + 1 - int(npad[b]):] = 0
    tokens, target = seq[:, :-1].contiguous(), seq[:, 1:].contiguous()
    if conditioning in ("continuous_token", "continuous_concat"):
        cond = torch.rand(B, 2, generator=g) * 2 - 1
    else:
        cond = torch.full((B, 2), float("nan"))
    if conditioning == "continuous_token":
        target = F.pad(target, (2, 0), value=0)   # loader.py:184-187
    return tokens, cond, target


def main():
    os.makedirs(OUT, exist_ok=True)
    for idx, (name, c) in enumerate(CASES.items()):
        conditioning, V, NL, H, d, di, dc, B, L, tail_pad, e_scale = c
        cfg = dict(vocab_size=V, n_layer=NL, n_head=H, d_model=d, d_inner=di, dropout=0.0,
                   d_condition=dc, conditioning=conditioning)
        torch.manual_seed(1234 + idx)
        model, _ = build_model(dict(cfg))
        model.eval()
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith("rga.E"):
                    p.mul_(e_scale)
                if "bias" in n or "layernorm" in n:
                    # defaults are 0 / 1: perturb so bias and LN affine paths are exercised
                    p.add_(0.05 * torch.randn_like(p))
        tokens, cond, target = make_batch(conditioning, V, B, L, tail_pad, 1000 + idx)

        out = {}
        for k, v in model.state_dict().items():
            out["param::" + k] = v.detach().numpy().copy()
        out["tokens"], out["cond"], out["target"] = tokens.numpy(), cond.numpy(), target.numpy()

        # fp32 forward + loss + grads (train.py:276-292,317)
        model.zero_grad()
        logits = model(tokens, cond)
        loss = F.cross_entropy(logits.reshape(-1, logits.size(-1)), target.reshape(-1), ignore_index=0)
        loss.backward()
        out["logits_fp32"] = logits.detach().numpy().copy()
        out["loss_fp32"] = np.float32(loss.item())
        for n, p in model.named_parameters():
            out["grad::" + n] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()

        # bf16 autocast forward (train.py:281 / generate.py:116 with bf16 instead of fp16)
        with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
            lb = model(tokens, cond)
        out["logits_bf16"] = lb.float().numpy().copy()

        # no-cache decode contract: full-prefix forward, last position (generate.py:99-122)
        prefixes = sorted(set([1, 2, 3, L // 2, L - tail_pad if L - tail_pad > 0 else L]))
        out["decode_prefix_lens"] = np.array(prefixes, dtype=np.int64)
        with torch.no_grad():
            for t in prefixes:
                o = model(tokens[:, :t], cond)
                out[f"decode_last::{t}"] = o[:, -1, :].numpy().copy()

        # known answers for helpers
        import models.music_multi as mm
        out["pe_table"] = model.pos_encoding.positional_embedding[0, :256].numpy().copy()
        mtoks = F.pad(tokens, (2, 0), value=-1) if conditioning == "continuous_token" else tokens
        out["mask"] = mm.generate_mask(mtoks, 0).numpy().copy()

        out["cfg_keys"] = np.array(list(cfg.keys()))
        out["cfg_vals"] = np.array([str(v) for v in cfg.values()])
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: loss={loss.item():.6f} logits{tuple(logits.shape)} -> {os.path.getsize(path)/1024:.0f} KiB")

    # full-width positional table rows for d=768 (spot rows only: bit-parity of the fp64->fp32 recipe)
    import models.music_multi as mm
    rows = [0, 1, 2, 17, 255, 1023, 2047]
    full = np.asarray(mm.sinusoid(2048, 768))[0]
    np.savez_compressed(os.path.join(OUT, "pe_768_rows.npz"), rows=np.array(rows),
                        values=full[rows].astype(np.float32))
    print("pe_768_rows written")


if __name__ == "__main__":
    main()
