"""Golden vectors for the oracle's restatement of the regression side model (models/music_regression.py, SURVEY.md
8f rank 4), produced by the UNMODIFIED reference package: build_model({"regression": True, ...}).

Usage:  PYTHONDONTWRITEBYTECODE=1 python scripts/make_golden_regression.py     (build container only)
"""
import os
import sys

sys.dont_write_bytecode = True
import numpy as np
import torch

sys.path.insert(0, "/root/reference/src")
from models.build_model import build_model  # noqa: E402  (the reference, unmodified)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
CASES = {
    # name: (V, n_layer, n_head, d_model, d_inner, B, L, tail_pad, e_scale)
    "regression_tiny": (67, 2, 2, 64, 128, 3, 24, 6, 1.0),
    "regression_dh48_L70": (1007, 1, 2, 96, 192, 2, 70, 9, 0.25),
}


def main():
    for idx, (name, (V, NL, H, d, di, B, L, tail_pad, e_scale)) in enumerate(CASES.items()):
        cfg = dict(vocab_size=V, n_layer=NL, n_head=H, d_model=d, d_inner=di, dropout=0.0, d_condition=-1,
                   conditioning="none", regression=True)
        torch.manual_seed(4321 + idx)
        model, _ = build_model(dict(cfg))
        model.eval()
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith("rga.E"):
                    p.mul_(e_scale)
                if "bias" in n or "layernorm" in n:
                    p.add_(0.05 * torch.randn_like(p))
        g = torch.Generator().manual_seed(2000 + idx)
        tokens = torch.randint(1, V, (B, L), generator=g)
        tokens[:, 0] = 1
        tokens[0, L - tail_pad:] = 0                      # pads are NOT masked by this model: they must matter
        out = model(tokens)
        loss = ((out - torch.tensor([[0.8, -0.8]])) ** 2).mean()
        loss.backward()
        rec = {"cfg_keys": np.array(list(cfg.keys())), "cfg_vals": np.array([str(v) for v in cfg.values()]),
               "tokens": tokens.numpy(), "out_fp32": out.detach().numpy(), "loss_fp32": np.float32(loss.item())}
        with torch.autocast("cpu", dtype=torch.bfloat16):
            rec["out_bf16"] = model(tokens).detach().float().numpy()
        for n, p in model.state_dict().items():
            rec["param::" + n] = p.detach().numpy()
        for n, p in model.named_parameters():
            rec["grad::" + n] = p.grad.detach().numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, out.detach().numpy().round(4).tolist())


if __name__ == "__main__":
    main()
