import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith(".npz") and not f.startswith(("pe_", "sampling_", "ce_", "regression_", "tokens_")))


def sampling_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("sampling_") and f.endswith(".npz"))


def load_sampling_golden(name):
    import numpy as np
    import torch

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}
    g["temperatures"] = [float(x) for x in z["temperatures"]]
    return g


def load_golden(name):
    import numpy as np
    import torch

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    cfg = {}
    for k, v in zip(z["cfg_keys"], z["cfg_vals"]):
        k, v = str(k), str(v)
        if k == "conditioning":
            cfg[k] = v
        elif k == "dropout":
            cfg[k] = float(v)
        else:
            cfg[k] = int(v)
    params = {k[len("param::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param::")}
    grads = {k[len("grad::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad::")}
    g = dict(cfg=cfg, params=params, grads=grads,
             tokens=torch.from_numpy(z["tokens"]), cond=torch.from_numpy(z["cond"]),
             target=torch.from_numpy(z["target"]),
             logits_fp32=torch.from_numpy(z["logits_fp32"]), logits_bf16=torch.from_numpy(z["logits_bf16"]),
             loss_fp32=float(z["loss_fp32"]), pe_table=torch.from_numpy(z["pe_table"]),
             mask=torch.from_numpy(z["mask"]),
             decode={int(t): torch.from_numpy(z[f"decode_last::{int(t)}"]) for t in z["decode_prefix_lens"]})
    return g


@pytest.fixture(params=golden_names())
def golden(request):
    return load_golden(request.param)


def ce_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("ce_") and f.endswith(".npz"))


@pytest.fixture(params=ce_golden_names())
def ce_golden(request):
    import numpy as np
    import torch

    z = np.load(os.path.join(GOLDEN_DIR, request.param + ".npz"))
    return {k: (torch.from_numpy(z[k]) if z[k].ndim else z[k].item()) for k in z.files}


@pytest.fixture(params=sampling_golden_names())
def sampling_golden(request):
    return load_sampling_golden(request.param)


def regression_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("regression_") and f.endswith(".npz"))


@pytest.fixture(params=regression_golden_names())
def regression_golden(request):
    import numpy as np
    import torch

    z = np.load(os.path.join(GOLDEN_DIR, request.param + ".npz"))
    cfg = {}
    for k, v in zip(z["cfg_keys"], z["cfg_vals"]):
        k, v = str(k), str(v)
        cfg[k] = v if k == "conditioning" else (v == "True") if k == "regression" else float(v) if k == "dropout" else int(v)
    return dict(cfg=cfg, tokens=torch.from_numpy(z["tokens"]), out_fp32=torch.from_numpy(z["out_fp32"]),
                out_bf16=torch.from_numpy(z["out_bf16"]), loss_fp32=float(z["loss_fp32"]),
                params={k[len("param::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param::")},
                grads={k[len("grad::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad::")})
