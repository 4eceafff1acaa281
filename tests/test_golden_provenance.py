"""The committed golden vectors really are outputs of the unmodified reference: where the reference tree is present
(the build container; it does not exist on the GPU box), its own model package is imported read-only and re-run on
the inputs stored in each tests/golden/*.npz, and must reproduce the stored logits, loss and gradients."""
import os
import sys

import pytest
import torch

REF = "/root/reference/src"

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")),
                                reason="the reference tree is only present in the build container")


@pytest.fixture(scope="module")
def ref_build_model():
    old_flag, sys.dont_write_bytecode = sys.dont_write_bytecode, True     # never write into /root/reference
    sys.path.insert(0, REF)
    try:
        from models.build_model import build_model
        yield build_model
    finally:
        sys.path.remove(REF)
        sys.dont_write_bytecode = old_flag
        for name in [m for m in sys.modules if m == "models" or m.startswith("models.")]:
            del sys.modules[name]


def test_reference_reproduces_the_golden_vectors(golden, ref_build_model):
    g = golden
    model, _ = ref_build_model(dict(g["cfg"]))
    missing = model.load_state_dict(g["params"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model.eval()
    out = model(g["tokens"], g["cond"])
    assert out.shape == g["logits_fp32"].shape
    torch.testing.assert_close(out.detach(), g["logits_fp32"], rtol=1e-5, atol=2e-6)
    assert torch.equal(out.argmax(-1), g["logits_fp32"].argmax(-1))
    loss = torch.nn.functional.cross_entropy(out.reshape(-1, out.size(-1)), g["target"].reshape(-1), ignore_index=0)
    assert float(loss.detach()) == pytest.approx(g["loss_fp32"], rel=1e-5)
    loss.backward()
    top = max(float(v.abs().max()) for v in g["grads"].values())
    for n, p in model.named_parameters():
        torch.testing.assert_close(p.grad, g["grads"][n], rtol=1e-4, atol=1e-6 * max(1.0, top), msg=lambda m: f"{n}: {m}")


def test_reference_reproduces_the_regression_goldens(regression_golden, ref_build_model):
    g = regression_golden
    model, _ = ref_build_model(dict(g["cfg"]))
    model.load_state_dict(g["params"], strict=True)
    model.eval()
    torch.testing.assert_close(model(g["tokens"]).detach(), g["out_fp32"], rtol=1e-5, atol=1e-6)
