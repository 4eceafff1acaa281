"""Pins oracle/midi_oracle.py against golden vectors generated from the unmodified reference
(scripts/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle import midi_oracle as O


def test_forward_fp32_matches_reference(golden):
    g = golden
    logits = O.forward(g["params"], g["cfg"], g["tokens"], g["cond"])
    ref = g["logits_fp32"]
    assert logits.shape == ref.shape
    # same op types in the same order -> expected bit-identical; allow 2e-6 for BLAS blocking
    assert torch.allclose(logits, ref, rtol=0, atol=2e-6), (logits - ref).abs().max()
    assert torch.equal(logits.argmax(-1), ref.argmax(-1))


def test_forward_bf16_autocast_matches_reference(golden):
    g = golden
    logits = O.forward(g["params"], g["cfg"], g["tokens"], g["cond"], autocast=torch.bfloat16).float()
    ref = g["logits_bf16"]
    rel = (logits - ref).norm() / ref.norm()
    assert rel < 2e-3, rel          # identical rounding points; typically exactly 0


def test_loss_and_grads_match_reference(golden):
    g = golden
    loss, _, grads = O.loss_and_grads(g["params"], g["cfg"], g["tokens"], g["cond"], g["target"])
    assert abs(float(loss) - g["loss_fp32"]) < 1e-5
    for k, ref in g["grads"].items():
        got = grads[k]
        denom = max(ref.abs().max().item(), 1e-8)
        assert (got - ref).abs().max().item() / denom < 2e-4, k


def test_decode_contract(golden):
    g = golden
    for t, ref in g["decode"].items():
        got = O.decode_last_logits(g["params"], g["cfg"], g["tokens"][:, :t], g["cond"])
        assert torch.allclose(got, ref, rtol=0, atol=2e-6), (t, (got - ref).abs().max())


def test_mask_known_answer(golden):
    g = golden
    toks = g["tokens"]
    if g["cfg"]["conditioning"] == "continuous_token":
        toks = torch.nn.functional.pad(toks, (2, 0), value=-1)
    assert torch.equal(O.key_mask(toks), g["mask"])


def test_positional_table_bit_exact(golden):
    g = golden
    d = g["cfg"]["d_model"]
    assert torch.equal(O.positional_table(d)[:256], g["pe_table"])


def test_positional_table_768_rows():
    z = np.load(os.path.join(GOLDEN_DIR, "pe_768_rows.npz"))
    tab = O.positional_table(768)
    for r, v in zip(z["rows"], z["values"]):
        assert np.array_equal(tab[int(r)].numpy(), v), r


def test_param_shapes_match_state_dict(golden):
    g = golden
    shapes = O.param_shapes(g["cfg"])
    assert set(shapes) == set(g["params"])
    for k, s in shapes.items():
        assert tuple(g["params"][k].shape) == s, k


def test_param_count_formula():
    # 754 287 parameters for 2L/128d/dc32/V1007 ... with E tables excluded
    cfg = dict(vocab_size=1007, n_layer=2, n_head=4, d_model=128, d_inner=512, d_condition=32,
               conditioning="continuous_concat", dropout=0.0)
    n = sum(int(np.prod(s)) for k, s in O.param_shapes(cfg).items())
    d, dc, V, di, NL, dh = 128, 32, 1007, 512, 2, 32
    expect = V * (d - dc) + (dc * 2 + dc) + NL * (2048 * dh + 4 * (d * d + d) + di * d + di + d * di + d + 4 * d) + V * d + V
    assert n == expect


def test_srel_closed_form_vs_explicit_loop():
    torch.manual_seed(0)
    B, H, L, dh = 1, 2, 9, 8
    q = torch.randn(B, H, L, dh)
    E = torch.randn(O.MAX_SEQ, dh)
    s = O.relative_logits(q, E)
    for i in range(L):
        for j in range(L):
            want = (q[0, 1, i] * E[O.MAX_SEQ - 1 - (i - j)]).sum() if j <= i else torch.tensor(0.0)
            assert abs(s[0, 1, i, j] - want) < 1e-5


def test_train_step_changes_params_and_is_finite(golden):
    g = golden
    state = {}
    loss, newp = O.train_step(g["params"], state, g["cfg"], g["tokens"], g["cond"], g["target"], lr=1e-3)
    assert np.isfinite(float(loss))
    moved = sum(float((newp[k] - g["params"][k]).abs().max()) > 0 for k in newp)
    assert moved > len(newp) // 2


# ---- regression side model (models/music_regression.py): oracle only, no CUDA path yet
def test_regression_oracle_matches_reference(regression_golden):
    g = regression_golden
    assert {k: tuple(v.shape) for k, v in g["params"].items()} == O.regression_param_shapes(g["cfg"])
    out = O.regression_forward(g["params"], g["cfg"], g["tokens"])
    torch.testing.assert_close(out, g["out_fp32"], rtol=1e-5, atol=1e-6)
    out16 = O.regression_forward(g["params"], g["cfg"], g["tokens"], autocast=torch.bfloat16)
    torch.testing.assert_close(out16.float(), g["out_bf16"], rtol=2e-2, atol=2e-2)


def test_regression_oracle_gradients_and_unmasked_pads(regression_golden):
    g = regression_golden
    leaves = {k: v.clone().requires_grad_(True) for k, v in g["params"].items()}
    out = O.regression_forward(leaves, g["cfg"], g["tokens"])
    loss = ((out - torch.tensor([[0.8, -0.8]])) ** 2).mean()
    assert float(loss.detach()) == pytest.approx(g["loss_fp32"], rel=1e-5)
    loss.backward()
    top = max(float(v.abs().max()) for v in g["grads"].values())
    for k, v in leaves.items():
        torch.testing.assert_close(v.grad, g["grads"][k], rtol=1e-4, atol=1e-6 * max(1.0, top), msg=lambda m: f"{k}: {m}")
    # no mask at all (music_regression.py:78): changing a pad position at the tail changes the pooled first position
    tok2 = g["tokens"].clone()
    tok2[0, -1] = 5
    assert not torch.equal(O.regression_forward(g["params"], g["cfg"], tok2)[0], out.detach()[0])
