"""Model-level parity on the B200: the CUDA path (through build_model / nn.Module / the C-ABI)
against the golden vectors recorded from the unmodified reference, and against the CPU oracle."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from gpu_util import record_parity, rel_err
    from midi_emotion_b200 import build_model
    from oracle import midi_oracle as O

# fp32: logits agree to accumulation-order noise and argmax token ids are identical (north_star).
FP32_ATOL = 3e-5
# bf16: north_star asks for 1e-3 relative against the reference's own bf16-autocast run.  Two bf16 runs cannot
# agree element-wise to 1e-3 (tests/test_gpu_parity_oracle.py, module docstring and profiles/parity_r02.json:
# 2.3e-3 for one layer, 4.4e-3 for twelve, with or without the reference's rounding sites inside attention), so
# each run is compared with the exact fp32 result: our error must not exceed the reference-autocast error by more
# than 15 % (+5e-4: the golden configurations are tiny), and our distance to the reference's bf16 logits must not exceed the reference's own distance to fp32
# by more than 10 % (measured at full size: 0.95-0.97 of it).
BF16_VS_REF_RATIO = 1.15


def _model(g, precision):
    model, _ = build_model(dict(g["cfg"]))
    model.load_state_dict(g["params"])
    model = model.cuda()
    model.precision = precision
    return model


def test_forward_fp32_matches_reference_golden(golden):
    g = golden
    model = _model(g, "fp32").eval()
    with torch.no_grad():
        logits = model(g["tokens"].cuda(), g["cond"].cuda()).cpu()
    ref = g["logits_fp32"]
    assert logits.shape == ref.shape and logits.dtype == torch.float32
    assert (logits - ref).abs().max().item() < FP32_ATOL * max(1.0, ref.abs().max().item())
    assert torch.equal(logits.argmax(-1), ref.argmax(-1))     # bit-exact token ids


def test_backward_fp32_matches_reference_golden(golden):
    g = golden
    model = _model(g, "fp32").train()     # dropout = 0 in the golden configs
    logits = model(g["tokens"].cuda(), g["cond"].cuda())
    loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.size(-1)), g["target"].cuda().reshape(-1),
                                             ignore_index=0)
    loss.backward()
    assert abs(loss.item() - g["loss_fp32"]) < 2e-5
    # rga.Wk.bias has an identically-zero gradient (softmax is invariant to a per-row constant), so
    # the reference holds only rounding noise there: scale errors by the largest gradient as well
    floor = 1e-4 * max(v.abs().max().item() for v in g["grads"].values())
    for name, p in model.named_parameters():
        ref = g["grads"][name]
        got = p.grad.cpu()
        denom = max(ref.abs().max().item(), floor)
        assert (got - ref).abs().max().item() / denom < 5e-4, name


def test_forward_bf16_close_to_reference_autocast(golden):
    g = golden
    model = _model(g, "auto").eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(g["tokens"].cuda(), g["cond"].cuda())
    assert logits.dtype == torch.bfloat16
    logits = logits.float().cpu()
    ref_bf16, ref_fp32 = g["logits_bf16"], g["logits_fp32"]
    ours = rel_err(logits, ref_fp32)
    theirs = rel_err(ref_bf16, ref_fp32)
    record_parity(f"golden_{g['cfg']['conditioning']}_{g['cfg']['d_model']}d_L{g['tokens'].shape[1]}",
                  {"bf16_vs_ref_fp32": ours, "ref_bf16_vs_ref_fp32": theirs, "bf16_vs_ref_bf16": rel_err(logits, ref_bf16)})
    assert ours <= BF16_VS_REF_RATIO * theirs + 5e-4, (ours, theirs)
    assert rel_err(logits, ref_bf16) <= BF16_VS_REF_RATIO * theirs + 5e-4, (rel_err(logits, ref_bf16), theirs)
    agree = (logits.argmax(-1) == ref_fp32.argmax(-1)).float().mean().item()
    ref_agree = (ref_bf16.argmax(-1) == ref_fp32.argmax(-1)).float().mean().item()
    assert agree >= ref_agree - 0.05


def test_backward_bf16_close_to_fp32_grads(golden):
    g = golden
    model = _model(g, "bf16").train()
    logits = model(g["tokens"].cuda(), g["cond"].cuda())
    loss = torch.nn.functional.cross_entropy(logits.float().reshape(-1, logits.size(-1)),
                                             g["target"].cuda().reshape(-1), ignore_index=0)
    loss.backward()
    assert abs(loss.item() - g["loss_fp32"]) < 3e-2
    worst = 0.0
    for name, p in model.named_parameters():
        ref = g["grads"][name]
        if ref.abs().max() < 1e-7:
            continue
        worst = max(worst, rel_err(p.grad.cpu(), ref))
        assert rel_err(p.grad.cpu(), ref) < 6e-2, (name, rel_err(p.grad.cpu(), ref))


def test_grad_accumulation_and_optimizer_step(golden):
    """train.py:307-325 caller contract: backward twice accumulates, clip + Adam update the
    parameters, and the next forward sees the new weights (packed device copies are refreshed)."""
    g = golden
    model = _model(g, "fp32").train()
    tokens, cond, target = g["tokens"].cuda(), g["cond"].cuda(), g["target"].cuda()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def loss_of():
        out = model(tokens, cond)
        return torch.nn.functional.cross_entropy(out.reshape(-1, out.size(-1)), target.reshape(-1), ignore_index=0)

    l0 = loss_of()
    l0.backward()
    g1 = [p.grad.clone() for p in model.parameters()]
    loss_of().backward()
    for a, p in zip(g1, model.parameters()):
        assert torch.allclose(p.grad, 2 * a, rtol=1e-4, atol=1e-7)
    opt.zero_grad()
    for _ in range(5):
        loss = loss_of()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        opt.zero_grad()
    assert loss_of().item() < l0.item()


def test_train_step_matches_oracle_train_step(golden):
    g = golden
    model = _model(g, "fp32").train()
    tokens, cond, target = g["tokens"].cuda(), g["cond"].cuda(), g["target"].cuda()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    out = model(tokens, cond)
    loss = torch.nn.functional.cross_entropy(out.reshape(-1, out.size(-1)), target.reshape(-1), ignore_index=0)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    _, want = O.train_step(g["params"], {}, g["cfg"], g["tokens"], g["cond"], g["target"], lr=1e-3)
    for name, p in model.named_parameters():
        # Adam's first step moves every weight by ~lr*sign(g): compare the update, not the weight
        upd_got = p.detach().cpu() - g["params"][name]
        upd_want = want[name] - g["params"][name]
        if name.endswith("rga.Wk.bias"):
            continue  # zero gradient up to rounding noise: Adam turns the noise's sign into a step
        big = g["grads"][name].abs() > 1e-4 * g["grads"][name].abs().max().clamp_min(1e-12)
        if big.any():
            assert torch.allclose(upd_got[big], upd_want[big], rtol=2e-2, atol=2e-5), name


def test_state_dict_round_trip(golden):
    g = golden
    model = _model(g, "fp32")
    sd = model.state_dict()
    assert set(sd) == set(g["params"])
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(g["params"][k].shape)
        assert torch.equal(v.cpu(), g["params"][k])


def test_cpu_input_is_an_error_not_a_fallback(golden):
    g = golden
    model = _model(g, "fp32")
    with pytest.raises(RuntimeError, match="CUDA"):
        model(g["tokens"], g["cond"])


def test_dropout_training_runs_and_is_seed_dependent(golden):
    g = golden
    model = _model(g, "fp32").train()
    model.dropout_p = 0.1
    tokens, cond = g["tokens"].cuda(), g["cond"].cuda()
    a = model(tokens, cond)
    b = model(tokens, cond)
    assert torch.isfinite(a).all() and not torch.equal(a, b)
    a.float().sum().backward()
    assert all(torch.isfinite(p.grad).all() for p in model.parameters())
    model.eval()
    with torch.no_grad():
        c = model(tokens, cond)
        d = model(tokens, cond)
    assert torch.equal(c, d)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_optimizer_step_reaches_the_next_forward(golden, precision):
    """torch's fused Adam updates parameters without bumping their version counters: the compute-type weight
    copies must be re-derived at every forward pass anyway (the reference re-casts under autocast per call)."""
    g = golden
    model = _model(g, precision).train()
    tok, cond, tgt = g["tokens"].cuda(), g["cond"].cuda(), g["target"].cuda()
    opt = torch.optim.Adam(model.parameters(), lr=1e-2, fused=True)
    before = model(tok, cond).detach().float()
    loss = torch.nn.functional.cross_entropy(model(tok, cond).float().reshape(-1, before.size(-1)), tgt.reshape(-1),
                                             ignore_index=0)
    loss.backward()
    opt.step()
    after = model(tok, cond).detach().float()
    assert not torch.equal(before, after), "the optimiser step did not reach the kernels"
    fresh = _model(dict(g, params=model.state_dict()), precision).train()
    assert torch.equal(fresh(tok, cond).detach().float(), after)
