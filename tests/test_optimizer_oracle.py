"""The oracle's restatement of the optimiser step (train.py:319-325) pinned on the third-party arithmetic the
reference calls there -- torch.nn.utils.clip_grad_norm_ and torch.optim.Adam, executed here on the CPU -- and the
host-side contract of ClipAdam (state-dict layout interchangeable with torch.optim.Adam, no CPU fallback)."""
import copy

import pytest
import torch

from oracle import midi_oracle as O

SHAPES = {"a": (7,), "b": (33, 5), "c": (1,), "d": (129, 16), "e": (3, 3, 3)}


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(*s, generator=g) * 0.1 for k, s in SHAPES.items()}


def _grads(seed, scale):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(*s, generator=g) * scale for k, s in SHAPES.items()}


@pytest.mark.parametrize("clip,grad_mag,wd", [(1.0, 3.0, 0.0), (1.0, 1e-3, 0.0), (None, 0.5, 0.0), (0.25, 0.5, 0.01)])
def test_oracle_update_matches_torch_clip_and_adam(clip, grad_mag, wd):
    p0 = _params(1)
    torch_p = [torch.nn.Parameter(v.clone()) for v in p0.values()]
    opt = torch.optim.Adam(torch_p, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    mine, state = {k: v.clone() for k, v in p0.items()}, {}
    for step in range(4):
        grads = _grads(10 + step, grad_mag)
        for tp, g in zip(torch_p, grads.values()):
            tp.grad = g.clone()
        want_norm = None
        if clip is not None:
            want_norm = torch.nn.utils.clip_grad_norm_(torch_p, clip)
        opt.step()
        mine, norm, skipped = O.clip_adam_update(mine, grads, state, lr=1e-3, clip=clip, weight_decay=wd)
        assert not skipped and state["step"] == step + 1
        if want_norm is not None:
            assert float(norm) == pytest.approx(float(want_norm), rel=1e-6)
        for tp, (k, v) in zip(torch_p, mine.items()):
            torch.testing.assert_close(v, tp.detach(), rtol=2e-6, atol=1e-9, msg=lambda m: f"step {step} {k}: {m}")


def test_oracle_unscale_and_skip_on_overflow():
    p0, state = _params(2), {}
    grads = _grads(3, 1.0)
    scaled = {k: g * 1024.0 for k, g in grads.items()}
    a, na, _ = O.clip_adam_update(p0, grads, {}, lr=1e-3, clip=1.0)
    b, nb, _ = O.clip_adam_update(p0, scaled, state, lr=1e-3, clip=1.0, grad_scale=1024.0)
    assert float(na) == pytest.approx(float(nb), rel=1e-6)
    for k in a:
        torch.testing.assert_close(a[k], b[k], rtol=1e-6, atol=1e-9)
    bad = {k: g.clone() for k, g in scaled.items()}
    bad["b"][0, 0] = float("inf")
    c, nc, skipped = O.clip_adam_update(b, bad, state, lr=1e-3, clip=1.0, grad_scale=1024.0)
    assert skipped and state["step"] == 1 and not torch.isfinite(nc)
    for k in b:
        assert torch.equal(b[k], c[k])                       # GradScaler.step skips the update (train.py:323)


def test_clip_adam_state_dict_is_interchangeable_with_torch_adam():
    from midi_emotion_b200 import ClipAdam
    ps = [torch.nn.Parameter(v.clone()) for v in _params(4).values()]
    ref = torch.optim.Adam(ps, lr=3e-4, betas=(0.8, 0.99), eps=1e-7)
    for p, g in zip(ps, _grads(5, 1.0).values()):
        p.grad = g.clone()
    ref.step()
    ref.step()
    sd = copy.deepcopy(ref.state_dict())

    mine = ClipAdam([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=1.0, max_grad_norm=1.0)
    mine.load_state_dict(sd)
    back = mine.state_dict()
    assert set(back["param_groups"][0]) == set(sd["param_groups"][0])
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad"):
        assert back["param_groups"][0][k] == sd["param_groups"][0][k]
    assert set(back["state"]) == set(sd["state"])
    for i, st in sd["state"].items():
        assert set(back["state"][i]) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(back["state"][i]["step"]) == float(st["step"]) == 2.0
        assert back["state"][i]["step"].dtype == torch.float32 and back["state"][i]["step"].device.type == "cpu"
        assert torch.equal(back["state"][i]["exp_avg"], st["exp_avg"])
        assert torch.equal(back["state"][i]["exp_avg_sq"], st["exp_avg_sq"])
    # ... and the other way: torch.optim.Adam takes ClipAdam's state dict and keeps stepping
    other = torch.optim.Adam([torch.nn.Parameter(p.detach().clone()) for p in ps], lr=1.0)
    other.load_state_dict(back)
    for p, g in zip(other.param_groups[0]["params"], _grads(6, 1.0).values()):
        p.grad = g.clone()
    other.step()
    assert float(other.state[other.param_groups[0]["params"][0]]["step"]) == 3.0


def test_clip_adam_has_no_cpu_fallback():
    from midi_emotion_b200 import ClipAdam
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        ClipAdam([p], lr=1e-3).step()
    with pytest.raises(ValueError):
        ClipAdam([p], lr=-1.0)
