"""ClipAdam (me_grad_sqnorm_partials / me_adam_prepare / me_adam_update through the C-ABI) against the arithmetic the
reference calls at train.py:319-325 -- torch.nn.utils.clip_grad_norm_ + torch.optim.Adam run on the same device --
and against the CPU oracle's restatement of it."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from midi_emotion_b200 import ClipAdam, build_model
    from oracle import midi_oracle as O

# chunk size of the kernels is 8192 elements, 64 tensors per launch: sizes on both sides of every boundary
SIZES = [1, 2, 3, 4, 5, 31, 768, 1007, 4096, 8191, 8192, 8193, 16384, 70001, 3 * 8192 + 2]


def _tensors(n_tensors, seed, offset_views=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    out = []
    for i in range(n_tensors):
        n = SIZES[i % len(SIZES)]
        if offset_views and i % 3 == 1:                         # 4-byte offset into an allocation: scalar path
            buf = torch.randn(n + 1, device="cuda", generator=g) * 0.1
            out.append(buf[1:].detach())
        else:
            out.append(torch.randn(n, device="cuda", generator=g) * 0.1)
    return out


def _set_grads(params, seed, scale):
    g = torch.Generator(device="cuda").manual_seed(seed)
    for p in params:
        p.grad = torch.randn(p.shape, device="cuda", generator=g) * scale


def _pair(n_tensors, seed, offset_views=False):
    base = _tensors(n_tensors, seed, offset_views)
    a = [torch.nn.Parameter(t) for t in base]                    # keeps the offset storage on our side
    b = [torch.nn.Parameter(t.detach().clone()) for t in base]
    return a, b


@pytest.mark.parametrize("n_tensors,clip,grad_mag,wd,offset_views", [
    (15, 1.0, 1.0, 0.0, False),       # clipping active (norm >> 1)
    (150, 1.0, 1e-5, 0.0, False),     # three launches per pass, clipping inactive
    (70, None, 0.3, 0.0, True),       # no norm pass at all, unaligned tensors
    (33, 0.5, 0.3, 0.01, True),       # L2 weight decay
])
def test_clip_adam_matches_torch_clip_grad_norm_and_adam(n_tensors, clip, grad_mag, wd, offset_views):
    ours, theirs = _pair(n_tensors, 1, offset_views)
    opt = ClipAdam(ours, lr=1e-3, weight_decay=wd, max_grad_norm=clip)
    ref = torch.optim.Adam(theirs, lr=1e-3, weight_decay=wd)
    for step in range(4):
        _set_grads(ours, 100 + step, grad_mag)
        _set_grads(theirs, 100 + step, grad_mag)
        kept = [p.grad.clone() for p in ours]
        norm = opt.step()
        want_norm = torch.nn.utils.clip_grad_norm_(theirs, clip) if clip is not None else None
        ref.step()
        torch.cuda.synchronize()
        if clip is not None:
            assert float(norm) == pytest.approx(float(want_norm), rel=1e-5)
            assert float(opt.last_stats[1]) == pytest.approx(min(1.0, clip / (float(want_norm) + 1e-6)), rel=1e-5)
        assert float(opt.last_stats[2]) == 0.0 and float(opt.last_stats[3]) == step + 1
        # tolerances: a few ulp of the operands (the two implementations contract multiply-adds differently); the
        # Adam step itself is ~lr = 1e-3 per element, five orders of magnitude above the parameter tolerance
        gm = grad_mag * (float(opt.last_stats[1]) if clip is not None else 1.0)
        for i, (p, q) in enumerate(zip(ours, theirs)):
            assert torch.equal(p.grad, kept[i]), "gradients are read-only"
            torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-5, atol=1e-8, msg=lambda m: f"step {step} tensor {i}: {m}")
            torch.testing.assert_close(opt.state[p]["exp_avg"], ref.state[q]["exp_avg"], rtol=1e-5, atol=1e-5 * gm)
            torch.testing.assert_close(opt.state[p]["exp_avg_sq"], ref.state[q]["exp_avg_sq"], rtol=1e-5, atol=1e-5 * gm * gm)
        opt.zero_grad(set_to_none=True)
        ref.zero_grad(set_to_none=True)


def test_clip_adam_matches_the_cpu_oracle():
    base = _tensors(20, 3)
    names = [f"t{i}" for i in range(len(base))]
    ours = [torch.nn.Parameter(t.clone()) for t in base]
    opt = ClipAdam(ours, lr=1e-3, max_grad_norm=1.0)               # train.py:321-322 clips at 1.0
    want, state = {n: t.cpu().clone() for n, t in zip(names, base)}, {}
    for step in range(3):
        _set_grads(ours, 200 + step, 0.002 * (step + 1) ** 2)      # norm below, near and above the clip threshold
        grads = {n: p.grad.cpu() for n, p in zip(names, ours)}
        norm = opt.step()
        want, want_norm, skipped = O.clip_adam_update(want, grads, state, lr=1e-3, clip=1.0)
        assert not skipped
        assert float(norm) == pytest.approx(float(want_norm), rel=1e-5)
        for n, p in zip(names, ours):
            torch.testing.assert_close(p.detach().cpu(), want[n], rtol=1e-5, atol=1e-8)


def test_unscale_and_skip_on_non_finite_gradients():
    ours, theirs = _pair(12, 5)
    opt = ClipAdam(ours, lr=1e-3, max_grad_norm=1.0)
    ref = ClipAdam(theirs, lr=1e-3, max_grad_norm=1.0)
    _set_grads(theirs, 7, 0.01)
    for p, q in zip(ours, theirs):
        p.grad = q.grad * 1024.0                                   # power-of-two scale: unscaling is exact
    n1 = opt.step(grad_scale=1024.0)
    n2 = ref.step()
    assert float(n1) == pytest.approx(float(n2), rel=1e-6)
    for p, q in zip(ours, theirs):
        assert torch.equal(p.detach(), q.detach())
    # overflow: nothing moves, the step count stays, the flag is raised
    snap = [p.detach().clone() for p in ours]
    m_snap = [opt.state[p]["exp_avg"].clone() for p in ours]
    ours[3].grad[0] = float("inf")
    opt.step(grad_scale=1024.0)
    assert float(opt.last_stats[2]) == 1.0 and float(opt.last_stats[3]) == 1.0
    assert float(opt.state[ours[0]]["step"]) == 1.0
    for p, s, m in zip(ours, snap, m_snap):
        assert torch.equal(p.detach(), s) and torch.equal(opt.state[p]["exp_avg"], m)
    ours[3].grad[0] = float("nan")
    opt.step()
    assert float(opt.last_stats[2]) == 1.0 and float(opt.state[ours[0]]["step"]) == 1.0
    # and the next finite step is step 2 on both sides
    _set_grads(ours, 8, 0.01)
    _set_grads(theirs, 8, 0.01)
    opt.step()
    ref.step()
    assert float(opt.last_stats[3]) == 2.0
    for p, q in zip(ours, theirs):
        assert torch.equal(p.detach(), q.detach())


def test_state_dict_moves_between_clip_adam_and_torch_adam():
    ours, theirs = _pair(9, 11)
    opt = ClipAdam(ours, lr=1e-3)
    for step in range(2):
        _set_grads(ours, 300 + step, 0.1)
        opt.step()
    sd = copy.deepcopy(opt.state_dict())
    assert all(float(st["step"]) == 2.0 and st["step"].device.type == "cpu" for st in sd["state"].values())
    with torch.no_grad():
        for p, q in zip(ours, theirs):
            q.copy_(p)
    ref = torch.optim.Adam(theirs, lr=1e-3)
    ref.load_state_dict(sd)
    again = ClipAdam([torch.nn.Parameter(p.detach().clone()) for p in ours], lr=1e-3)
    again.load_state_dict(copy.deepcopy(ref.state_dict()))     # (load_state_dict aliases the tensors it is given)
    mine = again.param_groups[0]["params"]
    for params in (ours, theirs, mine):
        _set_grads(params, 400, 0.1)
    opt.step()
    ref.step()
    again.step()
    assert float(opt.last_stats[3]) == 3.0 and float(again.last_stats[3]) == 3.0
    for p, q, r in zip(ours, theirs, mine):
        torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-5, atol=1e-8)
        assert torch.equal(p.detach(), r.detach())


def test_training_with_clip_adam_tracks_clip_grad_norm_plus_adam(golden):
    """The whole step of train.py:307-325 on the model: same losses over a few steps with either optimiser."""
    g = golden
    tokens, cond, target = g["tokens"].cuda(), g["cond"].cuda(), g["target"].cuda()
    losses = {}
    for kind in ("torch", "fused"):
        model, _ = build_model(dict(g["cfg"]))
        model.load_state_dict(g["params"])
        model = model.cuda().train()
        model.precision = "fp32"
        opt = (ClipAdam(model.parameters(), lr=1e-3, max_grad_norm=1.0) if kind == "fused"
               else torch.optim.Adam(model.parameters(), lr=1e-3))
        seq = []
        for _ in range(4):
            out = model(tokens, cond)
            loss = torch.nn.functional.cross_entropy(out.reshape(-1, out.size(-1)), target.reshape(-1), ignore_index=0)
            loss.backward()
            if kind == "torch":
                torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
            seq.append(loss.item())
        losses[kind] = seq
    assert losses["fused"][-1] < losses["fused"][0]
    for a, b in zip(losses["fused"], losses["torch"]):
        assert a == pytest.approx(b, rel=2e-4)


def test_clip_adam_rejects_what_it_cannot_do():
    p = torch.nn.Parameter(torch.zeros(8, device="cuda", dtype=torch.bfloat16))
    p.grad = torch.ones_like(p)
    with pytest.raises(RuntimeError, match="float32"):
        ClipAdam([p], lr=1e-3).step()
    q = torch.nn.Parameter(torch.zeros(8, device="cuda"))
    q.grad = torch.ones_like(q)
    with pytest.raises(RuntimeError, match="not implemented"):
        opt = ClipAdam([q], lr=1e-3)
        opt.param_groups[0]["amsgrad"] = True
        opt.step()
