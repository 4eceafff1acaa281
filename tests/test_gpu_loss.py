"""Fused cross-entropy (me_cross_entropy / midi_emotion_b200.cross_entropy) on the B200 against golden vectors made
with nn.CrossEntropyLoss and the reference's utils.accuracy, and inside the training step."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from midi_emotion_b200 import build_model, cross_entropy


def test_fused_cross_entropy_matches_reference(ce_golden):
    g = ce_golden
    x = g["logits"].cuda().requires_grad_(True)
    loss, stats = cross_entropy(x, g["target"].cuda(), ignore_index=0, return_stats=True)
    loss.backward()
    assert abs(loss.item() - g["loss"]) < 2e-5 * max(1.0, abs(g["loss"]))
    assert int(stats["count"].item()) == g["count"]
    assert int(stats["top1"].item()) == g["top1"] and int(stats["top5"].item()) == g["top5"]
    assert torch.allclose(x.grad.cpu(), g["grad"], rtol=1e-4, atol=1e-8)
    assert (x.grad.cpu()[g["target"] == 0] == 0).all()   # ignored rows get no gradient


def test_fused_cross_entropy_bf16_logits(ce_golden):
    """bf16 logits (what the bf16 head hands over): fp32 arithmetic on the rounded values, as CrossEntropyLoss does
    under autocast; the gradient is rounded to bf16 on the way out."""
    g = ce_golden
    t = g["target"].cuda()
    x = g["logits"].cuda().to(torch.bfloat16).requires_grad_(True)
    loss, stats = cross_entropy(x, t, ignore_index=0, return_stats=True)
    loss.backward()
    ref = x.detach().float().requires_grad_(True)
    want = torch.nn.functional.cross_entropy(ref, t, ignore_index=0)
    want.backward()
    assert abs(loss.item() - want.item()) < 2e-5 * max(1.0, abs(want.item()))
    assert torch.allclose(x.grad.float(), ref.grad, rtol=1e-2, atol=1e-7)
    valid = t != 0
    rank = (ref.detach()[valid] > ref.detach()[valid].gather(1, t[valid, None])).sum(-1)
    assert int(stats["top1"].item()) == int((rank < 1).sum()) and int(stats["top5"].item()) == int((rank < 5).sum())


def test_fused_cross_entropy_on_padded_pitch_and_upstream_scale():
    torch.manual_seed(0)
    M, V, Vp = 64, 1007, 1008
    base = torch.randn(M, Vp, device="cuda").to(torch.bfloat16)
    x = base[:, :V].view(4, 16, V).detach().requires_grad_(True)    # what the model returns: a slice of padded rows
    t = torch.randint(0, V, (4, 16), device="cuda")
    (3.0 * cross_entropy(x, t, ignore_index=0)).backward()
    ref = x.detach().float().reshape(-1, V).clone().requires_grad_(True)
    (3.0 * torch.nn.functional.cross_entropy(ref, t.reshape(-1), ignore_index=0)).backward()
    assert torch.allclose(x.grad.float().reshape(-1, V), ref.grad, rtol=2e-2, atol=1e-6)


def test_training_step_with_fused_loss_matches_pytorch_loss(golden):
    g = golden
    grads = {}
    for fused in (False, True):
        model, _ = build_model(dict(g["cfg"]))
        model.load_state_dict(g["params"])
        model = model.cuda().train()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = model(g["tokens"].cuda(), g["cond"].cuda())
        tgt = g["target"].cuda()
        if fused:
            loss = cross_entropy(logits, tgt, ignore_index=0)
        else:
            loss = torch.nn.functional.cross_entropy(logits.reshape(-1, logits.size(-1)).float(), tgt.reshape(-1),
                                                     ignore_index=0)
        loss.backward()
        grads[fused] = (loss.item(), {n: p.grad.clone() for n, p in model.named_parameters()})
    assert abs(grads[True][0] - grads[False][0]) < 1e-4 * max(1.0, abs(grads[False][0]))
    top = max(v.abs().max().item() for v in grads[False][1].values())
    for n, ref in grads[False][1].items():
        got = grads[True][1][n]
        assert (got - ref).abs().max().item() <= 2e-2 * max(ref.abs().max().item(), 1e-3 * top), n


def test_fused_head_cross_entropy_kernel_against_ce_goldens(ce_golden):
    """me_head_cross_entropy with an identity head (logits = x I^T + 0): loss, gradient and top-k counts of the
    golden logits (rounded to bf16, as the head hands them over) without a logits tensor."""
    from midi_emotion_b200 import _lib
    from midi_emotion_b200._lib import ptr
    g = ce_golden
    t = g["target"].cuda().reshape(-1)
    x32 = g["logits"].cuda().reshape(t.numel(), -1)
    M, V = x32.shape
    Kp = (V + 7) // 8 * 8
    x = torch.zeros(M, Kp, device="cuda", dtype=torch.bfloat16)
    x[:, :V] = x32.to(torch.bfloat16)
    W = torch.zeros(V, Kp, device="cuda", dtype=torch.bfloat16)
    W[:, :V] = torch.eye(V, device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(V, device="cuda")
    grad = torch.full((M, Kp), float("nan"), device="cuda", dtype=torch.bfloat16)
    stats = torch.empty(4, device="cuda")
    _lib.call("me_head_cross_entropy", ptr(x), ptr(W), ptr(bias), M, V, Kp, Kp, Kp, ptr(t), 0, ptr(grad), Kp, ptr(stats),
              torch.cuda.current_stream().cuda_stream)
    ref = x[:, :V].float().requires_grad_(True)
    want = torch.nn.functional.cross_entropy(ref, t, ignore_index=0)
    want.backward()
    s = stats.cpu()
    assert int(s[1]) == int((t != 0).sum())
    assert abs(float(s[0] / s[1]) - want.item()) < 2e-5 * max(1.0, abs(want.item()))
    assert torch.isfinite(grad.float()).all() and (grad[:, V:] == 0).all()
    assert torch.allclose(grad[:, :V].float(), ref.grad, rtol=1e-2, atol=1e-7)
    valid = t != 0
    rank = (ref.detach()[valid] > ref.detach()[valid].gather(1, t[valid, None])).sum(-1)
    assert int(s[2]) == int((rank < 1).sum()) and int(s[3]) == int((rank < 5).sum())


def test_model_loss_fused_head_matches_unfused_composition(golden):
    """model.loss(x, cond, target) == cross_entropy(model(x, cond), target): value, top-k counts and every gradient."""
    g = golden
    res = {}
    for fused in (False, True):
        model, _ = build_model(dict(g["cfg"]))
        model.load_state_dict(g["params"])
        model = model.cuda().train()
        tok, cond, tgt = g["tokens"].cuda(), g["cond"].cuda(), g["target"].cuda()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if fused:
                loss, stats = model.loss(tok, cond, tgt, ignore_index=0, return_stats=True)
            else:
                loss, stats = cross_entropy(model(tok, cond), tgt, ignore_index=0, return_stats=True)
        (2.0 * loss).backward()
        res[fused] = (loss.item(), {k: int(v.item()) for k, v in stats.items()},
                      {n: p.grad.clone() for n, p in model.named_parameters()})
    assert abs(res[True][0] - res[False][0]) < 1e-4 * max(1.0, abs(res[False][0]))
    assert res[True][1] == res[False][1]
    top = max(v.abs().max().item() for v in res[False][2].values())
    for n, ref in res[False][2].items():
        got = res[True][2][n]
        assert (got - ref).abs().max().item() <= 2e-2 * max(ref.abs().max().item(), 1e-3 * top), n


def test_model_loss_fp32_path_is_the_composition(golden):
    g = golden
    model, _ = build_model(dict(g["cfg"]))
    model.load_state_dict(g["params"])
    model = model.cuda().train()
    model.precision = "fp32"
    loss = model.loss(g["tokens"].cuda(), g["cond"].cuda(), g["target"].cuda())
    assert abs(loss.item() - g["loss_fp32"]) < 2e-5
