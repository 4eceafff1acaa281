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
