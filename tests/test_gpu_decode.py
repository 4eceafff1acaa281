"""KV-cache decode (new capability) against the reference's full-prefix recompute
(generate.py:99-122): golden last-position logits recorded from the unmodified reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from gpu_util import rel_err
    from midi_emotion_b200 import KVCacheDecoder, build_model


def _model(g):
    model, _ = build_model(dict(g["cfg"]))
    model.load_state_dict(g["params"])
    return model.cuda().eval()


@pytest.mark.parametrize("use_graph", [False, True])
def test_decode_fp32_matches_reference_recompute(golden, use_graph):
    g = golden
    model = _model(g)
    tokens, cond = g["tokens"].cuda(), g["cond"].cuda()
    B, L = tokens.shape
    dec = KVCacheDecoder(model, B, max_len=256, precision="fp32", use_cuda_graph=use_graph)
    checked = 0
    logits = dec.prefill(tokens[:, :1], cond)
    for t in range(1, max(g["decode"]) + 1):
        if t in g["decode"]:
            ref = g["decode"][t]
            got = logits.float().cpu()
            assert (got - ref).abs().max().item() < 3e-5 * max(1.0, ref.abs().max().item()), t
            assert torch.equal(got.argmax(-1), ref.argmax(-1)), t
            checked += 1
        if t < L:
            logits = dec.step(tokens[:, t])
    assert checked == len(g["decode"])


def test_decode_prefill_then_steps_equals_full_forward(golden):
    g = golden
    model = _model(g)
    model.precision = "fp32"
    tokens, cond = g["tokens"].cuda(), g["cond"].cuda()
    B, L = tokens.shape
    t0 = L // 2
    dec = KVCacheDecoder(model, B, max_len=512, precision="fp32")
    logits = dec.prefill(tokens[:, :t0], cond)
    with torch.no_grad():
        full = model(tokens, cond)            # [B, Ls, V]
    off = 2 if g["cfg"]["conditioning"] == "continuous_token" else 0
    assert torch.allclose(logits, full[:, off + t0 - 1], rtol=0, atol=3e-5)
    for t in range(t0, L):
        logits = dec.step(tokens[:, t])
        assert torch.allclose(logits, full[:, off + t], rtol=0, atol=3e-5), t
        assert torch.equal(logits.argmax(-1), full[:, off + t].argmax(-1))


def test_decode_bf16_tracks_fp32(golden):
    g = golden
    model = _model(g)
    tokens, cond = g["tokens"].cuda(), g["cond"].cuda()
    B, L = tokens.shape
    d32 = KVCacheDecoder(model, B, max_len=256, precision="fp32")
    d16 = KVCacheDecoder(model, B, max_len=256, precision="bf16")
    a = d32.prefill(tokens[:, :3], cond)
    b = d16.prefill(tokens[:, :3], cond)
    assert rel_err(b.float(), a) < 3e-2
    for t in range(3, min(L, 40)):
        a = d32.step(tokens[:, t])
        b = d16.step(tokens[:, t])
        assert rel_err(b.float(), a) < 3e-2, t


def test_decode_refuses_to_slide_the_window(golden):
    g = golden
    model = _model(g)
    tokens, cond = g["tokens"].cuda(), g["cond"].cuda()
    B = tokens.shape[0]
    dec = KVCacheDecoder(model, B, max_len=8, precision="fp32", use_cuda_graph=False)
    dec.prefill(tokens[:, :4], cond)
    with pytest.raises(RuntimeError, match="cache is full"):
        for t in range(4, 12):
            dec.step(tokens[:, t])
    with pytest.raises(ValueError):
        KVCacheDecoder(model, B, max_len=4096)
