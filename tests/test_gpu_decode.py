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


def test_decoder_reuse_with_a_new_condition_after_graph_capture(golden):
    """the captured step must not keep the first prompt's condition (or weight copies) by address."""
    g = golden
    if g["cfg"]["conditioning"] not in ("continuous_concat", "continuous_token"):
        pytest.skip("condition is not read in this mode")
    model = _model(g)
    tokens = g["tokens"].cuda()
    B = tokens.shape[0]
    cond_a = g["cond"].cuda()
    cond_b = (-cond_a).clone()
    dec = KVCacheDecoder(model, B, max_len=64, precision="fp32", use_cuda_graph=True)
    dec.prefill(tokens[:, :2], cond_a)
    for t in range(2, 6):
        dec.step(tokens[:, t])                      # eager, capture, replays
    assert dec.graph is not None
    got = dec.prefill(tokens[:, :2], cond_b.clone())   # a temporary: its storage may be gone at replay time
    got = [got.clone()] + [dec.step(tokens[:, t]).clone() for t in range(2, 6)]
    fresh = KVCacheDecoder(model, B, max_len=64, precision="fp32", use_cuda_graph=False)
    want = [fresh.prefill(tokens[:, :2], cond_b).clone()] + [fresh.step(tokens[:, t]).clone() for t in range(2, 6)]
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    # weights re-allocated (load into new storage): the graph is dropped, results follow the new weights
    with torch.no_grad():
        for p in model.parameters():
            p.data = p.data.clone() * 1.01
    after = dec.prefill(tokens[:, :2], cond_b)
    after = dec.step(tokens[:, 2]).clone()
    fresh2 = KVCacheDecoder(model, B, max_len=64, precision="fp32", use_cuda_graph=False)
    fresh2.prefill(tokens[:, :2], cond_b)
    assert torch.equal(after, fresh2.step(tokens[:, 2]))


@pytest.mark.parametrize("varying", [False, True])
def test_generate_slides_the_window_like_the_reference(golden, varying):
    """generate.py:99-122: past max_input_len the model sees the last max_input_len tokens only (positions restart
    at 0 every step).  generate() must equal that full-window recompute token for token."""
    from midi_emotion_b200 import Sampler, generate
    g = golden
    cfg = g["cfg"]
    model = _model(g)
    model.precision = "fp32"
    tokens, cond = g["tokens"].cuda(), g["cond"].cuda()
    B, V = tokens.shape[0], cfg["vocab_size"]
    t0, gen_len, max_input_len = 3, 14, 10
    extra = 2 if cfg["conditioning"] == "continuous_token" else 0
    window = max_input_len - extra
    vc = None
    if varying:
        if cfg["conditioning"] not in ("continuous_concat", "continuous_token"):
            pytest.skip("condition is not read in this mode")
        gg = torch.Generator().manual_seed(3)
        vc = (torch.rand(B, gen_len, generator=gg).cuda() * 2 - 1, torch.rand(B, gen_len, generator=gg).cuda() * 2 - 1)
    excl = torch.zeros(V, dtype=torch.uint8)
    excl[:2] = 1
    got = generate(model, tokens[:, :t0], cond, gen_len, Sampler(B, V, exclude=excl, seed=11), precision="fp32",
                   max_input_len=max_input_len, varying_condition=vc)
    # the reference's loop, literally: whole window through the model, last position, sample
    ref_sampler = Sampler(B, V, exclude=excl, seed=11)
    song = tokens[:, :t0].clone()
    for i in range(gen_len):
        c = cond if vc is None else torch.stack([vc[0][:, i], vc[1][:, i]], -1)
        with torch.no_grad():
            logits = model(song[:, -window:].contiguous(), c)[:, -1, :].contiguous()
        nxt = ref_sampler.sample(logits, song[:, -1].contiguous()).clone()
        song = torch.cat([song, nxt[:, None]], 1)
    assert torch.equal(got, song)
