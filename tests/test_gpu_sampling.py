"""me_sample_step / Sampler / generate() on the B200 against the reference-line golden vectors and the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from midi_emotion_b200 import Sampler, build_model, generate
    from oracle import midi_oracle as O


def _run(g, logits_dtype=torch.float32):
    B, V = g["logits"].shape
    s = Sampler(B, V, exclude=g["exclude"], is_timeshift=g["is_timeshift"], temperatures=g["temperatures"],
                penalty_coeff=g["penalty_coeff"], top_k=g["top_k"], top_p=g["top_p"])
    s.repeat_counts.copy_(g["repeat_counts"].to(torch.int32))
    probs = torch.full((B, V), -1.0, device="cuda")
    tok = s.sample(g["logits"].cuda().to(logits_dtype), g["prev"].cuda(), uniforms=g["uniforms"].cuda(), out_probs=probs)
    return tok.cpu(), probs.cpu(), s.num_choices.cpu(), s.repeat_counts.cpu()


def test_sampling_kernel_matches_reference_lines(sampling_golden):
    g = sampling_golden
    tok, probs, n, rc = _run(g)
    # probabilities to fp32 rounding (expf/logf vs torch's); the kept set may differ only where the cumulative
    # sum sits within rounding of top_p, which the golden cases avoid
    assert torch.allclose(probs, g["probs"], rtol=2e-5, atol=1e-8)
    assert torch.equal(n, g["num_choices"])
    assert torch.equal(rc, g["new_repeat_counts"])
    lz = g["logits"].nan_to_num(0.0)
    for j in range(tok.numel()):   # equal logits form a group whose internal order torch.topk leaves unspecified
        assert tok[j] == g["tokens"][j] or lz[j, tok[j]] == lz[j, g["tokens"][j]], j


def test_sampling_kernel_bf16_logits_match_the_oracle_on_the_same_rounded_logits(sampling_golden):
    g = dict(sampling_golden)
    g["logits"] = g["logits"].to(torch.bfloat16).float()          # what the bf16 decode path hands over
    want_tok, want_probs, want_n, want_rc = O.sample_step(
        g["logits"], g["prev"], g["repeat_counts"].tolist(), g["uniforms"], g["exclude"], g["is_timeshift"],
        temperatures=g["temperatures"], penalty_coeff=g["penalty_coeff"], top_k=g["top_k"], top_p=g["top_p"])
    tok, probs, n, rc = _run(g, torch.bfloat16)
    assert torch.allclose(probs, want_probs, rtol=2e-5, atol=1e-8)
    # bf16 rounding creates ties; both sides break them by the lower token id
    assert torch.equal(tok, want_tok) and torch.equal(n, want_n.int()) and rc.tolist() == want_rc


def test_sampling_edge_cases():
    V, B = 1007, 4
    logits = torch.zeros(B, V)
    logits[0] = float("nan")                                   # all NaN -> uniform over the allowed symbols
    logits[1, 10] = 80.0                                       # one dominant entry
    logits[2, :] = -1e30
    logits[2, 500] = 0.0
    logits[3] = torch.linspace(-5, 5, V)
    exclude = torch.zeros(V, dtype=torch.uint8)
    exclude[:5] = 1
    s = Sampler(B, V, exclude=exclude, top_p=0.7, penalty_coeff=0.5)
    probs = torch.empty(B, V, device="cuda")
    u = torch.tensor([0.999999, 0.0, 0.5, 0.25], device="cuda")
    tok = s.sample(logits.cuda(), torch.zeros(B, dtype=torch.int64, device="cuda"), uniforms=u, out_probs=probs).cpu()
    probs = probs.cpu()
    assert torch.isfinite(probs).all() and torch.allclose(probs.sum(-1), torch.ones(B), atol=1e-5)
    assert (probs[:, :5] == 0).all() and (tok >= 5).all()
    assert tok[1] == 10 and tok[2] == 500 and s.num_choices[1] == 1 and s.num_choices[2] == 1
    assert s.repeat_counts.tolist()[1:3] == [1, 1]             # <= 2 choices -> count + 1 (generate.py:187)


def test_generate_loop_runs_on_device_and_is_reproducible():
    cfg = dict(vocab_size=1007, n_layer=2, n_head=4, d_model=256, d_inner=512, dropout=0.0, d_condition=64,
               conditioning="continuous_concat")
    torch.manual_seed(3)
    model, _ = build_model(cfg)
    model = model.cuda().eval()
    B, t0, n = 8, 5, 40
    primer = torch.randint(5, 1007, (B, t0), device="cuda")
    primer[:, 0] = 1
    cond = torch.tensor([[0.8, 0.8], [-0.8, 0.8], [0.8, -0.8], [-0.8, -0.8]] * 2, device="cuda")
    exclude = torch.zeros(1007, dtype=torch.uint8)
    exclude[:5] = 1
    outs = []
    for _ in range(2):
        s = Sampler(B, 1007, exclude=exclude, seed=11)
        outs.append(generate(model, primer, cond, n, s, max_len=64))
    assert outs[0].shape == (B, t0 + n) and torch.equal(outs[0], outs[1])
    assert torch.equal(outs[0][:, :t0], primer) and (outs[0][:, t0:] >= 5).all()
    s = Sampler(B, 1007, exclude=exclude, seed=12)
    assert not torch.equal(generate(model, primer, cond, n, s, max_len=64), outs[0])
