"""The oracle's restatement of the sampling step (generate.py:122-189) against golden vectors produced by
executing the reference's own source lines (scripts/make_golden_sampling.py)."""
import torch

from oracle import midi_oracle as O


def test_oracle_sampling_matches_reference_lines(sampling_golden):
    g = sampling_golden
    tokens, probs, num_choices, new_counts = O.sample_step(
        g["logits"], g["prev"], g["repeat_counts"].tolist(), g["uniforms"], g["exclude"], g["is_timeshift"],
        temperatures=g["temperatures"], penalty_coeff=g["penalty_coeff"], top_k=g["top_k"], top_p=g["top_p"])
    # torch.topk leaves the order of exactly equal logits unspecified (e.g. the NaN -> 0 entries of row 0): a
    # draw that lands inside such a group may name any member of it
    for j in range(tokens.numel()):
        assert tokens[j] == g["tokens"][j] or g["logits"][j, tokens[j]] == g["logits"][j, g["tokens"][j]] or (
            torch.isnan(g["logits"][j, tokens[j]]) and g["logits"][j, g["tokens"][j]].nan_to_num(0.0) == 0), j
    assert torch.equal(num_choices.int(), g["num_choices"])
    assert new_counts == g["new_repeat_counts"].tolist()
    assert torch.allclose(probs, g["probs"], rtol=1e-6, atol=1e-9)
    assert int(g["ref_lines"][0]) == 122 and int(g["ref_lines"][1]) == 189


def test_oracle_sampling_never_picks_excluded_and_keeps_the_top_entry():
    torch.manual_seed(0)
    V, B = 300, 8
    logits = torch.randn(B, V) * 4
    exclude = torch.zeros(V, dtype=torch.uint8)
    exclude[:5] = 1
    logits[:, 2] = 50.0                       # the best logit is an excluded symbol
    ts = torch.zeros(V, dtype=torch.uint8)
    tokens, probs, n, _ = O.sample_step(logits, torch.zeros(B, dtype=torch.int64), [0] * B, torch.rand(B), exclude, ts,
                                        top_p=1e-6)
    assert (probs[:, :5] == 0).all() and (tokens >= 5).all()
    assert (n == 1).all()                      # top_p ~ 0 keeps exactly the first (best allowed) entry
    best = logits.clone()
    best[:, :5] = -float("inf")
    assert torch.equal(tokens, best.argmax(-1))


def test_oracle_cross_entropy_and_accuracy_match_the_reference(ce_golden):
    """oracle.ce_and_accuracy against nn.CrossEntropyLoss(ignore_index=0) (train.py:124) and the reference's own
    utils.accuracy (scripts/make_golden_ce.py)."""
    g = ce_golden
    loss, grad, hits, count = O.ce_and_accuracy(g["logits"], g["target"], ignore_index=0)
    assert abs(float(loss) - g["loss"]) < 1e-6 and count == g["count"]
    assert torch.allclose(grad, g["grad"], rtol=1e-6, atol=1e-9)
    assert hits[1] == g["top1"] and hits[5] == g["top5"]
