"""Device token pipeline (me_token_pipeline through TokenPipeline) against the vectors produced by the reference's
own loader lines, and against the oracle on larger random batches.  Bit-exact (integer work)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from test_tokens_oracle import NAMES, load_tokens_golden  # noqa: E402

if torch.cuda.is_available():
    from midi_emotion_b200.tokens import TokenPipeline
    from oracle import midi_oracle as O


def _pipe(g, tgt_len=None):
    return TokenPipeline(g["maps"], input_len=int(tgt_len or g["tgt_len"]), conditioning=g["conditioning"],
                         regression=bool(int(g["regression"])), use_cls_token=bool(int(g["use_cls_token"])))


@pytest.mark.parametrize("name", NAMES)
def test_token_pipeline_matches_reference_lines(name):
    g = load_tokens_golden(name)
    pipe = _pipe(g)
    B = int(g["B"])
    events = [torch.from_numpy(g["events"][i, :int(g["n_events"][i])]) for i in range(B)]
    emo = [tuple(int(x) for x in g["emotion_tokens"][i]) if g["emotion_tokens"][i][0] >= 0 else None for i in range(B)]
    inp, tgt = pipe(events, n_transpose=g["n_transpose"].tolist(), start=g["start"].tolist(), emotion_tokens=emo)
    assert torch.equal(inp.cpu(), torch.from_numpy(g["input"]))
    if int(g["regression"]):
        assert tgt is None
    else:
        assert torch.equal(tgt.cpu(), torch.from_numpy(g["target"]))
    assert int(pipe.last_status.sum()) == 0


@pytest.mark.parametrize("conditioning,tgt_len,B", [("continuous_concat", 1024, 32), ("continuous_token", 1024, 32),
                                                    ("continuous_concat", 2048, 16)])
def test_token_pipeline_matches_oracle_at_baseline_sizes(conditioning, tgt_len, B):
    g = load_tokens_golden("tokens_concat_L64")
    g["conditioning"] = conditioning
    pipe = _pipe(g, tgt_len)
    rng = np.random.RandomState(tgt_len + B)
    n_types = int(g["map_tuples"][:, 0].max()) + 1
    events, starts, trs = [], [], []
    for i in range(B):
        n = int(rng.choice([0, 7, pipe.input_len, pipe.input_len + 1, 3 * pipe.input_len])) if i < 6 \
            else int(rng.randint(1, 4 * pipe.input_len))
        ev = rng.randint(0, n_types, size=n)
        val = np.where(ev == n_types - 1, 8 * rng.randint(1, 126, size=n), rng.randint(21, 109, size=n))
        events.append(torch.tensor(np.stack([ev, val], 1).reshape(-1, 2), dtype=torch.int16))
        crop = n > pipe.input_len and rng.rand() > 0.5
        starts.append(int(rng.randint(0, n - pipe.input_len)) if crop else -1)
        trs.append(int(rng.randint(-3, 4)))
    inp, tgt = pipe(events, n_transpose=trs, start=starts)
    t2i = g["maps"]["tuple2idx"]
    for i in range(B):
        pre = [t2i["<START>"]] if starts[i] < 0 else []
        want_in, want_tg = O.token_pipeline_sample(events[i].numpy(), t2i, g["maps"]["transposable_event_inds"],
                                                   pipe.input_len, trs[i], starts[i], pre, t2i["<PAD>"],
                                                   pipe.target_left_pad)
        assert np.array_equal(inp[i].cpu().numpy(), want_in), i
        assert np.array_equal(tgt[i].cpu().numpy(), want_tg), i
    assert int(pipe.last_status.sum()) == 0


def test_unmapped_tuple_is_flagged():
    g = load_tokens_golden("tokens_concat_L64")
    pipe = _pipe(g)
    ev = torch.tensor([[0, 21], [0, 5]], dtype=torch.int16)     # pitch 5 has no token
    pipe([ev, ev[:1]])
    assert pipe.last_status.cpu().tolist() == [1, 0]
