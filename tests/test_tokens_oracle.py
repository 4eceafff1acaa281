"""The oracle's token pipeline (oracle.token_pipeline_sample) against vectors produced by executing the reference's
own loader / data_processing lines (scripts/make_golden_tokens.py).  CPU only."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import midi_oracle as O

NAMES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("tokens_") and f.endswith(".npz"))


def load_tokens_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    t2i = {(int(e), int(v)): int(i) for e, v, i in g["map_tuples"]}
    for n, i in zip(g["map_symbol_names"], g["map_symbol_ids"]):
        t2i[str(n)] = int(i)
    g["maps"] = {"tuple2idx": t2i, "transposable_event_inds": [int(x) for x in g["transposable_event_inds"]]}
    g["conditioning"] = str(g["conditioning"])
    return g


def prefix_of(g, i):
    """Token ids the reference prepends (loader.py:143-149,156-160,164-170), in final order."""
    t2i = g["maps"]["tuple2idx"]
    pre = []
    if g["emotion_tokens"][i][0] >= 0:
        pre += [int(g["emotion_tokens"][i][0]), int(g["emotion_tokens"][i][1])]
    if int(g["regression"]) and int(g["use_cls_token"]):
        pre.append(t2i["<CLS>"])
    if int(g["start"][i]) < 0:
        pre.append(t2i["<START>"])
    return pre


@pytest.mark.parametrize("name", NAMES)
def test_token_pipeline_oracle_matches_reference_lines(name):
    g = load_tokens_golden(name)
    ctoken = g["conditioning"] == "continuous_token"
    input_len = int(g["tgt_len"]) - (2 if ctoken else 0)
    for i in range(int(g["B"])):
        ev = g["events"][i, :int(g["n_events"][i])]
        inp, tgt = O.token_pipeline_sample(
            ev, g["maps"]["tuple2idx"], g["maps"]["transposable_event_inds"], input_len,
            n_transpose=int(g["n_transpose"][i]), start=int(g["start"][i]), prefix=prefix_of(g, i),
            pad_id=g["maps"]["tuple2idx"]["<PAD>"], target_left_pad=2 if ctoken else 0,
            want_target=not int(g["regression"]))
        assert np.array_equal(inp, g["input"][i]), (name, i)
        if not int(g["regression"]):
            assert np.array_equal(tgt, g["target"][i]), (name, i)


def test_cases_cover_crop_and_bar_start():
    starts = np.concatenate([load_tokens_golden(n)["start"] for n in NAMES])
    assert (starts >= 0).any() and (starts < 0).any()
