"""Golden vectors for the token pipeline, produced by EXECUTING the reference's own code.

data/loader.py and data/data_processing.py do not import here (pretty_midi / pypianoroll are absent), so
  * `get_maps`, `transpose`, `tensor_to_tuples`, `tuples_to_ind_tensor`, `tensor_to_ind_tensor` are cut out of
    data/data_processing.py with `ast` and exec'd unmodified;
  * the body of `Loader.__getitem__` after the bar window has been chosen -- from "# transpose" to the left-pad
    of the target, data/loader.py:124-187 -- is cut out of the source text and exec'd unmodified against a stub
    `self`, with `random.choice`, `np.random.uniform` and `np.random.randint` replaced by recorded draws (the
    decisions are inputs of the device pipeline).
Output: tests/golden/tokens_*.npz.

    python scripts/make_golden_tokens.py        (needs /root/reference; run in the build container)
"""
import ast
import os
import textwrap
from copy import deepcopy
from types import SimpleNamespace

import numpy as np
import torch

DP = "/root/reference/src/data/data_processing.py"
LOADER = "/root/reference/src/data/loader.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def reference_functions():
    src = open(DP).read()
    tree = ast.parse(src)
    want = {"get_maps", "transpose", "tensor_to_tuples", "tuples_to_ind_tensor", "tensor_to_ind_tensor"}
    ns = {"torch": torch, "np": np, "deepcopy": deepcopy}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in want:
            exec(compile(ast.Module([node], []), DP, "exec"), ns)
    return ns


def reference_segment():
    lines = open(LOADER).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip() == "# transpose")
    end = next(i for i, l in enumerate(lines) if "target = torch.nn.functional.pad(target, (condition.size(0), 0)" in l)
    return textwrap.dedent("\n".join(lines[start:end + 1])), (start + 1, end + 1)


def run_sample(fns, code, maps, bars, cfg, draws, meta):
    """One pass of the cut-out __getitem__ body.  draws: n_transpose, r, start."""
    self = SimpleNamespace(
        transpose_options=cfg["transpose_options"], maps=maps, bar_start_prob=cfg["bar_start_prob"],
        input_len=cfg["input_len"], start_token="<START>", regression=cfg["regression"],
        use_cls_token=cfg["use_cls_token"], cls_token="<CLS>", conditioning=cfg["conditioning"],
        always_use_discrete_condition=cfg["always"], data=[meta], pad_token="<PAD>",
        get_pad_idx=lambda: maps["tuple2idx"]["<PAD>"])
    fake_random = SimpleNamespace(choice=lambda opts: draws["n_transpose"])
    fake_np = SimpleNamespace(random=SimpleNamespace(uniform=lambda: draws["r"],
                                                      randint=lambda lo, hi: draws["start"]), nan=np.nan)
    ns = {"self": self, "bars": bars.clone(), "idx": 0, "torch": torch, "np": fake_np, "random": fake_random,
          "transpose": fns["transpose"], "tensor_to_ind_tensor": fns["tensor_to_ind_tensor"]}
    exec(code, ns)
    return ns["input_"], ns.get("target"), ns["condition"], ns["start_at_beginning"]


def main():
    fns = reference_functions()
    code, span = reference_segment()
    print(f"executing data/loader.py:{span[0]}-{span[1]}")
    base_maps = fns["get_maps"]()
    rng = np.random.RandomState(7)
    cases = {
        "tokens_concat_L64": dict(conditioning="continuous_concat", regression=False, use_cls_token=False, tgt_len=64),
        "tokens_ctoken_L64": dict(conditioning="continuous_token", regression=False, use_cls_token=False, tgt_len=64),
        "tokens_discrete_L48": dict(conditioning="discrete_token", regression=False, use_cls_token=False, tgt_len=48),
        "tokens_regression_cls_L40": dict(conditioning="none", regression=True, use_cls_token=True, tgt_len=40),
    }
    for name, c in cases.items():
        maps = deepcopy(base_maps)
        extra = []
        if c["conditioning"] == "discrete_token":       # loader.py:58-66
            extra = sorted(["<V-2>", "<V-1>", "<V0>", "<V1>", "<V2>", "<A-2>", "<A-1>", "<A0>", "<A1>", "<A2>"])
        if c["regression"] and c["use_cls_token"]:      # loader.py:68-69
            extra.append("<CLS>")
        if extra:                                        # loader.py:71-76
            ml = list(maps["idx2tuple"].values()) + extra
            maps["idx2tuple"] = {i: v for i, v in enumerate(ml)}
            maps["tuple2idx"] = {v: i for i, v in enumerate(ml)}
        input_len = c["tgt_len"] - (2 if c["conditioning"] == "continuous_token" else 0)   # loader.py:55-57
        cfg = dict(transpose_options=list(range(-3, 4)), bar_start_prob=0.5, input_len=input_len,
                   regression=c["regression"], use_cls_token=c["use_cls_token"], conditioning=c["conditioning"],
                   always=False)
        B = 8
        rec = {"conditioning": c["conditioning"], "regression": int(c["regression"]),
               "use_cls_token": int(c["use_cls_token"]), "tgt_len": c["tgt_len"], "B": B,
               "span": np.array(span)}
        n_event_types = len(maps["event2idx"])
        events, inputs, targets, conds, starts, transposes, emo = [], [], [], [], [], [], []
        for i in range(B):
            n = int(rng.choice([5, input_len // 2, input_len, input_len + 1, 2 * input_len + 3, 3 * input_len]))
            ev = rng.randint(0, n_event_types, size=n)
            val = np.where(ev == n_event_types - 1, 8 * rng.randint(1, 126, size=n), rng.randint(21, 109, size=n))
            bars = torch.tensor(np.stack([ev, val], 1), dtype=torch.int16)
            n_tr = int(rng.randint(-3, 4))
            r = float(rng.uniform())
            start = int(rng.randint(0, max(1, n - input_len)))
            meta = {"valence": ["<V-2>", "<V1>"][i % 2], "arousal": ["<A2>", "<A-1>"][i % 2]} \
                if c["conditioning"] == "discrete_token" else {"valence": 0.25 * i - 1, "arousal": 0.5 - 0.1 * i}
            inp, tgt, cond, sab = run_sample(fns, code, maps, bars, cfg, dict(n_transpose=n_tr, r=r, start=start), meta)
            events.append(bars.numpy())
            inputs.append(inp.numpy())
            if tgt is not None:
                targets.append(tgt.numpy())
            conds.append(cond.numpy())
            starts.append(-1 if sab else start)
            transposes.append(n_tr)
            if c["conditioning"] == "discrete_token" and sab:
                emo.append([maps["tuple2idx"][meta["valence"]], maps["tuple2idx"][meta["arousal"]]])
            else:
                emo.append([-1, -1])
        nmax = max(len(e) for e in events)
        ev_pad = np.zeros((B, nmax, 2), dtype=np.int16)
        for i, e in enumerate(events):
            ev_pad[i, :len(e)] = e
        rec.update(events=ev_pad, n_events=np.array([len(e) for e in events], dtype=np.int32),
                   n_transpose=np.array(transposes, dtype=np.int32), start=np.array(starts, dtype=np.int32),
                   emotion_tokens=np.array(emo, dtype=np.int32), input=np.stack(inputs),
                   condition=np.stack(conds).astype(np.float32))
        if targets:
            rec["target"] = np.stack(targets)
        # the maps as arrays: tuple tokens (event, value, id), symbols (name, id), transposable events
        tup = [(k[0], k[1], v) for k, v in maps["tuple2idx"].items() if isinstance(k, tuple)]
        sym = [(k, v) for k, v in maps["tuple2idx"].items() if isinstance(k, str)]
        rec["map_tuples"] = np.array(tup, dtype=np.int32)
        rec["map_symbol_names"] = np.array([k for k, _ in sym])
        rec["map_symbol_ids"] = np.array([v for _, v in sym], dtype=np.int32)
        rec["transposable_event_inds"] = np.array(maps["transposable_event_inds"], dtype=np.int32)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, "starts", starts, "input shape", rec["input"].shape)


if __name__ == "__main__":
    main()
