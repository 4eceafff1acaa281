"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
that include/midi_emotion_b200.h declares; the Python surface mirrors the reference's."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "midi_emotion_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(me_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    from midi_emotion_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert lib.me_version() >= 100
    assert isinstance(lib.me_last_error(), bytes)


def test_struct_mirrors_match_library():
    from midi_emotion_b200 import _lib
    lib = _lib.load()
    assert lib.me_sizeof_layer_args() == ctypes.sizeof(_lib.LayerArgs)
    assert lib.me_sizeof_layer_bwd_args() == ctypes.sizeof(_lib.LayerBwdArgs)
    assert lib.me_sizeof_attn_args() == ctypes.sizeof(_lib.AttnArgs)
    assert lib.me_sizeof_attn_bwd_args() == ctypes.sizeof(_lib.AttnBwdArgs)
    assert lib.me_sizeof_decode_layer_args() == ctypes.sizeof(_lib.DecodeLayerArgs)


def test_build_model_surface_matches_reference(golden):
    from midi_emotion_b200 import build_model
    g = golden
    model, args = build_model(dict(g["cfg"]))
    assert args["regression"] is False
    sd = model.state_dict()
    assert set(sd) == set(g["params"])
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(g["params"][k].shape), k
    model.load_state_dict(g["params"])
    for attr in ("max_seq", "num_layer", "embedding_dim", "vocab_size", "pad_token"):
        assert hasattr(model, attr)
    assert model.max_seq == 2048 and model.pad_token == 0


def test_build_model_load_config_dict_contract():
    from midi_emotion_b200 import build_model
    cfg = dict(vocab_size=67, n_layer=1, n_head=2, d_model=64, d_inner=128, dropout=0.3, d_condition=16,
               conditioning="continuous_concat", overwrite_dropout=True)
    model, args = build_model(None, load_config_dict=cfg)
    assert model.dropout_p == pytest.approx(0.3)
    with pytest.raises(KeyError):
        build_model(None, load_config_dict={k: v for k, v in cfg.items() if k != "overwrite_dropout"})
    with pytest.raises(AssertionError):       # models/music_regression.py:42: `assert d_condition <= 0`
        build_model(dict(cfg, regression=True))
    reg, _ = build_model(dict(cfg, regression=True, d_condition=-1, conditioning="none"))   # build_model.py:29-32
    assert type(reg).__name__ == "MusicRegression" and {"fc.0.weight", "fc.0.bias"} <= set(reg.state_dict())


def test_positional_table_matches_reference_rows():
    import numpy as np
    from conftest import GOLDEN_DIR
    from midi_emotion_b200 import positional_table
    z = np.load(os.path.join(GOLDEN_DIR, "pe_768_rows.npz"))
    tab = positional_table(768)
    for r, v in zip(z["rows"], z["values"]):
        assert np.array_equal(tab[int(r)].numpy(), v), r


def test_no_cpu_fallback_exists(golden):
    """The product refuses CPU tensors instead of computing on the host."""
    from midi_emotion_b200 import build_model
    g = golden
    model, _ = build_model(dict(g["cfg"]))
    with pytest.raises(RuntimeError, match="CUDA"):
        model(g["tokens"], g["cond"])


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "midi_emotion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f


def test_build_digest_does_not_depend_on_where_the_tree_lives(tmp_path):
    """The GPU box runs from a copy of the tree at another path: the library built here must be recognised as current
    there (otherwise every process -- every rank of a torchrun launch -- would rebuild it concurrently)."""
    import shutil
    import subprocess
    import sys
    from midi_emotion_b200 import build as b
    b.build()
    here = b._digest()
    assert open(os.path.join(b.LIBDIR, "build.sha256")).read().strip() == here
    pkg = tmp_path / "midi_emotion_b200"
    pkg.mkdir()
    shutil.copy(os.path.join(ROOT, "midi_emotion_b200", "build.py"), pkg / "build.py")
    (pkg / "__init__.py").write_text("")
    shutil.copytree(os.path.join(ROOT, "midi_emotion_b200", "csrc"), pkg / "csrc")
    shutil.copytree(os.path.join(ROOT, "include"), tmp_path / "include")
    out = subprocess.run([sys.executable, "-c", "from midi_emotion_b200 import build as b; print(b._digest())"],
                         cwd=tmp_path, capture_output=True, text=True, check=True).stdout.strip()
    assert out == here
