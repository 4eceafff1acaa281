"""bench.py's host-side contract, checked without a GPU: the workload table against BASELINE.json / SURVEY.md 8(d),
the synthetic batch contract of each conditioning mode (loader.py:55-75,169-195 as restated in SURVEY.md A0) and the
JSON line of the reference arm."""
import json
import math
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_flops_match_the_survey_figures():
    # cfg2 199.7 MFLOP/token forward, 599.2 fwd+bwd, 19.63 TFLOP per 32 x 1024 step; cfg3 757.0 / 2271.1
    f2 = bench.flops_per_token(bench.CFG2, 1024)
    assert abs(f2 / 1e6 - 199.7) < 0.05
    assert abs(3 * f2 / 1e6 - 599.2) < 0.1
    assert abs(3 * f2 * 32 * 1024 / 1e12 - 19.63) < 0.01
    f3 = bench.flops_per_token(bench.CFG3, 2048)
    assert abs(f3 / 1e6 - 757.0) < 0.05 and abs(3 * f3 / 1e6 - 2271.1) < 0.1


def test_default_workload_is_baseline_configs_1():
    cfg, L, Ls, B, metric, label = bench.workload("cfg2")
    assert (cfg["n_layer"], cfg["d_model"], cfg["n_head"], cfg["d_inner"]) == (12, 768, 12, 3072)
    assert cfg["conditioning"] == "continuous_concat" and cfg["d_condition"] == 192 and cfg["vocab_size"] == 1007
    assert (L, Ls, B) == (1024, 1024, 32)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert metric in base["metric"] and "configs[1]" in label


def test_cfg3_and_sweep_workloads():
    cfg, L, Ls, B, metric, _ = bench.workload("cfg3")
    assert (cfg["n_layer"], cfg["d_model"], cfg["n_head"], L, Ls) == (24, 1024, 16, 2048, 2048)
    assert "seq2048" in metric
    for mode in ("discrete_token", "continuous_token", "none"):
        cfg, L, Ls, B, metric, label = bench.workload(mode)
        assert cfg["conditioning"] == mode and cfg["d_condition"] == -1 and Ls == 1024 and "configs[4]" in label
        assert cfg["vocab_size"] == (1017 if mode == "discrete_token" else 1007)
        assert L == (1022 if mode == "continuous_token" else 1024)


@pytest.mark.parametrize("mode", ["continuous_concat", "discrete_token", "continuous_token", "none"])
def test_synthetic_batch_contract(mode):
    name = "cfg2" if mode == "continuous_concat" else mode
    cfg, L, Ls, _, _, _ = bench.workload(name)
    tokens, cond, target = bench.synthetic_batch(cfg, 3, L, seed=5)
    assert tokens.shape == (3, L) and tokens.dtype == torch.int64 and cond.shape == (3, 2)
    assert target.shape == (3, Ls) and target.dtype == torch.int64
    assert int(tokens.min()) >= 1 and int(tokens.max()) < cfg["vocab_size"]
    shift = Ls - L
    assert torch.equal(target[:, shift:-1], tokens[:, 1:])            # next-token targets
    if mode == "continuous_token":
        assert int(target[:, :2].abs().sum()) == 0                     # the two condition positions predict nothing
    if mode == "discrete_token":
        assert bool(((tokens[:, 0] >= 1007) & (tokens[:, 0] <= 1016)).all())
        assert int(tokens[:, 1:].max()) < 1007
    else:
        assert bool((tokens[:, 0] == 1).all())                          # <START>
    if mode in ("none", "discrete_token"):
        assert bool(torch.isnan(cond).all())
    else:
        assert bool(((cond >= -1) & (cond <= 1)).all())
    again = bench.synthetic_batch(cfg, 3, L, seed=5)
    assert torch.equal(again[0], tokens) and torch.equal(again[2], target)


def test_reference_arm_prints_one_contract_line(monkeypatch, capsys):
    """`bench.py --impl reference` (oracle port on the host cores) on a shrunk model: one JSON line with the keys the
    driver reads."""
    small = dict(bench.CFG2, n_layer=1, d_model=64, n_head=2, d_inner=128, d_condition=16)
    monkeypatch.setattr(bench, "workload", lambda name: (small, 48, 48, 2, bench.METRIC, "shrunk"))

    class A:
        steps, warmup, gpus, workload = 1, 0, 1, "cfg2"
    bench.run_reference(A, rank=1)
    assert capsys.readouterr().out == ""                                # other ranks print nothing
    bench.run_reference(A, rank=0)
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "tokens/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and math.isfinite(j["value"]) and j["gpu_launches"] == 0
    # "reference" where /root/reference is on the machine (the unmodified model package is timed), else the port
    assert j["cpu_baseline"]["kind"] in ("port", "reference") and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["kind"] == ("reference" if os.path.isdir("/root/reference/src/models") else "port")
    assert j["e2e"] == {"value": j["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_b200_arm_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_decode_cpu_baseline_is_the_no_cache_recompute(monkeypatch):
    small = dict(bench.CFG2, n_layer=1, d_model=64, n_head=2, d_inner=128, d_condition=16)
    monkeypatch.setattr(bench, "CFG2", small)
    n = torch.get_num_threads()
    try:
        r = bench.cpu_decode_tokens_per_s(prefix=80, B=2, steps=1, threads=2)
    finally:
        torch.set_num_threads(n)
    assert r["kind"] == "port" and r["cores"] == 2 and r["unit"] == "tokens/s"
    assert r["value"] > 0 and math.isfinite(r["value"]) and "prefix 80" in r["sample"]
    assert r["value"] == pytest.approx(2 / (r["ms_per_step"] / 1e3))


def test_committed_bench_line_satisfies_the_driver_contract():
    """The last bench line measured on the B200 (profiles/r02_e_bench_final.json, written by `python bench.py`) carries
    every key the driver and the judge read, with consistent values."""
    lines = [l for l in open(os.path.join(ROOT, "profiles", "r02_e_bench_final.json")) if l.startswith("{")]
    assert len(lines) == 1                                              # ONE JSON line
    j = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in j, k
    assert j["metric"] == bench.METRIC and j["unit"] == "tokens/s" and j["dtype"] == "bf16" and j["vs_baseline"] is None
    assert j["warmup"] >= 3 and j["n_gpus"] == 1 and j["scaling"] == "weak" and j["data"] == "synthetic"
    assert "workload" in j["config"] and "model" not in j["config"]
    tokens = j["config"]["global_batch"] * j["config"]["seq_len"]
    assert j["value"] == pytest.approx(tokens / (j["ms_per_step"] / 1e3), rel=1e-6)
    e = j["e2e"]
    assert e["unit"] == "tokens/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != j["value"]                                     # a separately timed region
    r = j["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and 0 < r["frac"] < 1
    assert r["traffic"] is None or r["traffic"] > 0
    c = j["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert j["gpu_launches"] > 0
    # round-2 keys: live per-kernel roofline, the real generation, configs[2]
    assert {k["kernel"].split(" ")[0] for k in r["by_kernel"]} >= {"attn_fwd_tc_kernel", "attn_bwd_tc_kernel", "attn_bwd_q_tc_kernel"}
    assert 0 < r["whole_step_frac"] < 1 and r["largest_single_kernel"]["avg_launch_us"] > 0
    assert j["decode"]["roofline"]["bound"] == "hbm" and 0 < j["decode"]["roofline"]["frac"] < 1
    assert j["cfg3"]["value"] > 0
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert not bad & set(j["clocks"]["reasons"]) and j["clocks"]["sm_mhz"] > 0.8 * j["clocks"]["sm_max_mhz"]
