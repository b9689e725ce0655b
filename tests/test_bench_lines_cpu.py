"""The bench lines kept under profiles/ (stdout of `bench.py` on a B200, the code as committed at the end of the
round) carry every key the measurement contract names: a change to bench.py that drops one shows up here."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FINAL = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_*_final.json")))

BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"]


def _line(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_final_lines_exist():
    names = {os.path.basename(p) for p in FINAL}
    for n in (1, 2, 4, 8):
        assert "r02_bench_n%d_cfg2_final.json" % n in names


@pytest.mark.parametrize("path", FINAL, ids=[os.path.basename(p) for p in FINAL])
def test_line_follows_the_contract(path):
    j = _line(path)
    for k in BASE_KEYS:
        assert k in j, k
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert baseline["metric"].startswith(j["metric"])          # BASELINE's metric names both parts; `value` is the first
    assert j["unit"] == "samples/s" and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["dtype"] == "f32" and j["data"] == "synthetic" and j["scaling"] in ("weak", "strong")
    n = int(os.path.basename(path).split("_")[2][1:])
    assert j["n_gpus"] == n and j["steps"] >= 1 and j["warmup"] >= 0
    assert "workload" in j["config"] and "model" not in j["config"]
    # value = samples of all ranks / step time
    B = j["config"]["global_batch"]
    assert abs(j["value"] - B / (j["ms_per_step"] * 1e-3)) <= 1e-6 * j["value"]
    e = j["e2e"]
    assert e["unit"] == j["unit"] and e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != j["value"]
    c = j["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"]
    assert not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert j["gpu_launches"] > 0
    r = j["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    if n == 1 and "cfg2" in path:
        cb = j["cpu_baseline"]
        assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
        assert r["traffic"] and r["traffic"] >= 0.9 * r["algorithmic_bytes_per_launch"]
    if n > 1:
        assert {"dir_shard_gather_send", "dir_embed_bwd_reduce_emit_to", "dir_shard_owner_update"} <= set(
            j["sharded_calls_us"])
        assert 0 < j["nvlink"]["frac"] <= 1.0
