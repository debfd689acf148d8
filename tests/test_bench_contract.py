"""bench.py's JSON-line contract (CPU): the reference arm is run here for real (it is a CPU program: the
unmodified reference class from oracle/_ref, or the oracle port when the reference is not built); the GPU arm's
line is checked on the recorded run in profiles/ (the GPU tests cannot run here)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def last_json_line(text: str) -> dict:
    lines = [l for l in text.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, "bench.py must print exactly one JSON line"
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = last_json_line(r.stdout)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["value"] > 0 and d["unit"] == "song-pairs/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["scaling"] == "weak"  # (at least one warm-up step is always run)
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_under_torchrun_only_rank0_prints():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29599")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


@pytest.mark.parametrize("name", ["r1_bench_n1.json", "r1_bench_n8.json", "r2_bench_n1.json", "r2_bench_n2.json", "r2_bench_n4.json", "r2_bench_n8.json"])
def test_recorded_gpu_line_has_every_contract_field(name):
    d = last_json_line(open(os.path.join(ROOT, "profiles", name)).read())
    assert BASE_KEYS - {"cpu_baseline"} <= set(d) and "impl" not in d
    ref = last_json_line(open(os.path.join(ROOT, "profiles", name[:2] + "_bench_reference_arm.json")).read())
    assert d["metric"] == ref["metric"] and d["unit"] == ref["unit"]
    assert d["gpu_launches"] > 0 and d["steps"] >= 1 and d["warmup"] >= 3 and d["vs_baseline"] is None
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] <= d["value"] * 1.01
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    rf = d["roofline"]
    assert set(rf) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and 0 < rf["frac"] < 1
    assert "workload" in d["config"] and d["dtype"] == "f32" and d["data"] == "synthetic"
    if d["n_gpus"] == 1:
        cb = d["cpu_baseline"]
        assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_round2_lines_carry_parity_and_the_strong_scaling_record(n):
    """Round 2: every case of the line ends with an oracle check of 32 sampled queries (0 mismatches), config 4 is
    measured as a strong-scaling record at every N (100 M songs total), config 5 with sharded queries."""
    d = last_json_line(open(os.path.join(ROOT, "profiles", f"r2_bench_n{n}.json")).read())
    assert d["n_gpus"] == n and d["scaling"] == "weak"
    assert d["parity_check"]["queries"] >= 32 and d["parity_check"]["mismatches"] == 0
    st = d["strong_scaling"]
    assert st["scaling"] == "strong" and st["n_gpus"] == n and st["parity_check"]["mismatches"] == 0
    assert "100000000 songs TOTAL" in st["config"]["workload"] and st["ms_per_step"] > 0
    ap = d["all_pairs_1M"]
    assert ap["parity_check"]["mismatches"] == 0 and ap["parity_check"]["table_complete"] and ap["seconds"] > 0
    if n == 1:
        assert d["config3_top100"]["parity_check"]["mismatches"] == 0
        assert d["roofline"]["traffic_source"].startswith("stored")
        cases = {c["queries"]: c for c in d["roofline_hbm_regime"]["cases"]}
        assert cases[1]["frac_call"] >= 0.75 and cases[1]["frac_scan_kernel"] >= 0.9
