"""bench.py's contract, checked on the CPU: the committed bench lines of the round (profiles/r2_bench_*_head.json, written by
`python bench.py` on B200 boxes) carry every key the driver reads and their numbers are consistent with each other, with
MEASURED_PEAKS.json and with the workload; and bench.py refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = {n: os.path.join(ROOT, "profiles", "r2_bench_n%d_head.json" % n) for n in (1, 2, 8)}
STATES = 100_000_000                    # BASELINE.json configs[4]: synthetic toggle, ~1e8 FSP states


def _line(n):
    if not os.path.exists(LINES[n]):
        pytest.skip("no committed bench line for %d GPUs" % n)
    with open(LINES[n]) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def _peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if not os.path.exists(p):
        pytest.skip("MEASURED_PEAKS.json absent")
    return json.load(open(p))["hbm_gbs"]


@pytest.mark.parametrize("n", [1, 2, 8])
def test_line_has_the_contract_keys(n):
    d = _line(n)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == n and d["warmup"] >= 3 and d["steps"] >= 1
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["higher_is_better"] is True
    assert d["scaling"] == "strong"                      # 1e8 states in total at every GPU count
    assert d["vs_baseline"] is None                      # BASELINE.json publishes no number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["states"] == STATES
    assert d["gpu_launches"] > 0
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # value = state updates of the whole job / device time of a solve
    assert d["value"] == pytest.approx(STATES * d["spmv_per_solve"] / (1e-3 * d["ms_per_step"]), rel=1e-9)


@pytest.mark.parametrize("n", [1, 2, 8])
def test_roofline_is_consistent(n):
    d = _line(n)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    # the measured copy bandwidth of the pool's B200s (MEASURED_PEAKS.json is re-written by the driver per pod, so the committed
    # line is only required to sit within 5 % of the file's current value)
    assert r["peak_kind"] == "measured" and r["peak"] == pytest.approx(_peak(), rel=0.05)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-12)
    assert 0.0 < r["frac"] < 1.0
    # achieved = algorithmic bytes per launch / measured launch time
    assert r["achieved"] == pytest.approx(r["algorithmic_bytes_per_launch"] / (1e-3 * r["avg_launch_ms"]) / 1e9, rel=1e-9)
    # the bytes are this rank's rows x the per-state figure of DESIGN.md section 4 (<= 40 B/state for a column launch)
    per_state = r["algorithmic_bytes_per_launch"] / (STATES / n)
    assert per_state == pytest.approx(r["algorithmic_bytes_per_state_per_launch"], rel=1e-9)
    assert 16.0 <= per_state <= 40.0
    # the dominant kernel's time inside one solve cannot exceed the solve
    assert 0.0 < r["share_of_step"] <= 1.0
    # DRAM traffic (ncu, per launch) is named with its source and is not below the algorithmic bytes by more than rounding
    assert r["traffic"] is None or (r["traffic_source"] and r["traffic"] >= 0.99 * r["algorithmic_bytes_per_launch"])


@pytest.mark.parametrize("n", [1, 2, 8])
def test_e2e_and_parity(n):
    d = _line(n)
    e = d["e2e"]
    assert e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] >= 8 * STATES and e["d2h_bytes_per_step"] >= 8 * STATES     # p in, p out: fp64 per state
    assert e["value"] < d["value"]                        # transfers are inside the timed region
    assert e["value"] == pytest.approx(STATES * d["spmv_per_solve"] / (1e-3 * e["ms_per_step"]), rel=1e-9)
    assert d["parity"]["lattice_vs_explicit_bit_identical"] is True and d["parity"]["lattice_vs_explicit_max_abs"] == 0.0
    assert d["parity"]["states_compared"] == STATES
    if n > 1:
        assert d["dist_parity"]["dist_bit_identical"] is True
    else:
        c = d["cpu_baseline"]
        assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
        assert c["unit"] == d["unit"]


def test_reference_arm_line():
    p = os.path.join(ROOT, "profiles", "r2_bench_reference_head.json")
    if not os.path.exists(p):
        pytest.skip("no committed reference-arm line")
    d = json.loads(open(p).read().strip().splitlines()[-1])
    own = _line(1)
    assert d["impl"] == "reference"
    for k in ("metric", "unit", "higher_is_better"):
        assert d[k] == own[k]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
    assert not any(l.startswith("{") for l in r.stdout.splitlines())       # no metric line from a machine without a GPU
