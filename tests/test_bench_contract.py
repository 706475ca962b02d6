"""The JSON line contract of bench.py, checked on the CPU: the reference arm runs here (it is the CPU implementation), and the
committed B200 line of the default workload (profiles/) must carry every key the driver and the judge read."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "e2e", "gpu_launches"}


def _check_common(d):
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert "workload" in d["config"] and "model" not in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["vs_baseline"] is None                      # BASELINE.md publishes no number for this metric
    assert isinstance(d["higher_is_better"], bool) and d["scaling"] in ("weak", "strong")


def test_committed_default_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_n1_session4_final.json")))
    _check_common(d)
    assert d["n_gpus"] == 1 and d["gpu_launches"] > 0 and d["warmup"] >= 3
    assert d["e2e"]["h2d_bytes_per_step"] == (1 << 22) * 16 and d["e2e"]["value"] < d["value"]      # host buffers in: strictly slower than resident
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.5 < r["int_pipe"]["frac_of_alu_pipe"] <= 1.0                                            # the binding roofline
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("reference", "port") and c["unit"] == d["unit"]
    k = d["clocks"]
    assert k["sm_mhz"] >= 0.95 * k["sm_max_mhz"] and not any("slowdown" in x for x in k["reasons"])


def test_committed_round2_line_states_the_binding_roofline_and_carries_the_sub_records():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n1_final.json")))
    _check_common(d)
    assert d["n_gpus"] == 1 and d["gpu_launches"] > 0 and d["warmup"] >= 3
    assert d["e2e"]["h2d_bytes_per_step"] == (1 << 22) * 16 and d["e2e"]["value"] < d["value"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] == "int-alu-pipe" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] <= 1.05
    assert r["hbm"]["bound"] == "hbm" and r["hbm"]["frac"] < 0.1                  # the HBM view is reported, and is not the binding one
    assert d["parity"]["columns_checked_on_rank0"] == [0]                          # the oracle digest was reproduced before timing
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample", "extrapolated"} <= set(c) and c["kind"] == "port" and c["unit"] == d["unit"]
    assert "EXTRAPOLATION" in c["extrapolated"]["how"]
    k = d["clocks"]
    assert k["sm_mhz"] >= 0.95 * k["sm_max_mhz"] and not any("slowdown" in x for x in k["reasons"])
    sub = d["configs"]
    for key in ("configs[3]", "configs[4]"):
        _check_common(sub[key])
        assert sub[key]["parity"]["checked_before_timing"] and sub[key]["clocks"]["sm_mhz"] and sub[key]["e2e"]["h2d_bytes_per_step"] > 0


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-log-n", "10"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    _check_common(d)
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    # the arm reports the workload the B200 arm runs (same config object); its bounded sample is described in cpu_baseline
    assert d["config"]["log_n"] == 24 and d["steps"] == 1 and d["warmup"] == 0 and d["cpu_baseline"]["sample_log_n"] == 10
