"""CPU checks of bench.py's measurement contract: the dt schedule, the clock/throttle sampler's parsing, the peak lookup,
the reference arm's JSON line (run on a tiny sample grid) and the command-line surface the driver uses."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_dt_schedule_is_the_capped_doubling_list():
    assert bench.dt_schedule(7) == pytest.approx([0.1, 0.2, 0.4, 0.8, 1.6, 1.6, 1.6])


def test_clock_sampler_parses_nvidia_smi_lines():
    s = bench.ClockSampler(0)
    s.samples = [["1965", "1965", "Not Active", "Not Active", "Not Active", "Active"],
                 ["1770", "1965", "Not Active", "Not Active", "Not Active", "Active"],
                 ["garbage"],
                 ["1800", "1965", "Not Active", "Active", "Not Active", "Not Active"]]
    out = s.finish()
    assert out == {"sm_mhz": 1800.0, "sm_max_mhz": 1965.0, "reasons": ["hw_thermal_slowdown", "sw_power_cap"], "samples": 3}


def test_peak_comes_from_the_driver_file_or_the_documented_fallback():
    peak, source = bench.measured_peak()
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
        assert source.startswith("measured") and peak == json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    else:
        assert source.startswith("fallback") and peak == 6650.0


def test_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the oracle port on the host): exactly one stdout line carrying the base contract's
    keys, the same metric/unit as the b200 arm, impl = reference, an e2e object with zero copy bytes and a cpu_baseline."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-grid", "12"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "e2e", "cpu_baseline"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == bench.host_cores() and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["workload"].startswith("b13d_kuhn3d_")


def test_command_line_surface():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--grid", "--workload", "--basis", "--jac-fp32"):
        assert flag in r.stdout, flag
