"""bench.py's contract pieces that run without a GPU: the reference arm (the compiled, unmodified fish_mod.cpp when
oracle/_ref is present, else the oracle port) and the JSON keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_valid_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--fish", "200000"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "agent_steps_per_sec" and d["unit"] == "agent-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_roofline_helpers():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.flops_alg(10, 2) == 13 * 10 + 20 * 2
    peaks, how = bench.measured_peaks()
    assert peaks["hbm_gbs"] > 1000 and how in ("measured", "fallback")
    cap = bench.ncu_capture("neighbour_kernel_v4")
    assert cap is None or (cap["dram_bytes"] > 0 and 0 < cap["issue_active_pct"] <= 100)
