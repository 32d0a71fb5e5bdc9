"""The CPU-runnable half of bench.py's contract: `--impl reference` prints ONE JSON line with the keys the driver reads,
from the unmodified reference in baseline/_ref when it is installed (kind "reference") or from the C port (kind "port")."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "agent-steps/s" and d["unit"] == "agent-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("c3:")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]
    assert cb["kind"] == ("reference" if os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "dist_classicrl")) else "port")
    assert d["e2e"] == {"value": d["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


import pytest  # noqa: E402


@pytest.mark.gpu
def test_our_arm_prints_the_full_contract_line():
    """One short run of the real thing on the GPU: every key the driver and the judge read is there and consistent."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "16", "--warmup", "3"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "cpu_baseline", "atomics_mode"):
        assert key in d, key
    assert d["metric"] == "agent-steps/s" and d["n_gpus"] == 1 and d["steps"] == 16 and d["warmup"] >= 3
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("c3:") and "model" not in d["config"]
    n = d["config"]["agents"]
    assert abs(d["value"] - n / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["algorithmic_bytes_per_agent_step"] == 8 * d["config"]["actions"] + 12
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["avg_launch_ms"] * 1e-3) / 1e9) <= 1e-6 * r["achieved"]
    assert 0.0 < r["frac"] < 1.0 and (r["traffic"] is None or r["traffic"] > 0)
    e = d["e2e"]
    assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] >= n * 16 and e["d2h_bytes_per_step"] >= n * 4
    assert d["gpu_launches"] == 2  # 16 steps = two fused launches of 8
    c = d["clocks"]
    assert c["samples"] >= 1 and c["sm_mhz"] and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] > 0 and cb["cores"] >= 1 and cb["sample"]
    assert d["atomics_mode"]["value"] > d["value"]  # dropping the ordering guarantee is never slower
    assert d["config"]["td_update_form"] == "one-pass pipeline" and r["kernel"].startswith("fused_flow_kernel")
    assert d["config"]["warmup_requested"] == 3 and d["config"]["warmup_used"] == d["warmup"]
    vl = d["value_long"]
    assert 0 < vl["value"] and vl["vector_steps"].startswith("8..") and len(vl["ms_per_step_by_32_step_window"]) >= 15
    assert r["gather_peak"]["value"] > r["achieved"]  # dependency-free gathers over the same table bound the loop from above
    assert "multiprocessing" in cb and cb["mpi"]["value"] is None
