"""bench.py contract (CPU side): the reference arm prints ONE JSON line with the agreed keys, runs without a GPU and
never touches /root/reference at run time (it times oracle/layers_ref.py, the pinned port)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--batch", "2"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "scene-flow pairs/sec @8192 pts" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and abs(d["value"] - 2e3 / d["ms_per_step"]) < 1e-6 * d["value"]       # --batch pairs per step
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("configs[2]") and d["vs_baseline"] is None
    # same ``config`` object as the kdpc arm prints for the same --batch (the driver's same_config check)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(2)


def test_roofline_traffic_comes_from_the_committed_capture():
    sys.path.insert(0, ROOT)
    import bench
    cap = bench.ncu_capture_of_roofline_kernel()
    assert cap["source"] and cap["source"].startswith("profiles/") and cap["traffic"] > 1e6
