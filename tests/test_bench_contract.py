"""bench.py contract, CPU side: the reference arm prints ONE JSON line with the keys the driver
reads, on the arm's own config / metric / unit, and never touches the CUDA library."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "cyl3d-30k"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "DoF-timesteps/s"
    assert d["metric"].startswith("DoF-timesteps/sec") and d["dtype"] == "f64" and d["vs_baseline"] is None
    # the arm labels its config with the mesh it actually runs and says which headline workload it samples
    assert d["config"]["workload"] == "cyl3d-30k" and d["config"]["sample_of"] == "cyl3d-20M" and d["gpu_launches"] == 0
    assert d["steps"] == 1 and d["detail"]["truncated"] is False
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] > 0 and "cyl3d-30k" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_other_ranks_of_the_reference_arm_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_2d_sample():
    """configs[3] (2D refined cylinder, aSIMPLE) has a CPU arm too."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--workload", "cyl2d-2M", "--cpu-sample", "cyl2d-3k"], capture_output=True, text=True, timeout=600,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip())
    assert d["config"]["variant"] == "NavierStokes2D" and d["config"]["preconditioner"] == "asimple"
    assert d["config"]["sample_of"] == "cyl2d-2M" and d["steps"] == 2 and d["value"] > 0
