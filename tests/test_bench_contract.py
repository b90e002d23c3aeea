"""bench.py contract, CPU side: the reference arm prints ONE JSON line with the keys the driver
reads, on the arm's own config / metric / unit, and never touches the CUDA library."""
import json
import os
import subprocess
import sys

import pytest

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


# ---- the GPU arm's host logic with a fake engine (no device): budget guard, failure handling, the JSON line ----------
class _FakeEngine:
    def __init__(self, fail_at=None):
        self.calls, self.fail_at, self.dev_ms = 0, fail_at, 0.0

    def launch_count(self, reset=False):
        return 1234

    def stat(self, key):
        if key == "t_step_dev_ms_reset":
            v, self.dev_ms = self.dev_ms, 0.0
            return v
        return {"t_prec_ms": 2.0, "t_solve_ms": 8.0}.get(key, 7.0)

    def bench_kernel(self, name, iters=5, flush_l2=True):
        if name == "spmv_S":
            raise RuntimeError("cannot be timed alone")
        return 1.0, 2.0e9


def _fake_gpurun(bench, step_s, fail_at=None, slow_after=None):
    import time as _time

    class FakeRun:
        def __init__(self, workload, args, world, rank, local_rank, uid=None):
            import types

            self.e = _FakeEngine()
            self.mesh = types.SimpleNamespace(n_cells=1000)
            self.prob = types.SimpleNamespace(_dir_rows=list(range(30)), N=5000)
            self.variant, self.n_dofs, self.dt = bench.WORKLOADS[workload][0], 5000, 2e-4
            self.ilu_ordering, self.ilu_ordering_schur, self.n = 2, 1, 0

        def step(self):
            self.n += 1
            if fail_at is not None and self.n == fail_at:
                raise RuntimeError("SolverControl::NoConvergence: inner GMRES")
            dt = step_s * (20.0 if slow_after is not None and self.n > slow_after else 1.0)
            _time.sleep(dt)
            self.e.dev_ms += 1e3 * dt
            return 40 + self.n

        def prepare(self):
            return [self.step(), self.step()]

    return FakeRun


def _run_gpu_arm(monkeypatch, capsys, argv, step_s=0.01, fail_at=None, budget=None, slow_after=None):
    import importlib
    import types

    import torch

    sys.path.insert(0, ROOT)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.delenv("RANK", raising=False)
    if budget is not None:
        monkeypatch.setenv("NSB_BENCH_BUDGET_S", str(budget))
    bench = importlib.import_module("bench")
    monkeypatch.setattr(bench, "GpuRun", _fake_gpurun(bench, step_s, fail_at, slow_after))
    monkeypatch.setattr(torch.cuda, "set_device", lambda *_: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *_: None)
    monkeypatch.setattr(bench, "ClockSampler", lambda *_: types.SimpleNamespace(
        start=lambda: None, stop=lambda: {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []}))
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    bench.main()
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.strip()]
    assert len(lines) == 1
    return bench, json.loads(lines[0])


def test_gpu_arm_prints_the_contract_line(monkeypatch, capsys):
    bench, d = _run_gpu_arm(monkeypatch, capsys, ["--steps", "4", "--warmup", "3", "--no-cpu-baseline"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "roofline", "clocks", "gpu_launches"):
        assert key in d, key
    assert d["steps"] == 4 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["gpu_launches"] == 1234
    assert d["detail"]["truncated"] is False and d["detail"]["failure"] is None
    assert d["value"] == pytest.approx(5000 * 4 / (d["ms_per_step"] * 4e-3)) and d["e2e"]["value"] <= d["value"] * 1.001
    assert d["e2e"]["h2d_bytes_per_step"] == 30 * 8 and d["e2e"]["d2h_bytes_per_step"] == 5000 * 8
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["frac"] == pytest.approx(r["achieved"] / r["peak"]) and "spmv_S" not in r["kernels"]
    assert d["detail"]["same_mesh"]["steps"] == 5  # the like-for-like leg on the mesh of the CPU arm


def test_gpu_arm_times_fewer_steps_when_the_budget_runs_out(monkeypatch, capsys):
    """20 steps of 0.4 s do not fit what a 4 s budget leaves: fewer steps are timed and the line says so."""
    _, d = _run_gpu_arm(monkeypatch, capsys, ["--steps", "20", "--warmup", "1", "--no-cpu-baseline"], step_s=0.4, budget=29.0)
    assert 1 <= d["steps"] < 20 and d["detail"]["truncated"] is True and d["detail"]["steps_requested"] == 20
    assert len(d["detail"]["outer_iterations"]) == d["steps"]


def test_gpu_arm_stops_inside_the_timed_loop_when_steps_turn_out_slower(monkeypatch, capsys):
    """The warm-up estimate (0.05 s per step) lets all 6 steps through, but the timed steps cost 1 s each: the guard
    inside the loop ends the measurement instead of overrunning the budget."""
    _, d = _run_gpu_arm(monkeypatch, capsys, ["--steps", "6", "--warmup", "1", "--no-cpu-baseline"], step_s=0.05, budget=47.0,
                        slow_after=3)
    assert 1 <= d["steps"] < 6 and d["detail"]["truncated"] is True and d["detail"]["failure"] is None
    assert d["ms_per_step"] == pytest.approx(1000.0, rel=0.05)


def test_gpu_arm_keeps_the_finished_steps_when_a_step_raises(monkeypatch, capsys):
    # 2 start-up + 1 warm-up steps, then the 3rd timed step (6th call) raises
    _, d = _run_gpu_arm(monkeypatch, capsys, ["--steps", "5", "--warmup", "1", "--no-cpu-baseline"], fail_at=6)
    assert d["steps"] == 2 and d["detail"]["truncated"] is True and "NoConvergence" in d["detail"]["failure"]
    assert len(d["detail"]["outer_iterations"]) == 2 and d["value"] > 0


def test_gpu_arm_on_two_ranks_over_gloo():
    """world_size 2: one line from rank 0, times are the max over ranks (rank 1 is twice as slow), both ranks stop at
    the same step when the budget guard trips."""
    env = dict(os.environ, NSB_BENCH_BUDGET_S="27.5", OMP_NUM_THREADS="1")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29653", os.path.join(ROOT, "tests", "_bench_fake_rank.py"), "--gpus", "2",
                        "--steps", "6", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["scaling"] == "strong" and d["config"]["transport"] == "p2p" and "cpu_baseline" not in d
    assert 1 <= d["steps"] < 6 and d["detail"]["truncated"] is True  # extras 25 s + 0.8 s per slow step against 27.5 s
    assert d["ms_per_step"] == pytest.approx(800.0, rel=0.1)         # rank 1's time, not rank 0's 400 ms


def test_ordering_selection_of_the_gpu_arm():
    """What bench.py runs by default: block multicolour F_s + point multicolour Schur factors at 19.9 M DoF on one GPU,
    point multicolour for both from two GPUs on, block-ordered Schur factors for the 2D family; explicit choices win."""
    sys.path.insert(0, ROOT)
    import bench

    n = bench.N_DOFS["cyl3d-20M"]
    assert bench.pick_orderings("3d", n, 1) == (2, 1)
    assert [bench.pick_orderings("3d", n, w) for w in (2, 4, 8)] == [(1, 1)] * 3
    assert bench.pick_orderings("3d", bench.N_DOFS["cyl3d-2M"], 1) == (1, 1)
    assert bench.pick_orderings("2d", 1962041, 1) == (1, 2) and bench.pick_orderings("2d", 1962041, 2) == (1, 2)
    assert bench.pick_orderings("3d", n, 1, 0, -1) == (0, 0) and bench.pick_orderings("3d", n, 1, 3, 1) == (3, 1)
    assert bench.pick_orderings("3d", n, 8, 2, -1) == (2, 1)
