"""One rank of bench.py's GPU arm with fake engines over gloo (launched by tests/test_bench_contract.py through
torch.distributed.run): the multi-rank control flow -- unique-id broadcast, barriers, max-over-ranks reductions, the
common decision of the budget guard, rank 0 printing the line -- without a device."""
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
os.environ["NSB_BENCH_BACKEND"] = "gloo"

import torch  # noqa: E402

import bench  # noqa: E402
import navierstokes_project_nm4pde_b200 as pkg  # noqa: E402
from test_bench_contract import _FakeEngine  # noqa: E402

rank = int(os.environ["RANK"])


class FakeRun:
    def __init__(self, workload, args, world, rank_, local_rank, uid=None):
        assert uid == b"u" * 128  # rank 0's id reached every rank
        self.e = _FakeEngine()
        self.mesh = types.SimpleNamespace(n_cells=1000)
        self.prob = types.SimpleNamespace(_dir_rows=list(range(30)), N=2600, transport="p2p")
        self.variant, self.n_dofs, self.dt = bench.WORKLOADS[workload][0], 5000, 2e-4
        self.ilu_ordering, self.ilu_ordering_schur, self.n = 1, 1, 0

    def step(self):
        self.n += 1
        dt = 0.02 * (1 + rank) * (20.0 if self.n > 3 else 1.0)  # rank 1 is the slower one; timed steps are slow
        time.sleep(dt)
        self.e.dev_ms += 1e3 * dt
        return 40 + self.n

    def prepare(self):
        return [self.step(), self.step()]


bench.GpuRun = FakeRun
pkg.Engine.unique_id = staticmethod(lambda: b"u" * 128)
torch.cuda.set_device = lambda *_: None
torch.cuda.synchronize = lambda *_: None
bench.ClockSampler = lambda *_: types.SimpleNamespace(start=lambda: None,
                                                      stop=lambda: {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []})
sys.argv = ["bench.py"] + sys.argv[1:]
bench.main()
