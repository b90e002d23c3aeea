"""Profiling driver for the launch-chain / L2 knobs of the triangular solves: one setup, then the ILU applies are
re-captured and timed under each environment variant (the knobs are read when a solve is captured:
csrc/kernels_linalg.cu `ilu_reset_graphs`).
usage: prof_variants.py <workload> <ilu_ordering> <iters> "<VAR=val,VAR=val;VAR=val;...>"   (';' separates variants)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import bench
from navierstokes_project_nm4pde_b200 import HostMesh, NavierStokes

wl, order, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
variants = [v for v in sys.argv[4].split(";")] if len(sys.argv) > 4 else [""]
KNOBS = ("NSB_PDL", "NSB_L2_PERSIST_MB", "NSB_L2_FETCH", "NSB_BSELL_PREFETCH", "NSB_BSELL_PIPE")
s, nz = bench.WORKLOADS[wl][1]
prob = NavierStokes(HostMesh.cylinder3d(s, nz), "3d", T=1.0, deltat=bench.DELTAT["3d"], test_case=2, ilu_ordering=order)
prob.setup()
e = prob.engine
rng = np.random.default_rng(20240607)
x = rng.uniform(-1, 1, prob.N)
x[prob.n_u:] = 0
e.set_solution(x)
e.set_dirichlet_values(prob.dirichlet_values(bench.DELTAT["3d"]))
e.assemble_first(bench.DELTAT["3d"])
e.precond_init()
print(f"{wl} ordering {order}: sweeps F {e.stat('sweeps_F'):.0f}, S {e.stat('sweeps_S'):.0f}", flush=True)
for var in variants:
    for k in KNOBS:
        os.environ.pop(k, None)
    for kv in filter(None, var.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    e.bench_kernel("reset_graphs", iters=1, flush_l2=False)
    out = []
    for kern in ("ilu_F", "ilu_S", "spmv_F"):
        ms, b = e.bench_kernel(kern, iters=iters, flush_l2=True)
        out.append(f"{kern} {ms:8.4f} ms {b / ms / 1e6:7.1f} GB/s")
    print(f"[{var or 'base':40s}] " + " | ".join(out), flush=True)
