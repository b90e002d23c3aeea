"""Profiling driver: set up a workload, assemble once, then launch each hot kernel a few times."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from navierstokes_project_nm4pde_b200 import HostMesh, NavierStokes
wl = sys.argv[1] if len(sys.argv) > 1 else "cyl3d-2M"
order = int(sys.argv[2]) if len(sys.argv) > 2 else 1
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
s, nz = bench.WORKLOADS[wl][1]
prob = NavierStokes(HostMesh.cylinder3d(s, nz), "3d", T=1.0, deltat=bench.DELTAT['3d'], test_case=2, ilu_ordering=order)
prob.setup()
e = prob.engine
rng = np.random.default_rng(20240607)
x = rng.uniform(-1, 1, prob.N); x[prob.n_u:] = 0
e.set_solution(x)
e.set_dirichlet_values(prob.dirichlet_values(bench.DELTAT['3d']))
e.assemble_first(bench.DELTAT['3d'])
e.precond_init()
kernels = sys.argv[4].split(",") if len(sys.argv) > 4 else ["assemble_step", "spmv_F", "spmv_system", "ilu_F", "spmv_S", "ilu_S", "dot", "axpy", "add_and_dot"]
for k in kernels:
    ms, b = e.bench_kernel(k, iters=iters, flush_l2=True)
    print(f"{k:14s} {ms:9.4f} ms  {b/1e6:10.1f} MB  {b/ms/1e6:8.1f} GB/s", flush=True)
print("levels", [e.stat(k) for k in ("levels_F_fwd","levels_F_bwd","levels_S_fwd","levels_S_bwd","sweeps_F","sweeps_S")])
