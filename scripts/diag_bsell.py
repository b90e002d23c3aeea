"""Diagnostic (GPU): block multicolour ILU apply -- determinism, linearity, and comparison with the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from navierstokes_project_nm4pde_b200 import HostMesh, NavierStokes
wl = sys.argv[1] if len(sys.argv) > 1 else "cyl3d-270k"
order = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with_oracle = len(sys.argv) > 3 and sys.argv[3] == "oracle"
DT = bench.DELTAT["3d"]
s, nz = bench.WORKLOADS[wl][1]
mesh = HostMesh.cylinder3d(s, nz)
p = NavierStokes(mesh, "3d", T=1.0, deltat=DT, test_case=2, ilu_ordering=order, orthogonalisation=1)
p.setup()
e = p.engine
rng = np.random.default_rng(20240607)
x0 = np.zeros(p.N); x0[:p.n_u] = 0.05 * rng.uniform(-1, 1, p.n_u)
e.set_solution(x0); e.set_dirichlet_values(p.dirichlet_values(DT)); e.assemble_first(DT); e.precond_init()
for which, n in ((0, p.n_u), (1, p.n_p)):
    x, z = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    a1, a2 = e.ilu_apply(which, x), e.ilu_apply(which, x)
    print(wl, "which", which, "deterministic:", np.array_equal(a1, a2), "max diff", np.abs(a1 - a2).max())
    lhs = e.ilu_apply(which, 0.5 * x - 3.0 * z); rhs = 0.5 * a1 - 3.0 * e.ilu_apply(which, z)
    print("   linearity rel err", np.linalg.norm(lhs - rhs) / np.linalg.norm(rhs), "sweeps", e.stat("sweeps_F" if which == 0 else "sweeps_S"))
    if with_oracle:
        from oracle import ns_ref as R
        d = p.dofs
        if which == 0:
            num = dict(dim=3, cell_dofs=d.cell_dofs(), N=d.N, n_u=d.n_u, n_p=d.n_p, dpc=d.dpc)
            o = R.Oracle(3, "3d", mesh.vertices, mesh.cells, num, R.system_pattern(num), 1e-3, DT)
            ou = (3 * e.ilu_order(0)[:, None] + np.arange(3)[None, :]).ravel()
            o.set_ilu_order(ou, e.ilu_order(1))
            o.set_dirichlet(p._dir_rows, p.dirichlet_values(DT)); o.set_solution(x0); o.assemble_first(); o.precond_init("yosida")
        ref = o.ilu_apply(which, x)
        err = np.abs(a1 - ref)
        print("   vs oracle rel l2", np.linalg.norm(a1 - ref) / np.linalg.norm(ref), "worst entries", np.argsort(err)[-5:], err.max(), np.abs(ref).max())
