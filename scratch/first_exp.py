import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from navierstokes_project_nm4pde_b200 import HostMesh, NavierStokes
from oracle import ns_ref as R
DT=2e-4
s, nz = int(sys.argv[1]), int(sys.argv[2])
mesh = HostMesh.cylinder3d(s, nz)
prob = NavierStokes(mesh, "3d", T=1.0, deltat=DT, test_case=2)
prob.setup_host()
d = prob.dofs
num = dict(dim=3, cell_dofs=d.cell_dofs(), N=d.N, n_u=d.n_u, n_p=d.n_p, dpc=d.dpc)
pat = R.system_pattern(num)
for first_rtol, cap in [(1e-2, 0), (1e-4, 0), (1e-2, 28)]:
    o = R.Oracle(3, "3d", mesh.vertices, mesh.cells, num, pat, 1e-3, DT)
    cell_part = mesh.partition(8); cd = d.cell_dofs(copy=False)
    part = np.full(d.N, 8, np.int32); np.minimum.at(part, cd.ravel(), np.repeat(cell_part, cd.shape[1]))
    o.set_partition(part)
    o.set_dirichlet(prob._dir_rows, prob.dirichlet_values(DT))
    o.set_solution(np.zeros(d.N))
    o.assemble_first()
    o.set_options(inner_rtol=first_rtol, outer_maxit=cap if cap else 100000)
    t0=time.time(); rc,k,res = o.solve_step("yosida")
    print(f"first rtol={first_rtol} cap={cap}: rc {rc} {k} outer res {res:.2e}, F {o.stat('n_inner_F')}/{o.stat('n_F_solves')} S {o.stat('n_inner_S')}/{o.stat('n_S_solves')} {time.time()-t0:.1f}s", flush=True)
    o.set_options(inner_rtol=1e-2, outer_maxit=100000)
    for _ in range(4):
        t0=time.time(); o.assemble_step(); rc,k,res = o.solve_step("yosida")
        print(f"   step {k} outer, F {o.stat('n_inner_F')}/{o.stat('n_F_solves')} S {o.stat('n_inner_S')}/{o.stat('n_S_solves')} {time.time()-t0:.1f}s", flush=True)
