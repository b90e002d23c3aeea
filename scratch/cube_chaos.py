import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import helpers as T
from navierstokes_project_nm4pde_b200.distributed import build_local_problem
case = T.Case("cube")
d = case.dofs
cell_part = case.mesh.partition(2)
cd = d.cell_dofs(copy=False)
part = np.full(d.N, 2, np.int32); np.minimum.at(part, cd.ravel(), np.repeat(cell_part, cd.shape[1]))
for rtol in (1e-2, 1e-6):
    o = case.oracle(); o.set_partition(part)
    if rtol != 1e-2: o.set_options(inner_rtol=rtol)
    rows, vals = case.bc(0.0); o.set_dirichlet(rows, vals); o.set_solution(case.initial())
    t = 0.0; its=[]
    for step in range(3):
        o.set_neumann_rhs(case.neumann(t)); t += case.dt
        rows, vals = case.bc(t); o.set_dirichlet_values(vals)
        (o.assemble_first if step == 0 else o.assemble_step)()
        rc, k, _ = o.solve_step("yosida"); its.append(k)
    print("threads", os.environ.get("OMP_NUM_THREADS"), "inner_rtol", rtol, "N", d.N, its)
