import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np
from navierstokes_project_nm4pde_b200 import HostMesh, NavierStokes
from oracle import ns_ref as R
import scipy.sparse as sp
DT=2e-4
def greedy_colour(rp, ci, n):
    col = -np.ones(n, np.int32)
    for i in range(n):
        used = set(col[ci[rp[i]:rp[i+1]]].tolist())
        c = 0
        while c in used: c += 1
        col[i] = c
    return col
def run(s, nz, mode, nparts=0, steps=2):
    mesh = HostMesh.cylinder3d(s, nz)
    prob = NavierStokes(mesh, "3d", T=1.0, deltat=DT, test_case=2)
    prob.setup_host()
    d = prob.dofs
    num = dict(dim=3, cell_dofs=d.cell_dofs(), N=d.N, n_u=d.n_u, n_p=d.n_p, dpc=d.dpc)
    pat = R.system_pattern(num)
    o = R.Oracle(3, "3d", mesh.vertices, mesh.cells, num, pat, 1e-3, DT)
    if mode == "mc":
        rp, ci = pat[0], pat[1]
        A = sp.csr_matrix((np.ones(len(ci)), ci, rp), shape=(d.N, d.N))
        # node graph = F block rows 0::3, cols//3
        Fu = A[:d.n_u][:, :d.n_u].tocsr()
        Fn = Fu[0::3][:, 0::3].tocsr()
        coln = greedy_colour(Fn.indptr, Fn.indices, Fn.shape[0])
        order_n = np.argsort(coln, kind="stable")
        order_u = (3*order_n[:,None] + np.arange(3)[None,:]).ravel().astype(np.int32)
        # pressure: pattern of S = B Bt
        B = A[d.n_u:][:, :d.n_u]; S = (B @ B.T).tocsr()
        colp = greedy_colour(S.indptr, S.indices, S.shape[0])
        order_p = np.argsort(colp, kind="stable").astype(np.int32)
        print("colours", coln.max()+1, colp.max()+1)
        o.set_ilu_order(order_u, order_p)
    if nparts:
        cell_part = mesh.partition(nparts)
        cd = d.cell_dofs(copy=False)
        part = np.full(d.N, nparts, np.int32)
        np.minimum.at(part, cd.ravel(), np.repeat(cell_part, cd.shape[1]))
        o.set_partition(part)
    o.set_dirichlet(prob._dir_rows, prob.dirichlet_values(DT))
    o.set_solution(np.zeros(d.N))
    t0=time.time()
    o.assemble_first()
    rc,k,_ = o.solve_step("yosida")
    print(f"{mode} P={nparts} N={d.N}: first {k} outer, F {o.stat('n_inner_F')}/{o.stat('n_F_solves')} S {o.stat('n_inner_S')}/{o.stat('n_S_solves')}  {time.time()-t0:.1f}s", flush=True)
    for _ in range(steps):
        t0=time.time()
        o.assemble_step(); rc,k,_ = o.solve_step("yosida")
        print(f"   step {k} outer, F {o.stat('n_inner_F')}/{o.stat('n_F_solves')}={o.stat('n_inner_F')/max(1,o.stat('n_F_solves')):.1f} S {o.stat('n_inner_S')}/{o.stat('n_S_solves')}={o.stat('n_inner_S')/max(1,o.stat('n_S_solves')):.1f}  {time.time()-t0:.1f}s", flush=True)
s, nz = int(sys.argv[1]), int(sys.argv[2])
for mode, P in [("nat",0),("mc",0),("nat",64),("nat",512)]:
    run(s, nz, mode, P)
