import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, scipy.sparse as sp
from navierstokes_project_nm4pde_b200 import HostMesh, NavierStokes
from oracle import ns_ref as R
DT=2e-4
def greedy_colour(rp, ci, n, order=None):
    col = -np.ones(n, np.int32)
    it = range(n) if order is None else order
    for i in it:
        used = set(col[ci[rp[i]:rp[i+1]]].tolist())
        c = 0
        while c in used: c += 1
        col[i] = c
    return col
def bfs_blocks(rp, ci, n, bs=32):
    """greedy BFS aggregation: seeds in natural order; returns block id per node, position order"""
    blk = -np.ones(n, np.int64); order = []
    nb = 0
    from collections import deque
    for seed in range(n):
        if blk[seed] >= 0: continue
        q = deque([seed]); blk[seed] = nb; cnt = 1; members=[seed]
        while q and cnt < bs:
            u = q.popleft()
            for v in ci[rp[u]:rp[u+1]]:
                if blk[v] < 0:
                    blk[v] = nb; members.append(v); q.append(v); cnt += 1
                    if cnt >= bs: break
        order.extend(members); nb += 1
    return blk, np.array(order), nb
def stats(name, Fn, blk, nb):
    coo = Fn.tocoo()
    off = coo.row != coo.col
    intra = (blk[coo.row] == blk[coo.col]) & off
    print(f"{name}: blocks {nb}, avg size {Fn.shape[0]/nb:.1f}, intra frac {intra.sum()/off.sum():.3f}")
    # block graph
    BG = sp.csr_matrix((np.ones(off.sum()), (blk[coo.row[off]], blk[coo.col[off]])), shape=(nb, nb)); BG.sum_duplicates()
    BG.setdiag(0); BG.eliminate_zeros()
    deg = np.diff(BG.indptr)
    colb = greedy_colour(BG.indptr, BG.indices, nb)
    print(f"   block degree avg {deg.mean():.1f} max {deg.max()}, colours {colb.max()+1}")
    # external gather distinct (node, block) pairs: for each node j, number of distinct other blocks among its neighbours
    ext = off & ~intra
    pairs = np.unique(coo.col[ext].astype(np.int64) * nb + blk[coo.row[ext]])
    print(f"   distinct (node, gathering block) pairs per node {len(pairs)/Fn.shape[0]:.2f}; ext entries per row {ext.sum()/Fn.shape[0]:.1f}")
    return colb
s, nz = int(sys.argv[1]), int(sys.argv[2])
mesh = HostMesh.cylinder3d(s, nz)
prob = NavierStokes(mesh, "3d", T=1.0, deltat=DT, test_case=2)
prob.setup_host()
d = prob.dofs
num = dict(dim=3, cell_dofs=d.cell_dofs(), N=d.N, n_u=d.n_u, n_p=d.n_p, dpc=d.dpc)
pat = R.system_pattern(num)
rp, ci = pat[0], pat[1]
A = sp.csr_matrix((np.ones(len(ci)), ci, rp), shape=(d.N, d.N))
Fu = A[:d.n_u][:, :d.n_u].tocsr()
Fn = Fu[0::3][:, 0::3].tocsr(); Fn.sort_indices()
n = Fn.shape[0]
print("nodes", n, "nnz/row", Fn.nnz/n)
# natural consecutive blocks
blk_nat = np.arange(n)//32
stats("natural-32", Fn, blk_nat, (n+31)//32)
t0=time.time()
blk, order, nb = bfs_blocks(Fn.indptr, Fn.indices, n, 32)
print("bfs time", time.time()-t0)
colb = stats("bfs-32", Fn, blk, nb)
# morton
xyz = d.node_xyz
lo, hi = xyz.min(0), xyz.max(0)
h = 0.41/ (2**10)
q = np.minimum(((xyz-lo)/ (hi-lo).max() * (2**16-1)).astype(np.uint64), 2**16-1)
def spread(v):
    v = v.astype(np.uint64)
    out = np.zeros_like(v)
    for b in range(16): out |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(3*b)
    return out
code = spread(q[:,0]) | (spread(q[:,1])<<np.uint64(1)) | (spread(q[:,2])<<np.uint64(2))
om = np.argsort(code, kind="stable")
blk_m = np.empty(n, np.int64); blk_m[om] = np.arange(n)//32
stats("morton-32", Fn, blk_m, (n+31)//32)
if len(sys.argv) > 3:
    # convergence with ABMC ordering (bfs blocks)
    # order: by (block colour, block id, position in bfs order)
    pos = np.empty(n, np.int64); pos[order] = np.arange(n)
    key = np.lexsort((pos, blk, colb[blk]))
    order_n = key
    order_u = (3*order_n[:,None] + np.arange(3)[None,:]).ravel().astype(np.int32)
    o = R.Oracle(3, "3d", mesh.vertices, mesh.cells, num, pat, 1e-3, DT)
    o.set_ilu_order(order_u, None)
    o.set_dirichlet(prob._dir_rows, prob.dirichlet_values(DT))
    o.set_solution(np.zeros(d.N))
    o.assemble_first(); rc,k,_ = o.solve_step("yosida")
    print(f"abmc: first {k} outer, F {o.stat('n_inner_F')}/{o.stat('n_F_solves')}={o.stat('n_inner_F')/o.stat('n_F_solves'):.1f} S {o.stat('n_inner_S')}/{o.stat('n_S_solves')}")
    for _ in range(2):
        o.assemble_step(); rc,k,_ = o.solve_step("yosida")
        print(f"   step {k} outer, F {o.stat('n_inner_F')}/{o.stat('n_F_solves')}={o.stat('n_inner_F')/o.stat('n_F_solves'):.1f} S {o.stat('n_inner_S')/o.stat('n_S_solves'):.1f}", flush=True)
