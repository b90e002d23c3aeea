import sys, time, os
sys.path.insert(0, "/root/repo")
import numpy as np, scipy.sparse as sp
from navierstokes_project_nm4pde_b200 import HostMesh, NavierStokes
from oracle import ns_ref as R
from collections import deque
DT=2e-4
def greedy_colour(rp, ci, n):
    col = -np.ones(n, np.int32)
    for i in range(n):
        used = set(col[ci[rp[i]:rp[i+1]]].tolist())
        c = 0
        while c in used: c += 1
        col[i] = c
    return col
def bfs_blocks(rp, ci, n, bs=32, fill=False):
    blk = -np.ones(n, np.int64); order = []
    nb = 0; cnt = 0; members = []
    for seed in range(n):
        if blk[seed] >= 0: continue
        if not fill or cnt == 0:
            members = []; cnt = 0
        q = deque([seed]); blk[seed] = nb; cnt += 1; members.append(seed)
        while q and cnt < bs:
            u = q.popleft()
            for v in ci[rp[u]:rp[u+1]]:
                if blk[v] < 0:
                    blk[v] = nb; members.append(v); q.append(v); cnt += 1
                    if cnt >= bs: break
        if cnt >= bs or not fill:
            order.extend(members); nb += 1; cnt = 0; members = []
    if fill and cnt > 0:
        order.extend(members); nb += 1
    return blk, np.array(order), nb
def stats(name, Fn, blk, nb, order):
    n = Fn.shape[0]
    coo = Fn.tocoo()
    off = coo.row != coo.col
    intra = (blk[coo.row] == blk[coo.col]) & off
    BG = sp.csr_matrix((np.ones(off.sum()), (blk[coo.row[off]], blk[coo.col[off]])), shape=(nb, nb)); BG.sum_duplicates()
    BG.setdiag(0); BG.eliminate_zeros()
    colb = greedy_colour(BG.indptr, BG.indices, nb)
    ext = off & ~intra
    pairs = np.unique(coo.col[ext].astype(np.int64) * nb + blk[coo.row[ext]])
    # per-block max intra-lower entries per row; levels of the full ordering
    pos = np.empty(n, np.int64); pos[order] = np.arange(n)
    key = np.lexsort((pos, blk, colb[blk]))   # new order: row k = node key[k]
    inv = np.empty(n, np.int64); inv[key] = np.arange(n)
    P = Fn[key][:, key].tocsr(); P.sort_indices()
    lvl = np.zeros(n, np.int64)
    rp, ci = P.indptr, P.indices
    for i in range(n):
        c = ci[rp[i]:rp[i+1]]; c = c[c < i]
        if len(c): lvl[i] = lvl[c].max() + 1
    blk_new = blk[key]
    # intra lower count per row
    cooP = P.tocoo(); low = cooP.col < cooP.row
    intraP = low & (blk_new[cooP.row] == blk_new[cooP.col])
    il = np.bincount(cooP.row[intraP], minlength=n)
    extl = np.bincount(cooP.row[low & ~intraP], minlength=n)
    sizes = np.bincount(blk, minlength=nb)
    print(f"{name}: blocks {nb} avg size {n/nb:.1f} (min {sizes.min()}), colours {colb.max()+1}, intra frac {intra.sum()/off.sum():.3f}, pairs/node {len(pairs)/n:.2f}, factor levels {lvl.max()+1}, intra-lower per row max {il.max()} mean {il.mean():.1f}, ext-lower per row max {extl.max()} mean {extl.mean():.1f}")
    csz = np.bincount(colb, minlength=colb.max()+1)
    print("   blocks per colour", csz.tolist())
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
for fill in (False, True):
    blk, order, nb = bfs_blocks(Fn.indptr, Fn.indices, n, 32, fill)
    stats(f"F bfs-32 fill={fill}", Fn, blk, nb, order)
B = A[d.n_u:][:, :d.n_u]; S = (B @ B.T).tocsr(); S.sort_indices()
print("S rows", S.shape[0], "nnz/row", S.nnz/S.shape[0])
for fill in (False, True):
    blk, order, nb = bfs_blocks(S.indptr, S.indices, S.shape[0], 32, fill)
    stats(f"S bfs-32 fill={fill}", S, blk, nb, order)

def padstats(name, M, fill=True):
    n = M.shape[0]
    blk, order, nb = bfs_blocks(M.indptr, M.indices, n, 32, fill)
    coo = M.tocoo(); off = coo.row != coo.col
    BG = sp.csr_matrix((np.ones(off.sum()), (blk[coo.row[off]], blk[coo.col[off]])), shape=(nb, nb)); BG.sum_duplicates(); BG.setdiag(0); BG.eliminate_zeros()
    colb = greedy_colour(BG.indptr, BG.indices, nb)
    pos = np.empty(n, np.int64); pos[order] = np.arange(n)
    key = np.lexsort((pos, blk, colb[blk]))
    P = M[key][:, key].tocoo()
    sl = np.arange(n)//32   # slices in new order == blocks (all full)
    low = P.col < P.row; up = P.col > P.row
    intra = sl[P.row] == sl[P.col]
    ns = (n+31)//32
    for nm, m in (("ext-lower", low & ~intra), ("ext-upper", up & ~intra), ("intra-lower", low & intra), ("intra-upper", up & intra)):
        cnt = np.bincount(P.row[m], minlength=ns*32).reshape(ns, 32)
        mx = cnt.max(1)
        print(f"{name} {nm}: entries {m.sum()}, padded slots {32*mx.sum()}, waste x{32*mx.sum()/m.sum():.2f}, max len {mx.max()}")
padstats("F", Fn)
padstats("S", S)
