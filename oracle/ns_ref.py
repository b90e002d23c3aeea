"""CPU ORACLE helpers -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy/ctypes side of the oracle: quadrature tables, a small independent restatement of the
reference's `setup()` (DoF numbering + sparsity, `/root/reference/Navier-Stokes/src/
NavierStokes2D.cpp:58-156`), boundary data (`NavierStokes2D.hpp:18-81`, `NavierStokes3D.hpp:18-81`,
`Convergence3D.hpp:51-201`) and a ctypes wrapper around `ns_oracle.c`.

PARITY UNPINNED: see the header of ns_oracle.c.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SEED = 20240607

# ---------------------------------------------------------------------------------------------
# quadrature: QGaussSimplex<dim>(3)  [deal.II, restated; version dependent -- SURVEY.md H3]
# ---------------------------------------------------------------------------------------------


def _perm3(a):
    b = 1.0 - 2.0 * a
    return [(a, a), (b, a), (a, b)]


def quadrature(dim: int, rule: str = "wv"):
    """Return (xi[nq, dim], w[nq]).

    rule "wv"       : deal.II >= 9.4, QWitherdenVincentSimplex degree 5 (7 points in 2D, 14 in 3D)
    rule "dealii93" : deal.II 9.3.x hard-coded tables (7 points 2D with truncated constants,
                      10 points 3D, degree 3)
    `Convergence3D.cpp:772` needs QGaussSimplex<3>(4), which 9.3 does not implement, so the
    reference must have been run with >= 9.4: "wv" is the default everywhere in this repo.
    """
    if dim == 2 and rule == "wv":
        s15 = math.sqrt(15.0)
        a, b = (6.0 - s15) / 21.0, (6.0 + s15) / 21.0
        wa, wb = (155.0 - s15) / 2400.0, (155.0 + s15) / 2400.0
        pts = [(1.0 / 3.0, 1.0 / 3.0)] + _perm3(a) + _perm3(b)
        w = [9.0 / 80.0] + [wa] * 3 + [wb] * 3
        return np.array(pts), np.array(w)
    if dim == 2 and rule == "dealii93":
        pts = [(0.3333333333330, 0.3333333333330), (0.7974269853530, 0.1012865073230),
               (0.1012865073230, 0.7974269853530), (0.1012865073230, 0.1012865073230),
               (0.0597158717898, 0.4701420641050), (0.4701420641050, 0.0597158717898),
               (0.4701420641050, 0.4701420641050)]
        w = [0.5 * 0.225] + [0.5 * 0.125939180545] * 3 + [0.5 * 0.132394152789] * 3
        return np.array(pts), np.array(w)
    if dim == 3 and rule == "wv":
        a1, w1 = 0.31088591926330060980, 0.11268792571801585080 / 6.0
        a2, w2 = 0.092735250310891226402, 0.073493043116361949544 / 6.0
        c, w3 = 0.045503704125649649492, 0.042546020777081466438 / 6.0
        d = 0.5 - c
        pts, w = [], []
        for a, ww in ((a1, w1), (a2, w2)):
            b = 1.0 - 3.0 * a
            pts += [(a, a, a), (b, a, a), (a, b, a), (a, a, b)]
            w += [ww] * 4
        pts += [(c, c, d), (c, d, c), (d, c, c), (c, d, d), (d, c, d), (d, d, c)]
        w += [w3] * 6
        return np.array(pts), np.array(w)
    if dim == 3 and rule == "dealii93":
        a, b = 0.5684305841968444, 0.1438564719343852
        pts = [(a, b, b), (b, b, b), (b, b, a), (b, a, b),
               (0.0, 0.5, 0.5), (0.5, 0.0, 0.5), (0.5, 0.5, 0.0),
               (0.5, 0.0, 0.0), (0.0, 0.5, 0.0), (0.0, 0.0, 0.5)]
        w = [0.2177650698804054 / 6.0] * 4 + [0.0214899534130631 / 6.0] * 6
        return np.array(pts), np.array(w)
    if dim == 1:  # QGauss<1>(3) on [0,1]
        g = math.sqrt(3.0 / 5.0)
        return (np.array([[0.5 - 0.5 * g], [0.5], [0.5 + 0.5 * g]]), np.array([5.0, 8.0, 5.0]) / 18.0)
    raise ValueError((dim, rule))


# ---------------------------------------------------------------------------------------------
# tiny structured simplex meshes (independent of the product's generators)
# ---------------------------------------------------------------------------------------------


def _fix_orientation(verts, cells):
    dim = verts.shape[1]
    v = verts[cells]
    J = np.stack([v[:, k + 1] - v[:, 0] for k in range(dim)], axis=2)
    neg = np.linalg.det(J) < 0
    cells = cells.copy()
    cells[neg, 0], cells[neg, 1] = cells[neg, 1].copy(), cells[neg, 0].copy()
    return cells


def square_mesh(n: int, lo=(0.0, 0.0), hi=(1.0, 1.0), jitter: float = 0.0):
    xs = np.linspace(lo[0], hi[0], n + 1)
    ys = np.linspace(lo[1], hi[1], n + 1)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    verts = np.stack([X.ravel(), Y.ravel()], axis=1)
    if jitter:
        rng = np.random.default_rng(SEED)
        h = (hi[0] - lo[0]) / n
        inner = np.ones((n + 1, n + 1), bool)
        inner[0, :] = inner[-1, :] = inner[:, 0] = inner[:, -1] = False
        verts[inner.ravel()] += rng.uniform(-jitter * h, jitter * h, size=(int(inner.sum()), 2))
    idx = lambda i, j: i * (n + 1) + j
    cells = []
    for i in range(n):
        for j in range(n):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            cells += [(a, b, c), (a, c, d)]
    return verts, _fix_orientation(verts, np.array(cells, dtype=np.int32))


_KUHN = [(0, 1, 3, 7), (0, 1, 5, 7), (0, 2, 3, 7), (0, 2, 6, 7), (0, 4, 5, 7), (0, 4, 6, 7)]


def cube_mesh(n: int, lo=-1.0, hi=1.0, jitter: float = 0.0):
    xs = np.linspace(lo, hi, n + 1)
    X, Y, Z = np.meshgrid(xs, xs, xs, indexing="ij")
    verts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    if jitter:
        rng = np.random.default_rng(SEED)
        h = (hi - lo) / n
        inner = np.ones((n + 1,) * 3, bool)
        inner[0] = inner[-1] = False
        inner[:, 0] = inner[:, -1] = False
        inner[:, :, 0] = inner[:, :, -1] = False
        verts[inner.ravel()] += rng.uniform(-jitter * h, jitter * h, size=(int(inner.sum()), 3))
    idx = lambda i, j, k: (i * (n + 1) + j) * (n + 1) + k
    cells = []
    for i in range(n):
        for j in range(n):
            for k in range(n):
                c = [idx(i + (b & 1), j + ((b >> 1) & 1), k + ((b >> 2) & 1)) for b in range(8)]
                cells += [tuple(c[t] for t in tet) for tet in _KUHN]
    return verts, _fix_orientation(verts, np.array(cells, dtype=np.int32))


# ---------------------------------------------------------------------------------------------
# DoF numbering [deal.II distribute_dofs + DoFRenumbering::component_wise, restated]
# ---------------------------------------------------------------------------------------------
EDGES = {2: [(0, 1), (1, 2), (2, 0)], 3: [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]}


def number_dofs(dim: int, verts: np.ndarray, cells: np.ndarray):
    """Walk cells in order; per cell number un-numbered vertices ([u.., p]) then un-numbered
    edges ([u..]); then move velocity DoFs first keeping relative order.  Returns a dict with
    the compact node numbering and the reference-layout `cell_dofs`."""
    nv1 = dim + 1
    edges = EDGES[dim]
    n2 = nv1 + len(edges)
    dpc = dim * n2 + nv1
    node_of_vertex, node_of_edge, p_of_vertex = {}, {}, {}
    node_xyz = []
    nc = cells.shape[0]
    cell_nodes = np.zeros((nc, n2), np.int32)
    cell_p = np.zeros((nc, nv1), np.int32)
    for c in range(nc):
        vs = cells[c]
        for lv in range(nv1):
            g = int(vs[lv])
            if g not in node_of_vertex:
                node_of_vertex[g] = len(node_xyz)
                p_of_vertex[g] = len(p_of_vertex)
                node_xyz.append(verts[g])
            cell_nodes[c, lv] = node_of_vertex[g]
            cell_p[c, lv] = p_of_vertex[g]
        for le, (a, b) in enumerate(edges):
            ga, gb = int(vs[a]), int(vs[b])
            key = (min(ga, gb), max(ga, gb))
            if key not in node_of_edge:
                node_of_edge[key] = len(node_xyz)
                node_xyz.append(0.5 * (verts[ga] + verts[gb]))
            cell_nodes[c, nv1 + le] = node_of_edge[key]
    n_nodes = len(node_xyz)
    n_u, n_p = dim * n_nodes, len(p_of_vertex)
    cell_dofs = np.zeros((nc, dpc), np.int32)
    for lv in range(nv1):
        for comp in range(dim):
            cell_dofs[:, lv * (dim + 1) + comp] = dim * cell_nodes[:, lv] + comp
        cell_dofs[:, lv * (dim + 1) + dim] = n_u + cell_p[:, lv]
    for le in range(len(edges)):
        for comp in range(dim):
            cell_dofs[:, nv1 * (dim + 1) + le * dim + comp] = dim * cell_nodes[:, nv1 + le] + comp
    p_vertex = np.zeros(n_p, np.int64)
    for g, k in p_of_vertex.items():
        p_vertex[k] = g
    return dict(dim=dim, n_nodes=n_nodes, n_u=n_u, n_p=n_p, N=n_u + n_p, n2=n2, dpc=dpc,
                cell_nodes=cell_nodes, cell_p=cell_p, cell_dofs=cell_dofs,
                node_xyz=np.array(node_xyz), p_xyz=verts[p_vertex],
                node_of_vertex=node_of_vertex, node_of_edge=node_of_edge)


def local_dof_table(dim: int):
    """(comp, base) of each FESystem local DoF; mirrors local_dof() in ns_oracle.c."""
    nv1 = dim + 1
    n2 = nv1 + len(EDGES[dim])
    dpc = dim * n2 + nv1
    comp, base = np.zeros(dpc, int), np.zeros(dpc, int)
    for i in range(dpc):
        if i < nv1 * (dim + 1):
            comp[i], base[i] = i % (dim + 1), i // (dim + 1)
        else:
            r = i - nv1 * (dim + 1)
            comp[i], base[i] = r % dim, nv1 + r // dim
    return comp, base


def system_pattern(num):
    """DoFTools::make_sparsity_pattern with coupling 'always' except p-p 'none'
    (NavierStokes2D.cpp:109-124); plus the p-p pattern of pressure_mass (:127-142)."""
    import scipy.sparse as sp

    dim, cd, N, n_u = num["dim"], num["cell_dofs"], num["N"], num["n_u"]
    comp, _ = local_dof_table(dim)
    dpc = num["dpc"]
    ii, jj = np.meshgrid(np.arange(dpc), np.arange(dpc), indexing="ij")
    keep = ~((comp[ii] == dim) & (comp[jj] == dim))
    rows = cd[:, ii[keep]].ravel()
    cols = cd[:, jj[keep]].ravel()
    A = sp.csr_matrix((np.ones(rows.size, np.int8), (rows, cols)), shape=(N, N))
    A.sum_duplicates()
    A.sort_indices()
    pp = (comp[ii] == dim) & (comp[jj] == dim)
    rows = cd[:, ii[pp]].ravel() - n_u
    cols = cd[:, jj[pp]].ravel() - n_u
    M = sp.csr_matrix((np.ones(rows.size, np.int8), (rows, cols)), shape=(num["n_p"], num["n_p"]))
    M.sum_duplicates()
    M.sort_indices()
    return (A.indptr.astype(np.int32), A.indices.astype(np.int32),
            M.indptr.astype(np.int32), M.indices.astype(np.int32))


def boundary_faces(dim: int, cells: np.ndarray):
    """Faces that belong to exactly one cell -> (cell, local_face, vertex ids[dim])."""
    from collections import defaultdict

    nv1 = dim + 1
    cnt = defaultdict(list)
    for c in range(cells.shape[0]):
        for f in range(nv1):  # face opposite to local vertex f
            vs = tuple(sorted(int(cells[c, v]) for v in range(nv1) if v != f))
            cnt[vs].append((c, f))
    out = []
    for vs, lst in cnt.items():
        if len(lst) == 1:
            out.append((lst[0][0], lst[0][1], vs))
    out.sort()
    return out


def dirichlet_nodes(num, bfaces, face_ids, wanted_ids):
    """P2 nodes lying on boundary faces whose id is in `wanted_ids`, in first-visit order."""
    dim = num["dim"]
    seen, order = set(), []
    for (c, f, vs), fid in zip(bfaces, face_ids):
        if fid not in wanted_ids:
            continue
        nodes = [num["node_of_vertex"][v] for v in vs]
        for a in range(len(vs)):
            for b in range(a + 1, len(vs)):
                nodes.append(num["node_of_edge"][(min(vs[a], vs[b]), max(vs[a], vs[b]))])
        for nd in nodes:
            if nd not in seen:
                seen.add(nd)
                order.append(nd)
    return np.array(order, np.int64)


# ---------------------------------------------------------------------------------------------
# boundary / initial data of the three drivers
# ---------------------------------------------------------------------------------------------


def inlet_2d(xyz, t, test_case=2, H=0.41, u_m=1.5):
    """NavierStokes2D.hpp:26-44"""
    y = xyz[:, 1]
    v = np.zeros_like(xyz)
    if test_case == 2:
        v[:, 0] = 4.0 * u_m * y * (H - y) * math.sin(math.pi * t / 8.0) / (H * H)
    elif test_case != 1:
        v[:, 0] = 4.0 * u_m * y * (H - y) / (H * H)
    return v


def inlet_3d(xyz, t, test_case=2, H=0.41, u_m=9.0):
    """NavierStokes3D.hpp:26-43"""
    y, z = xyz[:, 1], xyz[:, 2]
    v = np.zeros_like(xyz)
    if test_case == 3:
        v[:, 0] = 16.0 * u_m * y * z * (H - z) * (H - y) * math.sin(math.pi * t / 8.0) / (H ** 4)
    elif test_case != 1:
        v[:, 0] = 16.0 * u_m * y * z * (H - z) * (H - y) / (H ** 4)
    return v


ES_NU, ES_A, ES_B = 1e-2, math.pi / 4.0, math.pi / 2.0


def ethier_steinman(xyz, t):
    """Convergence3D.hpp:59-73 -> (u[n,3], p[n])"""
    a, b, nu = ES_A, ES_B, ES_NU
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    e = math.exp(-nu * b * b * t)
    u = np.stack([
        -a * e * (np.exp(a * x) * np.sin(a * y + b * z) + np.exp(a * z) * np.cos(a * x + b * y)),
        -a * e * (np.exp(a * y) * np.sin(a * z + b * x) + np.exp(a * x) * np.cos(a * y + b * z)),
        -a * e * (np.exp(a * z) * np.sin(a * x + b * y) + np.exp(a * y) * np.cos(a * z + b * x)),
    ], axis=1)
    factor = -(a * a * math.exp(-2 * nu * b * b * t)) / 2.0
    t1 = 2.0 * np.sin(a * x + b * y) * np.cos(a * z + b * x) * np.exp(a * (y + z))
    t2 = 2.0 * np.sin(a * y + b * z) * np.cos(a * x + b * y) * np.exp(a * (x + z))
    t3 = 2.0 * np.sin(a * z + b * x) * np.cos(a * y + b * z) * np.exp(a * (x + y))
    t4 = np.exp(2 * a * x) + np.exp(2 * a * y) + np.exp(2 * a * z)
    return u, factor * (t1 + t2 + t3 + t4)


def function_h(xyz, t):
    """Convergence3D.hpp:159-175 (Neumann datum on boundary id 3)"""
    a, b, nu = ES_A, ES_B, ES_NU
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    _, p = ethier_steinman(xyz, t)
    e = math.exp(-nu * b * b * t)
    h0 = -nu * a * e * (a * np.exp(a * x) * np.cos(a * y + b * z) - b * np.exp(a * z) * np.sin(a * x + b * y))
    h1 = -nu * a * e * (a * np.exp(a * y) * np.sin(a * z + b * x) - a * np.exp(a * x) * np.sin(a * y + b * z)) - p
    h2 = -nu * a * e * (b * np.exp(a * z) * np.cos(a * x + b * y) + a * np.exp(a * y) * np.cos(a * z + b * x))
    return np.stack([h0, h1, h2], axis=1)


def neumann_rhs(num, verts, bfaces, face_ids, t, neumann_id=3):
    """Face integral of Convergence3D.cpp:309-330 with QGaussSimplex<2>(3) on P2 face traces."""
    dim = num["dim"]
    assert dim == 3
    xi, w = quadrature(2, "wv")
    lam = np.stack([1 - xi[:, 0] - xi[:, 1], xi[:, 0], xi[:, 1]], axis=1)  # [nq,3]
    out = np.zeros(num["N"])
    for (c, f, vs), fid in zip(bfaces, face_ids):
        if fid != neumann_id:
            continue
        P = verts[list(vs)]
        area2 = np.linalg.norm(np.cross(P[1] - P[0], P[2] - P[0]))  # = 2*area = |J| of the face map
        xq = lam @ P
        h = function_h(xq, t)
        nodes = [num["node_of_vertex"][v] for v in vs]
        shp = [lam[:, k] * (2 * lam[:, k] - 1) for k in range(3)]
        for a_, b_ in ((0, 1), (0, 2), (1, 2)):
            nodes.append(num["node_of_edge"][(min(vs[a_], vs[b_]), max(vs[a_], vs[b_]))])
            shp.append(4 * lam[:, a_] * lam[:, b_])
        for nd, s in zip(nodes, shp):
            for comp in range(3):
                out[3 * nd + comp] += float(np.sum(h[:, comp] * s * w) * area2)
    return out


# ---------------------------------------------------------------------------------------------
# ctypes wrapper
# ---------------------------------------------------------------------------------------------
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libns_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.nso_create.restype = C.c_void_p
        L.nso_ptr.restype = C.POINTER(C.c_double)
        L.nso_iptr.restype = C.POINTER(C.c_int)
        L.nso_stat.restype = C.c_long
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


VARIANT = {"2d": 0, "3d": 1, "conv": 2}
PRECOND = {"yosida": 0, "simple": 1, "ayosida": 2, "asimple": 3}


class Oracle:
    """State of one `NavierStokes` instance of the reference, CPU restatement."""

    def __init__(self, dim, variant, verts, cells, num, pattern, nu, dt, rule="wv"):
        L = lib()
        self.L, self.num, self.dim = L, num, dim
        self.N, self.n_u, self.n_p = num["N"], num["n_u"], num["n_p"]
        vc = np.ascontiguousarray(verts[cells], dtype=np.float64)
        cd = np.ascontiguousarray(num["cell_dofs"], dtype=np.int32)
        self.rowptr, self.colind, pm_rp, pm_ci = [np.ascontiguousarray(a, dtype=np.int32) for a in pattern]
        xi, w = quadrature(dim, rule)
        xi, w = np.ascontiguousarray(xi), np.ascontiguousarray(w)
        self.h = C.c_void_p(L.nso_create(dim, VARIANT[variant], cells.shape[0], _dp(vc), _ip(cd), self.N, self.n_u,
                                         _ip(self.rowptr), _ip(self.colind), len(w), _dp(xi), _dp(w),
                                         C.c_double(nu), C.c_double(dt)))
        L.nso_set_pressure_mass_pattern(self.h, _ip(pm_rp), _ip(pm_ci))
        self.nnz = int(self.rowptr[-1])
        self.pm_nnz = int(pm_rp[-1])

    def close(self):
        if self.h:
            self.L.nso_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_options(self, dirichlet_mode=0, gmres_tmp=0, outer_tol=0.0, outer_maxit=0, inner_rtol=0.0, inner_maxit=0):
        self.L.nso_set_options(self.h, dirichlet_mode, gmres_tmp, C.c_double(outer_tol), outer_maxit,
                               C.c_double(inner_rtol), inner_maxit)

    def set_orthogonalisation(self, mode):
        """0: modified Gram-Schmidt (deal.II); 1: batched classical Gram-Schmidt (engine throughput mode)."""
        self.L.nso_set_orthogonalisation(self.h, int(mode))

    def set_dirichlet(self, rows, vals):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.L.nso_set_dirichlet(self.h, len(rows), _ip(rows), _dp(vals))

    def set_dirichlet_values(self, vals):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        self.L.nso_set_dirichlet_values(self.h, _dp(vals))

    def set_neumann_rhs(self, add):
        add = np.ascontiguousarray(add, dtype=np.float64)
        self.L.nso_set_neumann_rhs(self.h, _dp(add))

    def set_solution(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        self.L.nso_set_solution(self.h, _dp(x))

    def set_partition(self, part):
        part = np.ascontiguousarray(part, dtype=np.int32)
        self.L.nso_set_partition(self.h, _ip(part))

    def set_ilu_order(self, order_u=None, order_p=None):
        ou = None if order_u is None else np.ascontiguousarray(order_u, dtype=np.int32)
        op = None if order_p is None else np.ascontiguousarray(order_p, dtype=np.int32)
        self.L.nso_set_ilu_order(self.h, None if ou is None else _ip(ou), None if op is None else _ip(op))

    def assemble_first(self):
        self.L.nso_assemble_first(self.h)

    def assemble_step(self):
        self.L.nso_assemble_step(self.h)

    def array(self, name, n):
        p = self.L.nso_ptr(self.h, name.encode())
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def matrix(self, name):
        import scipy.sparse as sp

        v = self.array(name, self.nnz)
        return sp.csr_matrix((v, self.colind, self.rowptr), shape=(self.N, self.N))

    def precond_init(self, ptype):
        self.L.nso_precond_init(self.h, PRECOND[ptype])

    def precond_vmult(self, ptype, src, dst0=None):
        src = np.ascontiguousarray(src, dtype=np.float64)
        dst = np.zeros(self.N) if dst0 is None else np.ascontiguousarray(dst0, dtype=np.float64).copy()
        self.L.nso_precond_vmult(self.h, PRECOND[ptype], _dp(src), _dp(dst))
        return dst

    def system_vmult(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.N)
        self.L.nso_system_vmult(self.h, _dp(x), _dp(y))
        return y

    def block_vmult(self, which, x, n_out):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(n_out)
        self.L.nso_block_vmult(self.h, which, _dp(x), _dp(y))
        return y

    def ilu_apply(self, which, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        self.L.nso_ilu_apply(self.h, which, _dp(x), _dp(y))
        return y

    def schur(self):
        import scipy.sparse as sp

        nnz = int(self.L.nso_stat(self.h, b"S_nnz"))
        rp = np.ctypeslib.as_array(self.L.nso_iptr(self.h, b"S_rowptr"), shape=(self.n_p + 1,)).copy()
        ci = np.ctypeslib.as_array(self.L.nso_iptr(self.h, b"S_colind"), shape=(nnz,)).copy()
        v = self.array("S_val", nnz)
        return sp.csr_matrix((v, ci, rp), shape=(self.n_p, self.n_p))

    def compute_forces(self, solution, face_cell, face_opp, xi_f, w_f, rho=1.0):
        """(drag, lift) face integrals of NavierStokes::compute_forces (NavierStokes2D.cpp:752-859,
        NavierStokes3D.cpp:744-840) of `solution` over the given faces (cell, opposite local vertex)."""
        sol = np.ascontiguousarray(solution, dtype=np.float64)
        fc = np.ascontiguousarray(face_cell, dtype=np.int32)
        fo = np.ascontiguousarray(face_opp, dtype=np.int32)
        xi = np.ascontiguousarray(xi_f, dtype=np.float64)
        w = np.ascontiguousarray(w_f, dtype=np.float64)
        out = np.zeros(2)
        self.L.nso_compute_forces(self.h, _dp(sol), len(fc), _ip(fc), _ip(fo), len(w), _dp(xi), _dp(w), C.c_double(rho), _dp(out))
        return out

    def solve_step(self, ptype):
        its, res = C.c_int(0), C.c_double(0.0)
        rc = self.L.nso_solve_step(self.h, PRECOND[ptype], C.byref(its), C.byref(res))
        return rc, its.value, res.value

    def stat(self, name):
        return int(self.L.nso_stat(self.h, name.encode()))

    def residual_history(self):
        n = self.stat("n_res_hist")
        return self.array("res_hist", n)
