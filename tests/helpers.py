"""Shared test plumbing: build the same problem for the CPU oracle and the CUDA engine."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from navierstokes_project_nm4pde_b200 import Engine, HostDofs, HostMesh, NavierStokes, gauss_simplex
from navierstokes_project_nm4pde_b200 import problem as P
from oracle import ns_ref as R

SEED = 20240607
VARIANT_PREC = {"2d": "asimple", "3d": "yosida", "conv": "yosida"}


def make_mesh(case):
    if case == "cyl2d":
        return HostMesh.cylinder2d(1), "2d"
    if case == "cyl3d":
        return HostMesh.cylinder3d(1, 3), "3d"
    if case == "box3d":
        return HostMesh.box(3, (4, 3, 3), (0.0, 0.0, 0.0), (0.8, 0.41, 0.41)), "3d"
    if case == "box2d":
        return HostMesh.box(2, (7, 5), (0.0, 0.0), (1.1, 0.41)), "2d"
    if case == "cube":
        return HostMesh.cube(3), "conv"
    raise KeyError(case)


class Case:
    """Host description shared by both sides."""

    def __init__(self, case, deltat=None, rule="wv"):
        self.mesh, self.variant = make_mesh(case)
        self.dim = self.mesh.dim
        self.dofs = HostDofs(self.mesh)
        self.rule = rule
        self.nu = 1e-2 if self.variant == "conv" else 1e-3
        self.dt = deltat if deltat is not None else {"2d": 0.01, "3d": 0.0002, "conv": 0.0004}[self.variant]
        self.prob = NavierStokes(self.mesh, self.variant, T=1.0, deltat=self.dt)
        self.N, self.n_u, self.n_p = self.dofs.N, self.dofs.n_u, self.dofs.n_p

    def bc(self, time):
        """Dirichlet rows/values through the product's host logic (needs no GPU)."""
        if not hasattr(self.prob, "dofs"):
            self.prob.test_case = 3 if self.dim == 2 else 2  # steady profiles: non-zero at any time
            self.prob.setup_host()
        return self.prob._dir_rows, self.prob.dirichlet_values(time)

    def neumann(self, time):
        full = np.zeros(self.N)
        full[: self.n_u] = self.prob.neumann_rhs(time)
        return full

    def initial(self):
        if not hasattr(self.prob, "dofs"):
            self.bc(0.0)
        return self.prob.initial_condition()

    def random_state(self, scale=1.0):
        rng = np.random.default_rng(SEED)
        return scale * rng.uniform(-1.0, 1.0, self.N)

    def oracle(self):
        num = R.number_dofs(self.dim, self.mesh.vertices, self.mesh.cells)
        assert np.array_equal(num["cell_dofs"], self.dofs.cell_dofs())
        pat = R.system_pattern(num)
        o = R.Oracle(self.dim, self.variant, self.mesh.vertices, self.mesh.cells, num, pat, self.nu, self.dt, self.rule)
        o.pattern4 = pat
        return o

    def engine(self, **params):
        e = Engine(self.dim)
        e.default_params(self.variant)
        e.set_mesh(self.dofs.cell_coords(), self.dofs.cell_dofs(), self.n_u, self.n_p)
        e.set_quadrature(*gauss_simplex(self.dim, self.rule))
        e.set_params(deltat=self.dt, nu=self.nu, **params)
        e.finalize()
        return e


def oracle_blocks(o, name):
    """Split the oracle's N x N CSR into the reference's blocks."""
    A = o.matrix(name)
    nu = o.n_u
    return {"F": A[:nu, :nu].tocsr(), "Bt": A[:nu, nu:].tocsr(), "B": A[nu:, :nu].tocsr()}


def same_pattern(A, rp, ci):
    A = A.tocsr()
    A.sort_indices()
    return np.array_equal(A.indptr, rp) and np.array_equal(A.indices, ci)


def entry_error(A, B):
    """max |a-b| / scale over stored entries, scale = max(|a|, |b|, ||row||_inf, ||column||_inf)
    (SURVEY.md H7 metric).  The column norm is included because on structured meshes whole rows
    of block (0,1) are cancellation zeros (~1e-18 sums of O(1e-2) cell contributions) that carry
    no significant digit in either implementation."""
    A, B = sp.csr_matrix(A), sp.csr_matrix(B)
    D = (A - B).tocoo()
    if D.nnz == 0:
        return 0.0
    rowmax = np.maximum(abs(A).max(axis=1).toarray().ravel(), abs(B).max(axis=1).toarray().ravel())
    colmax = np.maximum(abs(A).max(axis=0).toarray().ravel(), abs(B).max(axis=0).toarray().ravel())
    scale = np.maximum(rowmax[D.row], colmax[D.col])
    scale[scale == 0] = 1.0
    return float(np.max(np.abs(D.data) / scale))


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
