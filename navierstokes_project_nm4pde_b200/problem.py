"""Host-side mirror of the reference's `NavierStokes` class for the hot path.

Same public surface as the reference (`setup()`, `solve()`, the public timing vectors;
`Navier-Stokes/include/NavierStokes2D.hpp:105-119`) and the same protected methods
(`assemble`, `assemble_time_step`, `solve_time_step`; `NavierStokes2D.hpp:123-131`), each of which
is one call through the C ABI.  The three reference copies of the class (2D cylinder, 3D cylinder,
Ethier-Steinman convergence) are the `variant` argument.  Boundary and initial data restate
`NavierStokes2D.hpp:18-81`, `NavierStokes3D.hpp:18-81` and `Convergence3D.hpp:51-265`.
"""
from __future__ import annotations

import math

import numpy as np

from .engine import Engine, HostDofs, HostMesh
from .quadrature import gauss_simplex

_ES_A, _ES_B, _ES_NU = math.pi / 4.0, math.pi / 2.0, 1e-2


def inlet_velocity(dim, xyz, t, test_case):
    """InletVelocity::vector_value (NavierStokes2D.hpp:26-44, NavierStokes3D.hpp:25-43)."""
    H = 0.41
    v = np.zeros((xyz.shape[0], dim))
    if dim == 2:
        u_m, y = 1.5, xyz[:, 1]
        if test_case == 2:
            v[:, 0] = 4.0 * u_m * y * (H - y) * math.sin(math.pi * t / 8.0) / (H * H)
        elif test_case != 1:
            v[:, 0] = 4.0 * u_m * y * (H - y) / (H * H)
    else:
        u_m, y, z = 9.0, xyz[:, 1], xyz[:, 2]
        if test_case == 3:
            v[:, 0] = 16.0 * u_m * y * z * (H - z) * (H - y) * math.sin(math.pi * t / 8.0) / (H * H * H * H)
        elif test_case != 1:
            v[:, 0] = 16.0 * u_m * y * z * (H - z) * (H - y) / (H * H * H * H)
    return v


def mean_velocity(dim, t, test_case):
    """InletVelocity::getMeanVelocity (NavierStokes2D.hpp:64-75, NavierStokes3D.hpp:64-75).
    Note the reference swaps cases 2/3 between the profile and the mean in 2D; reproduced."""
    if test_case == 1:
        return 0.0
    if dim == 2:
        return 2.0 * 1.5 * math.sin(t * math.pi / 8.0) / 3.0 if test_case == 3 else 2.0 * 1.5 / 3.0
    return 4.0 * 9.0 * math.sin(t * math.pi / 8.0) / 9.0 if test_case == 3 else 4.0 * 9.0 / 9.0


def exact_solution(xyz, t):
    """ExactSolution::vector_value (Convergence3D.hpp:59-73): Ethier-Steinman, nu = 1e-2."""
    a, b, nu = _ES_A, _ES_B, _ES_NU
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    e = math.exp(-nu * b * b * t)
    u = np.stack([
        -a * e * (np.exp(a * x) * np.sin(a * y + b * z) + np.exp(a * z) * np.cos(a * x + b * y)),
        -a * e * (np.exp(a * y) * np.sin(a * z + b * x) + np.exp(a * x) * np.cos(a * y + b * z)),
        -a * e * (np.exp(a * z) * np.sin(a * x + b * y) + np.exp(a * y) * np.cos(a * z + b * x))], axis=1)
    f = -(a * a * math.exp(-2 * nu * b * b * t)) / 2.0
    p = f * (2.0 * np.sin(a * x + b * y) * np.cos(a * z + b * x) * np.exp(a * (y + z))
             + 2.0 * np.sin(a * y + b * z) * np.cos(a * x + b * y) * np.exp(a * (x + z))
             + 2.0 * np.sin(a * z + b * x) * np.cos(a * y + b * z) * np.exp(a * (x + y))
             + np.exp(2 * a * x) + np.exp(2 * a * y) + np.exp(2 * a * z))
    return u, p


def exact_gradient(xyz, t):
    """ExactSolution::gradient_tensor (Convergence3D.hpp:109-132): grad[n, i, j] = d u_i / d x_j."""
    a, b, nu = _ES_A, _ES_B, _ES_NU
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    e = -a * math.exp(-nu * b * b * t)
    ex, ey, ez = np.exp(a * x), np.exp(a * y), np.exp(a * z)
    g = np.zeros((xyz.shape[0], 3, 3))
    g[:, 0, 0] = e * (a * ex * np.sin(a * y + b * z) - a * ez * np.sin(a * x + b * y))
    g[:, 0, 1] = e * (a * ex * np.cos(a * y + b * z) - b * ez * np.sin(a * x + b * y))
    g[:, 0, 2] = e * (b * ex * np.cos(a * y + b * z) + a * ez * np.cos(a * x + b * y))
    g[:, 1, 0] = e * (b * ey * np.cos(a * z + b * x) + a * ex * np.cos(a * y + b * z))
    g[:, 1, 1] = e * (a * ey * np.sin(a * z + b * x) - a * ex * np.sin(a * y + b * z))
    g[:, 1, 2] = e * (a * ey * np.cos(a * z + b * x) - b * ex * np.sin(a * y + b * z))
    g[:, 2, 0] = e * (a * ez * np.cos(a * x + b * y) - b * ey * np.sin(a * z + b * x))
    g[:, 2, 1] = e * (b * ez * np.cos(a * x + b * y) + a * ey * np.cos(a * z + b * x))
    g[:, 2, 2] = e * (a * ez * np.sin(a * x + b * y) - a * ey * np.sin(a * z + b * x))
    return g


def function_h(xyz, t):
    """FunctionH::vector_value (Convergence3D.hpp:159-175): nu du/dn - p n with n = +e_y."""
    g = exact_gradient(xyz, t)
    _, p = exact_solution(xyz, t)
    h = _ES_NU * g[:, :, 1]
    h[:, 1] -= p
    return h


# local P2 nodes of the face opposite to local vertex f: vertices, then edges with both ends on the face
_EDGES = {2: [(0, 1), (1, 2), (2, 0)], 3: [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]}


def _face_nodes(dim, f):
    vs = [v for v in range(dim + 1) if v != f]
    es = [(dim + 1 + k, a, b) for k, (a, b) in enumerate(_EDGES[dim]) if a != f and b != f]
    return vs, es


class NavierStokes:
    """`NavierStokes problem(mesh, 2, 1, T, deltat, test_case); problem.setup(); problem.solve();`"""

    def __init__(self, mesh, variant="2d", T=8.0, deltat=0.01, test_case=2, device=0, rule="wv", nranks=1, rank=0,
                 unique_id=None, verbose=False, **param_overrides):
        self.mesh = HostMesh.read_msh(mesh) if isinstance(mesh, (str, bytes)) else mesh
        self.variant, self.T, self.deltat, self.test_case = variant, float(T), float(deltat), test_case
        self.dim = self.mesh.dim
        if (variant == "2d") != (self.dim == 2):
            raise ValueError("variant does not match the mesh dimension")
        self.device, self.rule, self.verbose = device, rule, verbose
        self.nranks, self.rank, self.unique_id = nranks, rank, unique_id
        self.param_overrides = param_overrides
        # public result vectors of the reference class (NavierStokes2D.hpp:113-119)
        self.vec_drag, self.vec_lift, self.vec_drag_coeff, self.vec_lift_coeff = [], [], [], []
        self.time_prec, self.time_solve, self.gmres_iterations = [], [], []
        self.engine = None
        self.forces_after, self.c_D_max, self.c_L_min, self.time = 0.1, -999.0, 999.0, 0.0

    # ------------------------------------------------------------------ setup()
    def setup_host(self):
        """Host-only part of setup(): DoF numbering and boundary lists (needs no GPU)."""
        self.dofs = d = HostDofs(self.mesh)
        self.n_u, self.n_p, self.N = d.n_u, d.n_p, d.N
        # Dirichlet rows (interpolate_boundary_values, NavierStokes2D.cpp:328-353)
        node_xyz = d.node_xyz
        if self.variant == "conv":
            self._dir_nodes = d.boundary_nodes([0, 1, 2, 4, 5])  # Convergence3D.cpp:364-368
            self._inlet_mask = None
        else:
            inlet = d.boundary_nodes([0])
            walls = d.boundary_nodes([2, 3])
            nodes = np.concatenate([inlet, walls[~np.isin(walls, inlet)]])
            self._dir_nodes = nodes
            self._inlet_mask = np.isin(nodes, inlet) & ~np.isin(nodes, walls)  # second call overwrites with zero
        self._dir_xyz = node_xyz[self._dir_nodes]
        rows = (self.dim * self._dir_nodes[:, None] + np.arange(self.dim)[None, :]).ravel()
        self._dir_rows = rows.astype(np.int32)
        if self.variant == "conv":
            self._prepare_neumann()
        self.solution = np.zeros(self.N)
        return self

    def setup(self):
        """NavierStokes::setup (NavierStokes2D.cpp:2-157): FE space, DoFs, sparsity, vectors."""
        self.setup_host()
        d = self.dofs
        e = self.engine = Engine(self.dim, self.device, self.nranks, self.rank, self.unique_id)
        e.default_params(self.variant)
        e.set_mesh(d.cell_coords(copy=False), d.cell_dofs(copy=False), d.n_u, d.n_p)
        e.set_quadrature(*gauss_simplex(self.dim, self.rule))
        e.set_params(deltat=self.deltat, **self.param_overrides)
        e.finalize()
        self.nu = e.params.nu
        e.set_dirichlet(self._dir_rows)
        if self.variant != "conv":  # obstacle faces of compute_forces (boundary id 3)
            e.set_force_faces(*d.boundary_faces(3), *gauss_simplex(self.dim - 1, self.rule))
        return self

    def dirichlet_values(self, time):
        if self.variant == "conv":
            u, _ = exact_solution(self._dir_xyz, time)
            return u.ravel()
        v = inlet_velocity(self.dim, self._dir_xyz, time, self.test_case)
        v[~self._inlet_mask] = 0.0
        return v.ravel()

    # Neumann face term of Convergence3D.cpp:309-330 (QGaussSimplex<2>(3) on faces with id 3)
    def _prepare_neumann(self):
        d, dim = self.dofs, self.dim
        fc, fl = d.boundary_faces(3)
        xi, w = gauss_simplex(2, self.rule)
        lam = np.stack([1.0 - xi[:, 0] - xi[:, 1], xi[:, 0], xi[:, 1]], axis=1)
        cc = d.cell_coords(copy=False)
        cd = d.cell_dofs(copy=False)
        nodes, shapes, xq, jw = [], [], [], []
        for c, f in zip(fc, fl):
            vs, es = _face_nodes(dim, f)
            P = cc[c][vs]
            area2 = np.linalg.norm(np.cross(P[1] - P[0], P[2] - P[0]))
            shp = [lam[:, k] * (2.0 * lam[:, k] - 1.0) for k in range(3)]
            nd = [cd[c][v * (dim + 1)] // dim for v in vs]
            for (ln, a, b) in es:
                shp.append(4.0 * lam[:, vs.index(a)] * lam[:, vs.index(b)])
                nd.append(cd[c][(dim + 1) * (dim + 1) + (ln - dim - 1) * dim] // dim)
            nodes.append(nd)
            shapes.append(np.stack(shp, axis=0) * (w * area2)[None, :])
            xq.append(lam @ P)
        self._neu = (np.array(nodes), np.array(shapes), np.array(xq)) if len(fc) else None

    def neumann_rhs(self, time):
        if self._neu is None:
            return np.zeros(self.dofs.n_u)
        nodes, shapes, xq = self._neu
        nf, nq = xq.shape[0], xq.shape[1]
        h = function_h(xq.reshape(-1, 3), time).reshape(nf, nq, 3)
        contrib = np.einsum("fnq,fqc->fnc", shapes, h)  # [face, node, comp]
        out = np.zeros((self.dofs.n_u // 3, 3))
        np.add.at(out, nodes.ravel(), contrib.reshape(-1, 3))
        return out.ravel()

    # ------------------------------------------------------------------ protected methods
    def assemble(self, time):
        """NavierStokes::assemble (NavierStokes2D.cpp:164-357)."""
        self.engine.set_dirichlet_values(self.dirichlet_values(time))
        self.engine.assemble_first(time)

    def assemble_time_step(self, time):
        """NavierStokes::assemble_time_step (NavierStokes2D.cpp:361-527)."""
        self.engine.set_dirichlet_values(self.dirichlet_values(time))
        self.engine.assemble_step(time)

    def solve_time_step(self, time=None):
        """NavierStokes::solve_time_step (NavierStokes2D.cpp:530-639)."""
        its, tp, ts = self.engine.solve_step()
        self.time_prec.append(tp)
        self.time_solve.append(ts)
        self.gmres_iterations.append(its)
        if self.verbose:
            print(f"Time taken to initialize preconditioner: {tp} seconds")
            print(f"Time taken to solve Navier Stokes problem: {ts} seconds")
            print(f"Result:  {its} GMRES iterations")
        return its

    def compute_forces(self, time=None, rho=1.0):
        """NavierStokes::compute_forces (NavierStokes2D.cpp:752-859, NavierStokes3D.cpp:744-840): the face
        integrals run on the device (nsb_compute_forces, summed over ranks); returns [c_d, c_l] with
        c = 2 F / (mean_v^2 D) in 2D and 2 F / (rho mean_v^2 D H) in 3D."""
        drag, lift = self.engine.compute_forces(rho)
        mean_v = mean_velocity(self.dim, self.time if time is None else time, self.test_case)
        D, H = 0.1, 0.41
        den = mean_v * mean_v * D if self.dim == 2 else rho * mean_v * mean_v * D * H
        c_d, c_l = 2.0 * drag / den, 2.0 * lift / den
        self.vec_drag.append(drag); self.vec_lift.append(lift)
        self.vec_drag_coeff.append(c_d); self.vec_lift_coeff.append(c_l)
        return [c_d, c_l]

    def initial_condition(self):
        """VectorTools::interpolate(dof_handler, u_0, solution_owned) (NavierStokes2D.cpp:708)."""
        x = np.zeros(self.N)
        if self.variant == "conv":
            u, _ = exact_solution(self.dofs.node_xyz, 0.0)
            _, p = exact_solution(self.dofs.p_xyz, 0.0)
            x[: self.n_u] = u.ravel()
            x[self.n_u:] = p
        return x

    # ------------------------------------------------------------------ solve()
    def solve(self, max_steps=None):
        """NavierStokes::solve (NavierStokes2D.cpp:699-750): the time loop."""
        e = self.engine
        e.set_solution(self.initial_condition())
        time, step = 0.0, 0
        while time < self.T - 0.5 * self.deltat:
            if self.variant == "conv":  # function_h.set_time(time) BEFORE the increment (Convergence3D.cpp:747-750)
                e.set_neumann_rhs(self.neumann_rhs(time))
            time += self.deltat
            step += 1
            if time == self.deltat:
                self.assemble(time)
            else:
                self.assemble_time_step(time)
            self.solve_time_step(time)
            self.time = time
            # NavierStokes2D.cpp:736-740 (every step), NavierStokes3D.cpp:728-733 (only once time > 0.1)
            if self.variant == "2d" or (self.variant == "3d" and time > self.forces_after):
                c = self.compute_forces(time)
                self.c_D_max, self.c_L_min = max(self.c_D_max, c[0]), min(self.c_L_min, c[1])
            if max_steps is not None and step >= max_steps:
                break
        self.time, self.n_steps = time, step
        self.solution = e.get_solution()
        return self.solution
