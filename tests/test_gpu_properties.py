"""Size-independent properties of the CUDA hot path at bench size (`-m gpu`).

The oracle cannot follow at millions of DoFs in test time, so these checks use identities of the
discretisation and of the solver that hold at any size, on the 2 M-DoF refined 3D cylinder
(`bench.py` workload `cyl3d-2M`, same mesh family as BASELINE.json configs[4]) in the throughput
configuration `bench.py` runs (block multicolour ILU(0), batched Gram-Schmidt) -- i.e. through the SELL-32 /
block-sequential kernels that the small parity cases only touch with a handful of slices.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from navierstokes_project_nm4pde_b200 import HostMesh, NavierStokes  # noqa: E402

pytestmark = pytest.mark.gpu

WORKLOAD = "cyl3d-2M"
DT = bench.DELTAT["3d"]


@pytest.fixture(scope="module")
def prob():
    s, nz = bench.WORKLOADS[WORKLOAD][1]
    p = NavierStokes(HostMesh.cylinder3d(s, nz), "3d", T=1.0, deltat=DT, test_case=2, ilu_ordering=2,
                     orthogonalisation=1)
    p.setup()
    rng = np.random.default_rng(20240607)
    x = np.zeros(p.N)
    x[: p.n_u] = rng.uniform(-1.0, 1.0, p.n_u)
    e = p.engine
    e.set_solution(x)
    e.set_dirichlet_values(p.dirichlet_values(DT))
    e.assemble_first(DT)
    p.rng = rng
    return p


def _interior_mask(p):
    m = np.ones(p.n_u, bool)
    m[p._dir_rows] = False
    return m


def test_divergence_of_constant_velocity_vanishes(prob):
    """Block (1,0): sum_j B_ij c = int psi_i div(c) = 0 for a constant field (columns are never
    eliminated, NavierStokes2D.cpp:354), for every one of the ~9e4 pressure rows."""
    e = prob.engine
    c = np.tile([0.3, -1.1, 0.7], prob.n_u // 3)
    y = e.block_vmult("B", c)
    scale = np.abs(e.block_vmult("B", prob.rng.uniform(-1.0, 1.0, prob.n_u))).max()  # size of un-cancelled row sums
    assert np.abs(y).max() < 1e-11 * scale


def test_spmv_is_linear_and_blocks_compose(prob):
    """system_matrix.vmult is linear and equals its blocks: y_u = F x_u + Bt x_p, y_p = B x_u."""
    e, rng, nu = prob.engine, prob.rng, prob.n_u
    x, z = rng.uniform(-1, 1, prob.N), rng.uniform(-1, 1, prob.N)
    a, b = 0.37, -2.5
    lhs = e.system_vmult(a * x + b * z)
    rhs = a * e.system_vmult(x) + b * e.system_vmult(z)
    assert np.linalg.norm(lhs - rhs) < 1e-12 * np.linalg.norm(rhs)
    y = e.system_vmult(x)
    yu = e.block_vmult("F", x[:nu]) + e.block_vmult("Bt", x[nu:])
    yp = e.block_vmult("B", x[:nu])
    assert np.linalg.norm(y[:nu] - yu) < 1e-13 * np.linalg.norm(yu)
    assert np.linalg.norm(y[nu:] - yp) < 1e-13 * np.linalg.norm(yp)


def test_step_assembly_rhs_identity(prob):
    """For a constant advecting state u = c: stiffness and convection rows sum to zero, so on every
    unconstrained row F 1 = (M/dt) 1 and the assembled rhs = M u / dt = c (F 1) -- ties the matrix
    scatter (atomics through mapF), the rhs and the K + C(u) composition of assemble_time_step
    together over all 4.7e5 cells."""
    e, nu = prob.engine, prob.n_u
    c = np.array([0.8, -0.4, 0.25])
    x = np.zeros(prob.N)
    x[:nu] = np.tile(c, nu // 3)
    saved = e.get_solution()
    e.set_solution(x)
    e.assemble_step(2 * DT)
    F1 = e.block_vmult("F", np.ones(nu))
    rhs = e.get_rhs()[:nu]
    m = _interior_mask(prob)
    expect = np.tile(c, nu // 3) * F1
    assert np.abs(rhs[m] - expect[m]).max() < 1e-11 * np.abs(expect[m]).max()
    # total mass: sum_i (M 1)_i / dt over ALL rows is |Omega| / dt per component; interior rows give less
    assert 0.0 < F1[m].sum() * DT / 3.0 < 2.5 * 0.41 * 0.41
    e.set_solution(saved)
    e.assemble_step(2 * DT)


def test_ilu_apply_is_linear_and_a_good_inverse(prob):
    """The multicolour SELL-32 sweeps: linear, and U^-1 D^-1 L^-1 F x ~ x for the mass-dominated F
    (dt = 2e-4): the preconditioned operator stays within 0.5 of the identity in relative l2."""
    e, rng, nu = prob.engine, prob.rng, prob.n_u
    e.precond_init()
    x, z = rng.uniform(-1, 1, nu), rng.uniform(-1, 1, nu)
    lhs = e.ilu_apply(0, 0.5 * x - 3.0 * z)
    rhs = 0.5 * e.ilu_apply(0, x) - 3.0 * e.ilu_apply(0, z)
    assert np.linalg.norm(lhs - rhs) < 1e-12 * np.linalg.norm(rhs)
    w = e.ilu_apply(0, e.block_vmult("F", x))
    assert np.linalg.norm(w - x) < 0.5 * np.linalg.norm(x)
    xp = rng.uniform(-1, 1, prob.n_p)
    ws = e.ilu_apply(1, e.block_vmult("S", xp))
    assert np.linalg.norm(ws - xp) < 0.9 * np.linalg.norm(xp)


def test_solve_reduces_the_preconditioned_residual(prob):
    """solve_time_step stops on GMRES's estimate of the preconditioned residual (absolute 1e-4,
    NavierStokes2D.cpp:535).  With the reference's inexact inner solves (rel. 1e-2) the preconditioner
    is a slightly different operator on every application, so the TRUE preconditioned residual
    recomputed from the operators is only guaranteed to the inner tolerance: it must have dropped to
    well below 2% of |P^-1 b| (here: a rough random state, pressure values of O(50))."""
    e = prob.engine
    its, _, _ = e.solve_step()
    assert its > 0
    assert e.stat("n_inner_F") > 0 and e.stat("n_F_solves") == 2 * e.stat("n_S_solves")  # Yosida: two F solves per vmult
    x, b = e.get_solution(), e.get_rhs()
    z = e.precond_vmult(b - e.system_vmult(x))
    z0 = e.precond_vmult(b)
    assert np.isfinite(z).all() and np.linalg.norm(z) < 2e-2 * np.linalg.norm(z0)


def test_midsize_parity_against_the_oracle():
    """0.53 M DoF (`cyl3d-500k`, the mesh of bench.py's CPU legs; ~5.5e3 blocks of 32 rows per factor, all
    16 block colours populated): the deterministic operators of the throughput configuration against the
    oracle in the same ILU ordering -- block SpMV, Schur product, both ILU(0) applies to 1e-10 -- and one
    preconditioner application / one whole time step to the accuracy the inexact inner solves allow."""
    import helpers as T
    from oracle import ns_ref as R

    s, nz = bench.WORKLOADS["cyl3d-500k"][1]
    mesh = HostMesh.cylinder3d(s, nz)
    p = NavierStokes(mesh, "3d", T=1.0, deltat=DT, test_case=2, ilu_ordering=2, orthogonalisation=1)
    p.setup()
    d, e = p.dofs, p.engine
    num = dict(dim=3, cell_dofs=d.cell_dofs(), N=d.N, n_u=d.n_u, n_p=d.n_p, dpc=d.dpc)
    o = R.Oracle(3, "3d", mesh.vertices, mesh.cells, num, R.system_pattern(num), 1e-3, DT)
    ou = (3 * e.ilu_order(0)[:, None] + np.arange(3)[None, :]).ravel()
    o.set_ilu_order(ou, e.ilu_order(1))
    o.set_orthogonalisation(1)
    assert e.stat("ilu_blocks_F") > 5000 and e.stat("sweeps_F") <= 24
    vals = p.dirichlet_values(DT)
    o.set_dirichlet(p._dir_rows, vals)
    e.set_dirichlet_values(vals)
    rng = np.random.default_rng(T.SEED)
    x0 = np.zeros(d.N)
    x0[: d.n_u] = 0.05 * rng.uniform(-1.0, 1.0, d.n_u)
    for side in (o, e):
        side.set_solution(x0)
        side.assemble_first()
    o.precond_init("yosida"); e.precond_init()
    x = rng.uniform(-1.0, 1.0, d.N)
    xu, xp = x[: d.n_u], x[d.n_u:]
    assert T.rel_l2(e.system_vmult(x), o.system_vmult(x)) < 1e-12
    assert T.rel_l2(e.ilu_apply(0, xu), o.ilu_apply(0, xu)) < 1e-10
    assert T.rel_l2(e.ilu_apply(1, xp), o.ilu_apply(1, xp)) < 1e-10
    assert T.rel_l2(e.block_vmult("S", xp), o.block_vmult(3, xp, d.n_p)) < 1e-11
    # one preconditioner application: inner Krylov solves stop on a tolerance, so agreement is to the
    # level at which both sides take the same inner iteration counts (they do unless a residual sits on
    # the threshold); 1e-6 still separates "same algorithm" from "different algorithm" by 4 digits
    ze, zo = e.precond_vmult(x), o.precond_vmult("yosida", x)
    assert T.rel_l2(ze, zo) < 1e-6
    # one whole time step, the reference's own first one (impulsive start from u = 0, ~80 outer iterations)
    for side in (o, e):
        side.set_solution(np.zeros(d.N))
        side.assemble_first()
    rc, its_o, _ = o.solve_step("yosida")
    its_e, _, _ = e.solve_step()
    assert rc == 0 and abs(its_e - its_o) <= 2, (its_e, its_o)
    assert T.rel_l2(e.get_solution()[: d.n_u], o.array("sol_owned", d.N)[: d.n_u]) < 1e-5


def test_2d_refined_cylinder_2M_properties():
    """BASELINE.json configs[3]: the globally refined 2D cylinder (~2 M DoF, `bench.py` workload `cyl2d-2M`,
    NavierStokes2D + aSIMPLE, NavierStokes2D.cpp:547, Preconditioners.hpp:254-311) through the 2D kernels at
    bench size: divergence of a constant field, linearity and block composition of the system SpMV, the
    rhs / F.1 identity of assemble_time_step (with the Temam term, which 2D keeps), linearity of both ILU(0)
    applies, and drag / lift on the device against the host face loop on the same field.  The time step itself is
    checked against the oracle on the 0.16 M-DoF mesh of the same family (next test): at 2 M DoF the reference's
    inner GMRES(28) + ILU(0) on the Schur complement needs thousands of iterations per solve -- on the CPU
    restatement 314 per solve at 0.16 M DoF and 1 058 at 0.64 M (natural ordering) -- which is the reference
    algorithm's own limit on this configuration (it throws NoConvergence at 10 000), not a property to assert on."""
    dt = bench.DELTAT["2d"]
    (s,) = bench.WORKLOADS["cyl2d-2M"][1]
    p = NavierStokes(HostMesh.cylinder2d(s), "2d", T=1.0, deltat=dt, test_case=2, ilu_ordering=1, orthogonalisation=1)
    p.setup()
    assert 1.5e6 < p.N < 3e6
    e, nu, rng = p.engine, p.n_u, np.random.default_rng(20240607)
    x0 = np.zeros(p.N)
    x0[:nu] = 0.1 * rng.uniform(-1.0, 1.0, nu)
    x0[nu:] = rng.uniform(-1.0, 1.0, p.N - nu)
    e.set_solution(x0)
    e.set_dirichlet_values(p.dirichlet_values(2.0 + dt))
    e.assemble_first(dt)
    # B c = 0
    y = e.block_vmult("B", np.tile([0.3, -1.1], nu // 2))
    scale = np.abs(e.block_vmult("B", rng.uniform(-1.0, 1.0, nu))).max()
    assert np.abs(y).max() < 1e-11 * scale
    # linearity + blocks
    x, z = rng.uniform(-1, 1, p.N), rng.uniform(-1, 1, p.N)
    lhs = e.system_vmult(0.37 * x - 2.5 * z)
    rhs = 0.37 * e.system_vmult(x) - 2.5 * e.system_vmult(z)
    assert np.linalg.norm(lhs - rhs) < 1e-12 * np.linalg.norm(rhs)
    yy = e.system_vmult(x)
    assert np.linalg.norm(yy[:nu] - e.block_vmult("F", x[:nu]) - e.block_vmult("Bt", x[nu:])) < 1e-13 * np.linalg.norm(yy[:nu])
    # drag / lift: device kernel == host face loop on the same (rough) field
    f_dev = e.compute_forces()
    f_host = p.dofs.boundary_forces(x0, 3, p.nu, 1.0)
    assert np.allclose(f_dev, f_host, rtol=1e-10, atol=1e-13 * np.abs(f_host).max())
    # constant advecting state: rhs = c (F 1) on unconstrained rows (Temam term vanishes for div c = 0)
    c = np.array([0.8, -0.4])
    xc = np.zeros(p.N)
    xc[:nu] = np.tile(c, nu // 2)
    e.set_solution(xc)
    e.assemble_step(2 * dt)
    F1 = e.block_vmult("F", np.ones(nu))
    m = np.ones(nu, bool)
    m[p._dir_rows] = False
    expect = np.tile(c, nu // 2) * F1
    assert np.abs(e.get_rhs()[:nu][m] - expect[m]).max() < 1e-11 * np.abs(expect[m]).max()
    # ILU(0) applies are linear (velocity block and Schur complement, through the colour sweeps)
    e.precond_init()
    for which, n in ((0, nu), (1, p.N - nu)):
        xu, zu = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
        lhs = e.ilu_apply(which, 0.5 * xu - 3.0 * zu)
        rhs = 0.5 * e.ilu_apply(which, xu) - 3.0 * e.ilu_apply(which, zu)
        assert np.isfinite(lhs).all() and np.linalg.norm(lhs - rhs) < 1e-12 * np.linalg.norm(rhs)


@pytest.mark.parametrize("ordering", [1, 2])
def test_2d_midsize_time_step_against_the_oracle(ordering):
    """0.16 M DoF (`cyl2d-160k`, the 2D mesh of bench.py's CPU legs): NavierStokes2D + aSIMPLE in the throughput
    configuration against the oracle factorising in the engine's ILU order -- ILU applies and Schur product to
    1e-10, then the reference's first time step.  Its inner GMRES on the Schur complement runs 300-500 iterations
    per solve (restarted GMRES(28) close to stagnation), so iteration counts are compared with a margin and the
    fields to the accuracy of the outer tolerance."""
    import helpers as T
    from oracle import ns_ref as R

    dt = bench.DELTAT["2d"]
    (s,) = bench.WORKLOADS["cyl2d-160k"][1]
    mesh = HostMesh.cylinder2d(s)
    p = NavierStokes(mesh, "2d", T=1.0, deltat=dt, test_case=2, ilu_ordering=ordering, orthogonalisation=1)
    p.setup()
    d, e = p.dofs, p.engine
    num = dict(dim=2, cell_dofs=d.cell_dofs(), N=d.N, n_u=d.n_u, n_p=d.n_p, dpc=d.dpc)
    o = R.Oracle(2, "2d", mesh.vertices, mesh.cells, num, R.system_pattern(num), 1e-3, dt)
    ou = (2 * e.ilu_order(0)[:, None] + np.arange(2)[None, :]).ravel()
    o.set_ilu_order(ou, e.ilu_order(1))
    o.set_orthogonalisation(1)
    vals = p.dirichlet_values(dt)
    o.set_dirichlet(p._dir_rows, vals)
    e.set_dirichlet_values(vals)
    for side in (o, e):
        side.set_solution(np.zeros(d.N))
        side.assemble_first()
    o.precond_init("asimple"); e.precond_init()
    rng = np.random.default_rng(T.SEED)
    x = rng.uniform(-1.0, 1.0, d.N)
    xu, xp = x[: d.n_u], x[d.n_u:]
    assert T.rel_l2(e.system_vmult(x), o.system_vmult(x)) < 1e-12
    assert T.rel_l2(e.ilu_apply(0, xu), o.ilu_apply(0, xu)) < 1e-10
    assert T.rel_l2(e.ilu_apply(1, xp), o.ilu_apply(1, xp)) < 1e-10
    assert T.rel_l2(e.block_vmult("S", xp), o.block_vmult(3, xp, d.n_p)) < 1e-11
    rc, its_o, _ = o.solve_step("asimple")
    its_e, _, _ = e.solve_step()
    assert rc == 0 and abs(its_e - its_o) <= max(3, its_o // 5), (its_e, its_o)
    assert e.stat("n_F_solves") == e.stat("n_S_solves") > 0  # aSIMPLE: one F and one Schur solve per vmult
    xe, xo = e.get_solution(), o.array("sol_owned", d.N)
    assert T.rel_l2(xe[: d.n_u], xo[: d.n_u]) < 1e-3
    # the true preconditioned residual of the engine's solution has dropped like the solver says
    b = e.get_rhs()
    zr, z0 = e.precond_vmult(b - e.system_vmult(xe)), e.precond_vmult(b)
    assert np.isfinite(zr).all() and np.linalg.norm(zr) < 5e-2 * np.linalg.norm(z0)
