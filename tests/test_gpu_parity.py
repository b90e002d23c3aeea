"""Parity of the CUDA hot path against the CPU oracle, called through the C ABI (`-m gpu`).

Tolerances are the north_star's: assembled entries 1e-12 relative (to the row scale, SURVEY.md
H7), fields 1e-8 relative L2 after each step.  Sizes are small enough for the oracle to finish
in seconds; size-independent properties at full size are in test_gpu_properties.py.
"""
import numpy as np
import pytest

import helpers as T

pytestmark = pytest.mark.gpu

CASES = ["cyl2d", "box2d", "box3d", "cube", "cyl3d"]
ENTRY_TOL = 1e-12
FIELD_TOL = 1e-8


def _setup(case_name, state_scale=0.3, time=4.0, **params):
    case = T.Case(case_name)
    rows, vals = case.bc(time)
    o, e = case.oracle(), case.engine(**params)
    o.set_dirichlet(rows, vals)
    e.set_dirichlet(rows)
    e.set_dirichlet_values(vals)
    x0 = state_scale * case.random_state()
    o.set_solution(x0)
    e.set_solution(x0)
    if case.variant == "conv":  # Neumann face term of Convergence3D.cpp:309-330
        neu = case.neumann(0.0)
        o.set_neumann_rhs(neu)
        e.set_neumann_rhs(neu[: case.n_u])
    return case, o, e


@pytest.mark.parametrize("case_name", CASES)
def test_sparsity_pattern_matches_reference_layout(case_name):
    """a5: make_sparsity_pattern with the coupling table of NavierStokes2D.cpp:109-119."""
    case = T.Case(case_name)
    o, e = case.oracle(), case.engine()
    rp, ci, pm_rp, pm_ci = o.pattern4
    import scipy.sparse as sp

    A = sp.csr_matrix((np.ones(len(ci)), ci, rp), shape=(case.N, case.N))
    nu = case.n_u
    for blk, M in (("F", A[:nu, :nu]), ("Bt", A[:nu, nu:]), ("B", A[nu:, :nu])):
        erp, eci = e.pattern(blk)
        assert T.same_pattern(M, erp, eci), blk
        M = M.tocsr(); M.sort_indices()
        e.check_pattern(blk, M.indptr, M.indices)  # the optional hand-in path accepts the reference pattern
    erp, eci = e.pattern("Mp")
    assert np.array_equal(erp, pm_rp) and np.array_equal(eci, pm_ci)
    # a corrupted pattern must be refused
    bad = eci.copy(); bad[0] += 1
    with pytest.raises(Exception):
        e.check_pattern("Mp", erp, bad)


@pytest.mark.parametrize("case_name", CASES)
def test_assemble_first_entries(case_name):
    """a1: NavierStokes::assemble -- all five matrices and the rhs, entrywise."""
    case, o, e = _setup(case_name)
    o.assemble_first()
    e.assemble_first()
    for mat, oname in (("system", "sys"), ("mass", "mass"), ("stiffness", "stiff")):
        ob = T.oracle_blocks(o, oname)
        for blk in ("F", "Bt", "B"):
            err = T.entry_error(e.matrix(mat, blk), ob[blk])
            assert err < ENTRY_TOL, (mat, blk, err)
    import scipy.sparse as sp

    Mp_o = sp.csr_matrix((o.array("pmass", o.pm_nnz), o.pattern4[3], o.pattern4[2]), shape=(case.n_p, case.n_p))
    assert T.entry_error(e.matrix("system", "Mp"), Mp_o) < ENTRY_TOL
    rhs_o, rhs_e = o.array("rhs", case.N), e.get_rhs()
    assert np.max(np.abs(rhs_e - rhs_o)) <= ENTRY_TOL * max(1.0, np.max(np.abs(rhs_o)))


@pytest.mark.parametrize("kernel", [0, 1])
@pytest.mark.parametrize("case_name", CASES)
def test_assemble_time_step_entries(case_name, kernel):
    """a2 + a4: assemble_time_step (both kernels) incl. Dirichlet row clearing, over 2 steps."""
    case, o, e = _setup(case_name, assembly_kernel=kernel)
    o.assemble_first()
    e.assemble_first()
    rng = np.random.default_rng(7)
    for step in range(2):
        x = 0.5 * rng.uniform(-1, 1, case.N)
        o.set_solution(x)
        e.set_solution(x)
        o.assemble_step()
        e.assemble_step()
        ob = T.oracle_blocks(o, "sys")
        for blk in ("F", "Bt", "B"):
            err = T.entry_error(e.matrix("system", blk), ob[blk])
            assert err < ENTRY_TOL, (step, blk, err)
        errc = T.entry_error(e.matrix("convection", "F"), T.oracle_blocks(o, "conv")["F"])
        assert errc < ENTRY_TOL, errc
        rhs_o, rhs_e = o.array("rhs", case.N), e.get_rhs()
        assert np.max(np.abs(rhs_e - rhs_o)) <= ENTRY_TOL * max(1.0, np.max(np.abs(rhs_o)))


@pytest.mark.parametrize("case_name", CASES)
def test_dirichlet_replace_mode(case_name):
    """a4: the alternative 'replace diagonal by dbar' reading of apply_boundary_values."""
    case, o, e = _setup(case_name, dirichlet_mode=1)
    o.set_options(dirichlet_mode=1)
    o.assemble_first()
    e.assemble_first()
    assert T.entry_error(e.matrix("system", "F"), T.oracle_blocks(o, "sys")["F"]) < ENTRY_TOL
    assert np.max(np.abs(e.get_rhs() - o.array("rhs", case.N))) <= ENTRY_TOL * max(1.0, np.max(np.abs(o.array("rhs", case.N))))


@pytest.mark.parametrize("case_name", CASES)
def test_spmv_blocks(case_name):
    """a9 operator: BlockSparseMatrix::vmult and the block products used by the preconditioners."""
    case, o, e = _setup(case_name)
    o.assemble_first()
    e.assemble_first()
    o.precond_init(T.VARIANT_PREC[case.variant])
    x = case.random_state()
    yo, ye = o.system_vmult(x), e.system_vmult(x)
    assert T.rel_l2(ye, yo) < 1e-13
    xu, xp = x[: case.n_u], x[case.n_u:]
    assert T.rel_l2(e.block_vmult("F", xu), o.block_vmult(0, xu, case.n_u)) < 1e-13
    assert T.rel_l2(e.block_vmult("Bt", xp), o.block_vmult(1, xp, case.n_u)) < 1e-13
    assert T.rel_l2(e.block_vmult("B", xu), o.block_vmult(2, xu, case.n_p)) < 1e-13


@pytest.mark.parametrize("ptype", ["asimple", "yosida", "simple", "ayosida"])
@pytest.mark.parametrize("case_name", ["cyl2d", "box3d", "cube"])
def test_preconditioner_setup(case_name, ptype):
    """a7/a8/a10/a11: diag extraction, Schur product (mmult) and the two ILU(0) factorisations."""
    case, o, e = _setup(case_name, precond_type=ptype)
    o.assemble_first()
    e.assemble_first()
    o.precond_init(ptype)
    e.precond_init()
    So, Se = o.schur(), e.schur()
    assert T.same_pattern(So, Se.indptr, Se.indices)
    assert T.entry_error(Se, So) < ENTRY_TOL
    x = case.random_state()
    xu, xp = x[: case.n_u], x[case.n_u:]
    assert T.rel_l2(e.ilu_apply(0, xu), o.ilu_apply(0, xu)) < 1e-11
    assert T.rel_l2(e.ilu_apply(1, xp), o.ilu_apply(1, xp)) < 1e-11
    assert T.rel_l2(e.block_vmult("S", xp), o.block_vmult(3, xp, case.n_p)) < 1e-12


@pytest.mark.parametrize("ptype", ["asimple", "yosida", "simple", "ayosida"])
@pytest.mark.parametrize("case_name", ["cyl2d", "box3d", "cube"])
def test_preconditioner_vmult(case_name, ptype):
    """a7/a8: one application of P^{-1} (inner GMRES/CG to rel 1e-2 with ILU(0))."""
    case, o, e = _setup(case_name, precond_type=ptype)
    o.assemble_first()
    e.assemble_first()
    o.precond_init(ptype)
    e.precond_init()
    src = case.random_state()
    dst0 = 0.1 * np.random.default_rng(3).uniform(-1, 1, case.N)  # aSIMPLE uses dst as initial guess
    yo = o.precond_vmult(ptype, src, dst0)
    ye = e.precond_vmult(src, dst0)
    assert e.stat("n_inner_F") == o.stat("n_inner_F")
    assert e.stat("n_inner_S") == o.stat("n_inner_S")
    assert T.rel_l2(ye, yo) < FIELD_TOL


@pytest.mark.parametrize("case_name,ptype", [("cyl2d", "asimple"), ("box2d", "asimple"), ("box3d", "yosida"),
                                             ("cube", "yosida"), ("cyl3d", "yosida"), ("cyl2d", "simple"),
                                             ("box3d", "ayosida")])
def test_time_steps_fields(case_name, ptype):
    """a6: three full time steps (assemble | assemble_time_step -> solve_time_step): same outer
    iteration counts and fields within 1e-8 relative L2 after each step."""
    case = T.Case(case_name)
    o, e = case.oracle(), case.engine(precond_type=ptype)
    rows, vals = case.bc(0.0)
    o.set_dirichlet(rows, vals)
    e.set_dirichlet(rows)
    x0 = case.initial()
    o.set_solution(x0)
    e.set_solution(x0)
    t = 0.0
    for step in range(3):
        t += case.dt
        tb = t if case.variant == "conv" else 2.0 + t  # non-trivial inlet amplitude
        rows, vals = case.bc(tb)
        o.set_dirichlet_values(vals)
        e.set_dirichlet_values(vals)
        if case.variant == "conv":
            neu = case.neumann(t - case.dt)  # function_h lags by one step (Convergence3D.cpp:747-750)
            o.set_neumann_rhs(neu)
            e.set_neumann_rhs(neu[: case.n_u])
        if step == 0:
            o.assemble_first(); e.assemble_first()
        else:
            o.assemble_step(); e.assemble_step()
        rc, its_o, _ = o.solve_step(ptype)
        its_e, _, _ = e.solve_step()
        assert rc == 0
        assert its_e == its_o, (step, its_e, its_o)
        xo, xe = o.array("sol_owned", case.N), e.get_solution()
        nu = case.n_u
        assert T.rel_l2(xe[:nu], xo[:nu]) < FIELD_TOL, (step, "velocity")
        assert np.linalg.norm(xe[nu:] - xo[nu:]) < FIELD_TOL * max(np.linalg.norm(xo[nu:]), np.linalg.norm(xo[:nu])), (step, "pressure")


@pytest.mark.parametrize("case_name,ptype", [("cyl2d", "asimple"), ("cyl3d", "yosida")])
def test_drag_lift_coefficients(case_name, ptype):
    """f1 / config 3: drag and lift of NavierStokes::compute_forces (NavierStokes2D.cpp:752-859,
    NavierStokes3D.cpp:744-840) computed ON THE DEVICE (nsb_compute_forces) after each of three time
    steps against the oracle's restatement evaluated on the oracle's own solution: within 1e-6
    relative (north_star tolerance); and against the oracle evaluated on the engine's solution
    (isolates the face kernel from the solver tolerance): 1e-12."""
    DRAG_LIFT_TOL = 1e-6
    case = T.Case(case_name)
    o, e = case.oracle(), case.engine(precond_type=ptype)
    rows, vals = case.bc(0.0)
    o.set_dirichlet(rows, vals)
    e.set_dirichlet(rows)
    fc, fl = case.dofs.boundary_faces(3)
    xi, w = T.gauss_simplex(case.dim - 1)
    assert len(fc) > 0
    e.set_force_faces(fc, fl, xi, w)
    x0 = case.initial()
    o.set_solution(x0); e.set_solution(x0)
    t = 0.0
    for step in range(3):
        t += case.dt
        rows, vals = case.bc(2.0 + t)
        o.set_dirichlet_values(vals); e.set_dirichlet_values(vals)
        if step == 0:
            o.assemble_first(); e.assemble_first()
        else:
            o.assemble_step(); e.assemble_step()
        rc, its_o, _ = o.solve_step(ptype)
        its_e, _, _ = e.solve_step()
        assert rc == 0 and its_e == its_o
        f_dev = e.compute_forces(rho=1.0)
        f_ref = o.compute_forces(o.array("sol_owned", case.N), fc, fl, xi, w, rho=1.0)
        f_same = o.compute_forces(e.get_solution(), fc, fl, xi, w, rho=1.0)
        scale = np.abs(f_ref).max()
        assert np.all(np.abs(f_dev - f_same) <= 1e-12 * scale), (step, f_dev, f_same)
        assert np.all(np.abs(f_dev - f_ref) <= DRAG_LIFT_TOL * np.abs(f_ref) + 1e-9 * scale), (step, f_dev, f_ref)
    # error behaviour: forces before the faces were handed in
    e2 = case.engine(precond_type=ptype)
    from navierstokes_project_nm4pde_b200._lib import NsbError
    with pytest.raises(NsbError):
        e2.compute_forces()


def test_bad_arguments_are_refused():
    """Error behaviour of the boundary: negative return codes + message, no exceptions across the ABI."""
    from navierstokes_project_nm4pde_b200._lib import NsbError

    case = T.Case("box2d")
    e = case.engine()
    with pytest.raises(NsbError) as ei:
        e.set_params(precond_type=7)  # reference: std::runtime_error("Invalid preconditioner type")
    assert "Invalid preconditioner type" in str(ei.value)
    with pytest.raises(NsbError):
        e.set_dirichlet(np.array([0], np.int32))  # only one component of a node
    with pytest.raises(NsbError):
        e.solve_step()  # nothing assembled
    cd = case.dofs.cell_dofs()
    cc = case.dofs.cell_coords().copy()
    cc[0, [0, 1]] = cc[0, [1, 0]]  # negative Jacobian
    from navierstokes_project_nm4pde_b200 import Engine

    e2 = Engine(2)
    with pytest.raises(NsbError):
        e2.set_mesh(cc, cd, case.n_u, case.n_p)


def _oracle_order(case, e):
    """Hand the engine's ILU orderings to the oracle (F order is per P2 node -> all components)."""
    ou = (case.dim * e.ilu_order(0)[:, None] + np.arange(case.dim)[None, :]).ravel()
    return ou, e.ilu_order(1)


@pytest.mark.parametrize("ordering", [1, 2, 3])
@pytest.mark.parametrize("case_name,ptype", [("cyl2d", "asimple"), ("box3d", "yosida"), ("cube", "yosida"),
                                             ("cyl3d", "yosida"), ("box2d", "simple"), ("box3d", "ayosida")])
def test_multicolour_ilu_mode(case_name, ptype, ordering, monkeypatch):
    """Throughput modes (ilu_ordering = 1: point multicolour, 2: block multicolour with sequential
    elimination inside 32-row blocks, 3: subdomain ordering -- parts solved by one CTA out of shared
    memory, separator rows last; small parts here so that the test meshes have several of them): ILU(0)
    of the permuted matrices.  Same checks as the replay
    mode, against the oracle factorising in the same ordering."""
    if ordering == 3:  # three levels of small parts, no lower bound on the rows of a level
        monkeypatch.setenv("NSB_SD_LEAF", "96,32,32")
        monkeypatch.setenv("NSB_SD_MIN_ACTIVE", "0")
    case = T.Case(case_name)
    o, e = case.oracle(), case.engine(precond_type=ptype, ilu_ordering=ordering)
    ou, op = _oracle_order(case, e)
    assert sorted(ou.tolist()) == list(range(case.n_u)) and sorted(op.tolist()) == list(range(case.n_p))
    assert not np.array_equal(op, np.arange(case.n_p))
    o.set_ilu_order(ou, op)
    rows, vals = case.bc(0.0)
    o.set_dirichlet(rows, vals)
    e.set_dirichlet(rows)
    x0 = case.initial()
    o.set_solution(x0)
    e.set_solution(x0)
    t = 0.0
    for step in range(3):
        t += case.dt
        rows, vals = case.bc(t if case.variant == "conv" else 2.0 + t)
        o.set_dirichlet_values(vals)
        e.set_dirichlet_values(vals)
        if case.variant == "conv":
            neu = case.neumann(t - case.dt)
            o.set_neumann_rhs(neu)
            e.set_neumann_rhs(neu[: case.n_u])
        if step == 0:
            o.assemble_first(); e.assemble_first()
        else:
            o.assemble_step(); e.assemble_step()
        if step == 0:
            o.precond_init(ptype); e.precond_init()
            x = case.random_state()
            xu, xp = x[: case.n_u], x[case.n_u:]
            assert T.rel_l2(e.ilu_apply(0, xu), o.ilu_apply(0, xu)) < 1e-11
            assert T.rel_l2(e.ilu_apply(1, xp), o.ilu_apply(1, xp)) < 1e-11
        rc, its_o, _ = o.solve_step(ptype)
        its_e, _, _ = e.solve_step()
        assert rc == 0 and its_e == its_o, (step, its_e, its_o)
        xo, xe = o.array("sol_owned", case.N), e.get_solution()
        assert T.rel_l2(xe[: case.n_u], xo[: case.n_u]) < FIELD_TOL, step
    # far fewer dependent launches per triangular solve than the natural ordering has levels
    assert e.stat("sweeps_F") <= 40 and e.stat("sweeps_S") <= 80


@pytest.mark.parametrize("case_name,ptype,ord_f,ord_s", [("cyl3d", "yosida", 2, 1), ("cyl2d", "asimple", 2, 1),
                                                         ("box3d", "yosida", 1, 0), ("cube", "yosida", 2, 1)])
def test_separate_ordering_of_the_schur_factors(case_name, ptype, ord_f, ord_s):
    """nsb_params.ilu_ordering_schur: the Schur-complement factors in a different elimination order than F_s
    (bench.py at 19.9 M DoF: block multicolour for F_s, point multicolour for the pressure matrix); the oracle
    factorises each matrix in the order the engine reports.  Batched Gram-Schmidt as in the bench."""
    case = T.Case(case_name)
    o = case.oracle()
    e = case.engine(precond_type=ptype, ilu_ordering=ord_f, ilu_ordering_schur=ord_s, orthogonalisation=1)
    ref = case.engine(precond_type=ptype, ilu_ordering=ord_s, orthogonalisation=1)
    assert np.array_equal(e.ilu_order(1), ref.ilu_order(1))  # the pressure ordering is the one asked for
    if ord_f != ord_s:
        assert not np.array_equal(e.ilu_order(0), ref.ilu_order(0))
    del ref
    o.set_ilu_order(*_oracle_order(case, e))
    o.set_orthogonalisation(1)
    rows, vals = case.bc(0.0)
    o.set_dirichlet(rows, vals)
    e.set_dirichlet(rows)
    x0 = case.initial()
    o.set_solution(x0)
    e.set_solution(x0)
    t = 0.0
    for step in range(2):
        t += case.dt
        rows, vals = case.bc(t if case.variant == "conv" else 2.0 + t)
        o.set_dirichlet_values(vals)
        e.set_dirichlet_values(vals)
        if case.variant == "conv":
            neu = case.neumann(t - case.dt)
            o.set_neumann_rhs(neu)
            e.set_neumann_rhs(neu[: case.n_u])
        if step == 0:
            o.assemble_first(); e.assemble_first()
            o.precond_init(ptype); e.precond_init()
            x = case.random_state()
            xu, xp = x[: case.n_u], x[case.n_u:]
            assert T.rel_l2(e.ilu_apply(0, xu), o.ilu_apply(0, xu)) < 1e-11
            assert T.rel_l2(e.ilu_apply(1, xp), o.ilu_apply(1, xp)) < 1e-11
        else:
            o.assemble_step(); e.assemble_step()
        rc, its_o, _ = o.solve_step(ptype)
        its_e, _, _ = e.solve_step()
        assert rc == 0 and its_e == its_o, (step, its_e, its_o)
        assert T.rel_l2(e.get_solution()[: case.n_u], o.array("sol_owned", case.N)[: case.n_u]) < FIELD_TOL, step


@pytest.mark.parametrize("ordering", [1, 2])
@pytest.mark.parametrize("case_name,ptype", [("cyl2d", "asimple"), ("box3d", "yosida"), ("cyl3d", "yosida"),
                                             ("cube", "yosida")])
def test_batched_gram_schmidt_mode(case_name, ptype, ordering):
    """Throughput mode of the Krylov solvers (orthogonalisation = 1): classical Gram-Schmidt with all
    coefficients of an Arnoldi step from one fused multi-dot (inner solves: deal.II's loss test on
    every vector; outer solve: two passes), on top of the multicolour ILU(0).  Checked against the
    oracle running the same variant: same iteration counts, fields within 1e-8."""
    case = T.Case(case_name)
    o, e = case.oracle(), case.engine(precond_type=ptype, ilu_ordering=ordering, orthogonalisation=1)
    o.set_ilu_order(*_oracle_order(case, e))
    o.set_orthogonalisation(1)
    rows, vals = case.bc(0.0)
    o.set_dirichlet(rows, vals)
    e.set_dirichlet(rows)
    x0 = case.initial()
    o.set_solution(x0)
    e.set_solution(x0)
    t = 0.0
    for step in range(3):
        t += case.dt
        rows, vals = case.bc(t if case.variant == "conv" else 2.0 + t)
        o.set_dirichlet_values(vals)
        e.set_dirichlet_values(vals)
        if case.variant == "conv":
            neu = case.neumann(t - case.dt)
            o.set_neumann_rhs(neu)
            e.set_neumann_rhs(neu[: case.n_u])
        if step == 0:
            o.assemble_first(); e.assemble_first()
        else:
            o.assemble_step(); e.assemble_step()
        rc, its_o, _ = o.solve_step(ptype)
        its_e, _, _ = e.solve_step()
        assert rc == 0 and its_e == its_o, (step, its_e, its_o)
        assert e.stat("n_inner_F") == o.stat("n_inner_F")
        xo, xe = o.array("sol_owned", case.N), e.get_solution()
        assert T.rel_l2(xe[: case.n_u], xo[: case.n_u]) < FIELD_TOL, step


@pytest.mark.parametrize("lanes", [1, 4])
@pytest.mark.parametrize("case_name,ptype", [("cyl3d", "yosida"), ("cyl2d", "asimple")])
def test_chunked_multicolour_ordering(case_name, ptype, lanes, monkeypatch):
    """Chunk-major multicolour ordering (rows sorted by (chunk, colour)): the L2-blocking variant of the
    throughput mode, here forced to tiny chunks.  Still an exact ILU(0) of the permuted matrix: the
    oracle factorising in the same order gives the same preconditioner and the same iterations."""
    monkeypatch.setenv("NSB_ILU_CHUNK", "97")
    monkeypatch.setenv("NSB_SELL_LANES", str(lanes))  # both SELL-32 layouts: one / four lanes per row
    case = T.Case(case_name)
    o, e = case.oracle(), case.engine(precond_type=ptype, ilu_ordering=1)
    assert e.stat("levels_F_fwd") > 40  # many (chunk, colour) groups
    o.set_ilu_order(*_oracle_order(case, e))
    rows, vals = case.bc(2.0)
    o.set_dirichlet(rows, vals)
    e.set_dirichlet(rows)
    e.set_dirichlet_values(vals)
    x0 = 0.1 * case.random_state()
    for side in (o, e):
        side.set_solution(x0)
        side.assemble_first()
    o.precond_init(ptype); e.precond_init()
    x = case.random_state()
    xu, xp = x[: case.n_u], x[case.n_u:]
    assert T.rel_l2(e.ilu_apply(0, xu), o.ilu_apply(0, xu)) < 1e-11
    assert T.rel_l2(e.ilu_apply(1, xp), o.ilu_apply(1, xp)) < 1e-11
    Fx = T.oracle_blocks(o, "sys")["F"] @ xu
    assert T.rel_l2(e.block_vmult("F", xu), Fx) < 1e-13
    rc, its_o, _ = o.solve_step(ptype)
    its_e, _, _ = e.solve_step()
    assert rc == 0 and its_e == its_o
    assert T.rel_l2(e.get_solution()[: case.n_u], o.array("sol_owned", case.N)[: case.n_u]) < FIELD_TOL
