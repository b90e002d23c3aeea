"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py): the oracle must keep
reproducing them (CPU), and the CUDA path must match the frozen numbers through the C ABI (GPU)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import helpers as T

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("cyl2d", "asimple"), ("box3d", "yosida"), ("cube", "yosida")]


def _drive(case, side, ptype, is_oracle, gold):
    rows, vals = case.bc(0.0)
    if is_oracle:
        side.set_dirichlet(rows, vals)
    else:
        side.set_dirichlet(rows)
    side.set_solution(case.initial())
    t = 0.0
    for step in range(2):
        t += case.dt
        rows, vals = case.bc(t if case.variant == "conv" else 2.0 + t)
        side.set_dirichlet_values(vals)
        if case.variant == "conv":
            neu = case.neumann(t - case.dt)
            side.set_neumann_rhs(neu if is_oracle else neu[: case.n_u])
        (side.assemble_first if step == 0 else side.assemble_step)()
        Fg = sp.csr_matrix((gold[f"F_data_{step}"], gold["F_indices"], gold["F_indptr"]), shape=(case.n_u, case.n_u))
        F = T.oracle_blocks(side, "sys")["F"] if is_oracle else side.matrix("system", "F")
        assert T.entry_error(F, Fg) < 1e-12, step
        rhs = side.array("rhs", case.N) if is_oracle else side.get_rhs()
        assert np.abs(rhs - gold[f"rhs_{step}"]).max() < 1e-12 * np.abs(gold[f"rhs_{step}"]).max()
        its = side.solve_step(ptype)[1] if is_oracle else side.solve_step()[0]
        assert its == int(gold[f"its_{step}"]), (step, its)
        x = side.array("sol_owned", case.N) if is_oracle else side.get_solution()
        assert T.rel_l2(x[: case.n_u], gold[f"sol_{step}"][: case.n_u]) < 1e-8, step


@pytest.mark.parametrize("case_name,ptype", CASES)
def test_oracle_reproduces_golden(case_name, ptype):
    gold = np.load(os.path.join(GOLD, f"{case_name}_{ptype}.npz"))
    case = T.Case(case_name)
    assert (case.n_u, case.n_p) == (int(gold["n_u"]), int(gold["n_p"]))
    _drive(case, case.oracle(), ptype, True, gold)


@pytest.mark.gpu
@pytest.mark.parametrize("case_name,ptype", CASES)
def test_cuda_path_reproduces_golden(case_name, ptype):
    gold = np.load(os.path.join(GOLD, f"{case_name}_{ptype}.npz"))
    case = T.Case(case_name)
    _drive(case, case.engine(precond_type=ptype), ptype, False, gold)


# ---------------------------------------------------------------------------------------------------------
# DFG benchmark 2D-3 (Schaefer & Turek 1996) = the reference's 2D driver with its own literals (main2D.cpp:
# test case 2, T = 8, dt = 0.01): the only externally published numbers for what this path computes.
# tests/golden/dfg2d3_s<k>.npz hold the oracle's (t, outer iterations, c_D, c_L) history on cylinder2d(k).
DFG_CD_MAX, DFG_T_CD_MAX, DFG_CL_MAX, DFG_DP = (2.93, 2.97), 3.93, (0.47, 0.49), (-0.115, -0.105)
DFG_REPLAY_STEPS = 25


def _dfg(s):
    path = os.path.join(GOLD, f"dfg2d3_s{s}.npz")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated (tests/golden/make_dfg2d3.py {s})")
    return np.load(path)


@pytest.mark.parametrize("s,tol", [(2, 0.05), (4, 0.02), (6, 0.01)])
def test_dfg_2d3_drag_maximum_against_the_published_interval(s, tol):
    """c_D,max of the stored oracle history lies in the published interval [2.93, 2.97] widened by the mesh
    tolerance (10 k DoF: 5 %, 41 k DoF: 2 %, 91 k DoF: 1 %; P2-P1 converges from below: 2.848, 2.932, 2.943) and is
    reached at the published time t = 3.93 +- 0.03.  Lift and Delta P(8 s) are only bounded: they converge with the
    mesh to c_L,max = 0.37 and Delta P = -0.096 against the published 0.47-0.49 and -0.115..-0.105, which are
    time-converged values -- the reference's first-order semi-implicit scheme at dt = 0.01 damps the vortex shedding."""
    g = _dfg(s)
    h = g["history"]
    assert h.shape[0] == 800 and abs(h[-1, 0] - 8.0) < 1e-9
    cd_max, t_cd = h[:, 2].max(), h[h[:, 2].argmax(), 0]
    assert DFG_CD_MAX[0] * (1 - tol) <= cd_max <= DFG_CD_MAX[1] * (1 + tol), cd_max
    assert abs(t_cd - DFG_T_CD_MAX) <= 0.03, t_cd
    assert 0.0 < h[:, 3].max() <= DFG_CL_MAX[1] * 1.05
    dp = float(g["pressure_difference"][0])
    if np.isfinite(dp):  # Delta P(8 s) between the front and the back of the cylinder: depends on the phase and
        # amplitude of the vortex shedding (c_L,max 0.20 / 0.36 / 0.37 against 0.48, see above) -> 15 % margin
        assert DFG_DP[0] * 1.15 <= dp <= DFG_DP[1] * 0.85, dp


def _dfg_replay(s, make_side, is_oracle):
    """First steps of the benchmark run through `side`; returns [(its, c_D, c_L)]."""
    import sys

    sys.path.insert(0, GOLD)
    import make_dfg2d3 as M
    from navierstokes_project_nm4pde_b200 import HostMesh, NavierStokes, gauss_simplex
    from navierstokes_project_nm4pde_b200 import problem as P

    mesh = HostMesh.cylinder2d(s)
    prob = NavierStokes(mesh, "2d", T=M.T_END, deltat=M.DT, test_case=M.TEST_CASE)
    out = []
    if is_oracle:
        return M.run(s, DFG_REPLAY_STEPS, log=False)["history"][:, 1:]
    prob.setup()
    e = prob.engine
    e.set_solution(np.zeros(prob.N))
    tm = 0.0
    for k in range(DFG_REPLAY_STEPS):
        tm += M.DT
        e.set_dirichlet_values(prob.dirichlet_values(tm))
        (e.assemble_first if k == 0 else e.assemble_step)(tm)
        its = e.solve_step()[0]
        f = e.compute_forces(1.0)
        mv = P.mean_velocity(2, tm, M.TEST_CASE)
        out.append((its, 2.0 * f[0] / (mv ** 2 * 0.1), 2.0 * f[1] / (mv ** 2 * 0.1)))
    return np.array(out)


def test_oracle_reproduces_the_dfg_history():
    g = _dfg(2)
    got = _dfg_replay(2, None, True)
    want = g["history"][:DFG_REPLAY_STEPS, 1:]
    assert np.array_equal(got[:, 0], want[:, 0])
    assert np.allclose(got[:, 1:], want[:, 1:], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
def test_cuda_path_reproduces_the_dfg_history():
    """The reference's time loop on the device (replay mode) against the frozen benchmark history: identical outer
    iteration counts, drag / lift coefficients within the north-star tolerance 1e-6 (relative to the drag scale)."""
    g = _dfg(2)
    got = _dfg_replay(2, None, False)
    want = g["history"][:DFG_REPLAY_STEPS, 1:]
    assert np.array_equal(got[:, 0], want[:, 0])
    scale = np.abs(want[:, 1]).max()
    assert np.abs(got[:, 1:] - want[:, 1:]).max() < 1e-6 * scale
