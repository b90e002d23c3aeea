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
