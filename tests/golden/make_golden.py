"""Generates tests/golden/*.npz: small fixed-input / fixed-output vectors of the hot path.

The reference ships no golden vectors and cannot be built here, so these fixtures
are produced by this repo's CPU oracle (oracle/ns_oracle.c, itself pinned by the sympy-exact and
analytic checks of tests/test_oracle_pins.py) and committed so that (a) the oracle cannot drift
silently and (b) the CUDA path is also checked against frozen numbers, not only against a checker
built from the same tree.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as T  # noqa: E402

CASES = {"cyl2d": "asimple", "box3d": "yosida", "cube": "yosida"}


def run(case_name, ptype):
    case = T.Case(case_name)
    o = case.oracle()
    rows, vals = case.bc(0.0)
    o.set_dirichlet(rows, vals)
    o.set_solution(case.initial())
    out = {"n_u": case.n_u, "n_p": case.n_p, "dt": case.dt}
    t = 0.0
    for step in range(2):
        t += case.dt
        rows, vals = case.bc(t if case.variant == "conv" else 2.0 + t)
        o.set_dirichlet_values(vals)
        if case.variant == "conv":
            o.set_neumann_rhs(case.neumann(t - case.dt))
        if step == 0:
            o.assemble_first()
        else:
            o.assemble_step()
        F = T.oracle_blocks(o, "sys")["F"]
        out[f"F_data_{step}"] = F.data.copy()
        if step == 0:
            out["F_indptr"], out["F_indices"] = F.indptr.copy(), F.indices.copy()
        out[f"rhs_{step}"] = o.array("rhs", case.N).copy()
        rc, its, _ = o.solve_step(ptype)
        assert rc == 0
        out[f"its_{step}"] = its
        out[f"sol_{step}"] = o.array("sol_owned", case.N).copy()
    return out


if __name__ == "__main__":
    for name, ptype in CASES.items():
        d = run(name, ptype)
        np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"{name}_{ptype}.npz"), **d)
        print(name, ptype, "its", d["its_0"], d["its_1"], "N", d["n_u"] + d["n_p"])
