"""Generates tests/golden/dfg2d3_s<k>.npz: the reference's 2D driver (main2D.cpp: test case 2 = inflow
4 U_m y (H - y) sin(pi t / 8) / H^2 with U_m = 1.5, T = 8, dt = 0.01, aSIMPLE) run through the CPU oracle on the
generated cylinder mesh `cylinder2d(s)`, with drag / lift coefficients from the oracle's restatement of
compute_forces (NavierStokes2D.cpp:752-859) and the pressure difference of compute_pressure_difference (:862-936).

This is the DFG benchmark 2D-3 of Schaefer & Turek, "Benchmark computations of laminar flow around a cylinder"
(Notes on Numerical Fluid Mechanics 52, 1996), whose published reference intervals are
    c_D,max in [2.93, 2.97]   (reached at t = 3.93),   c_L,max in [0.47, 0.49]   (t = 5.69),
    Delta P(t = 8 s) in [-0.115, -0.105].
They are the only externally published numbers for what this code path computes (the reference repository ships no
results), so the stored histories pin the oracle -- assembly, boundary rows, preconditioned solve, time loop and force
integrals together -- against an independent source; tests/test_oracle_pins.py checks the stored maxima against the
intervals with a tolerance for the mesh (s = 2: 10 k DoF, s = 4: 41 k DoF, s = 6: 91 k DoF) and replays the first steps.
Stored results: c_D,max 2.848 / 2.932 / 2.943 at t = 3.93 / 3.94 / 3.94; c_L,max 0.198 / 0.365 / 0.372;
Delta P(8 s) -0.109 / -0.096 / -0.096.

    python tests/golden/make_dfg2d3.py 2        # ~2 minutes on 8 cores
    python tests/golden/make_dfg2d3.py 4        # ~20 minutes
    python tests/golden/make_dfg2d3.py 6        # ~35 minutes

Lift and Delta P converge with the mesh to values below the published ones because the published values are converged in
time and the reference's scheme is first order with dt = 0.01: with half the time step (not the reference's literal,
printed only, no fixture written)
    python tests/golden/make_dfg2d3.py 4 0.005  # ~8 minutes
gives c_D,max 2.932 (unchanged), c_L,max 0.439 (0.365 at dt = 0.01) and Delta P(8 s) = -0.1035 (-0.096).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from navierstokes_project_nm4pde_b200 import HostMesh, NavierStokes, gauss_simplex  # noqa: E402
from navierstokes_project_nm4pde_b200 import problem as P  # noqa: E402
from oracle import ns_ref as R  # noqa: E402

DT, T_END, TEST_CASE = 0.01, 8.0, 2  # main2D.cpp:7,21-22


def run(s, nsteps=None, log=True, dt=None):
    DT = dt or globals()["DT"]
    mesh = HostMesh.cylinder2d(s)
    prob = NavierStokes(mesh, "2d", T=T_END, deltat=DT, test_case=TEST_CASE)
    prob.setup_host()
    d = prob.dofs
    num = dict(dim=2, cell_dofs=d.cell_dofs(), N=d.N, n_u=d.n_u, n_p=d.n_p, dpc=d.dpc)
    o = R.Oracle(2, "2d", mesh.vertices, mesh.cells, num, R.system_pattern(num), 1e-3, DT)
    fc, fl = d.boundary_faces(3)
    xi, w = gauss_simplex(1)
    nsteps = nsteps or int(round(T_END / DT))
    o.set_solution(np.zeros(d.N))
    hist = np.zeros((nsteps, 4))
    dp, tm, t0 = np.nan, 0.0, time.time()
    for k in range(nsteps):
        tm += DT
        vals = prob.dirichlet_values(tm)
        if k == 0:
            o.set_dirichlet(prob._dir_rows, vals)
            o.assemble_first()
        else:
            o.set_dirichlet_values(vals)
            o.assemble_step()
        rc, its, _ = o.solve_step("asimple")
        assert rc == 0, (k, rc)
        x = o.array("sol_owned", d.N)
        f = o.compute_forces(x, fc, fl, xi, w, rho=1.0)
        mean_v = P.mean_velocity(2, tm, TEST_CASE)  # 1.0: the 2D class takes the constant mean for this case
        hist[k] = (tm, its, 2.0 * f[0] / (mean_v ** 2 * 0.1), 2.0 * f[1] / (mean_v ** 2 * 0.1))
        if k == nsteps - 2 or k == int(round(T_END / DT)) - 2:  # time == T - deltat (NavierStokes2D.cpp:735)
            pv = lambda pt: (lambda v: np.nan if v is None else v[2])(d.point_value(x, pt))  # noqa: E731
            # benchmark points (front / back of the cylinder, nudged off the polygonal surface); the reference's points
            dp = (pv([0.15 - 1e-9, 0.2]) - pv([0.25 + 1e-9, 0.2]), pv([0.45, 0.2]) - pv([0.55, 0.2]))
        if log and k % 100 == 99:
            print(f"step {k + 1}: {its} its, c_D {hist[k, 2]:.5f}, c_L {hist[k, 3]:.5f}, {time.time() - t0:.0f} s", flush=True)
    return dict(s=s, n_dofs=d.N, dt=DT, history=hist, pressure_difference=np.array(dp, dtype=float))


if __name__ == "__main__":
    s = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dt = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = run(s, dt=dt)
    h = out["history"]
    print(f"s={s}, {out['n_dofs']} DoF: c_D,max {h[:, 2].max():.4f} at t={h[h[:, 2].argmax(), 0]:.2f}, "
          f"c_L,max {h[:, 3].max():.4f} at t={h[h[:, 3].argmax(), 0]:.2f}, dP (0.15 / 0.25; 0.45 / 0.55) {out['pressure_difference']}")
    if dt is None or dt == DT:  # only the reference's literal time step is a fixture
        np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), f"dfg2d3_s{s}.npz"), **out)
