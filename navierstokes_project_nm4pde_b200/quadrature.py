"""QGaussSimplex<dim>(n_points_1D) tables used by the engine's host side.

The reference builds `QGaussSimplex<dim>(fe->degree + 1)` = `QGaussSimplex(3)`
(Navier-Stokes/src/NavierStokes2D.cpp:45).  deal.II >= 9.4 forwards that to the
Witherden-Vincent degree-5 rules (7 points in 2D, 14 in 3D); deal.II 9.3.x hard-codes a
7-point 2D table with truncated constants and a 10-point degree-3 rule in 3D.
`Convergence3D.cpp:772` needs QGaussSimplex<3>(4), which only exists from 9.4 on, so "wv" is the
default.  The tables are data: any rule with <= 16 points can be passed to `nsb_set_quadrature`.
"""
from __future__ import annotations

import math

import numpy as np


def _sym3(a):
    return [(a, a), (1.0 - 2.0 * a, a), (a, 1.0 - 2.0 * a)]


def gauss_simplex(dim: int, rule: str = "wv"):
    """Return (points[nq, dim], weights[nq]); weights sum to the reference simplex volume."""
    if dim == 1:  # QGauss<1>(3) on [0, 1]
        g = math.sqrt(0.6)
        return np.array([[0.5 - 0.5 * g], [0.5], [0.5 + 0.5 * g]]), np.array([5.0, 8.0, 5.0]) / 18.0
    if dim == 2 and rule == "wv":
        r = math.sqrt(15.0)
        pts = [(1.0 / 3.0, 1.0 / 3.0)] + _sym3((6.0 - r) / 21.0) + _sym3((6.0 + r) / 21.0)
        w = [0.1125] + [(155.0 - r) / 2400.0] * 3 + [(155.0 + r) / 2400.0] * 3
        return np.array(pts), np.array(w)
    if dim == 2 and rule == "dealii93":
        pts = [(0.3333333333330, 0.3333333333330), (0.7974269853530, 0.1012865073230),
               (0.1012865073230, 0.7974269853530), (0.1012865073230, 0.1012865073230),
               (0.0597158717898, 0.4701420641050), (0.4701420641050, 0.0597158717898),
               (0.4701420641050, 0.4701420641050)]
        w = [0.1125] + [0.0629695902725] * 3 + [0.0661970763945] * 3
        return np.array(pts), np.array(w)
    if dim == 3 and rule == "wv":
        groups = ((0.31088591926330060980, 0.11268792571801585080 / 6.0),
                  (0.092735250310891226402, 0.073493043116361949544 / 6.0))
        pts, w = [], []
        for a, ww in groups:
            b = 1.0 - 3.0 * a
            pts += [(a, a, a), (b, a, a), (a, b, a), (a, a, b)]
            w += [ww] * 4
        c = 0.045503704125649649492
        d = 0.5 - c
        pts += [(c, c, d), (c, d, c), (d, c, c), (c, d, d), (d, c, d), (d, d, c)]
        w += [0.042546020777081466438 / 6.0] * 6
        return np.array(pts), np.array(w)
    if dim == 3 and rule == "dealii93":
        a, b = 0.5684305841968444, 0.1438564719343852
        pts = [(a, b, b), (b, b, b), (b, b, a), (b, a, b), (0.0, 0.5, 0.5), (0.5, 0.0, 0.5), (0.5, 0.5, 0.0),
               (0.5, 0.0, 0.0), (0.0, 0.5, 0.0), (0.0, 0.0, 0.5)]
        w = [0.2177650698804054 / 6.0] * 4 + [0.0214899534130631 / 6.0] * 6
        return np.array(pts), np.array(w)
    raise ValueError(f"no QGaussSimplex table for dim={dim}, rule={rule!r}")
