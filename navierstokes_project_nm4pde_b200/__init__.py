"""B200-native engine for the per-timestep hot path of lelecaruso/NavierStokes_Project_NM4PDE.

Only the hot path lives here: cell-loop assembly of the P2-P1 block system and the
block-preconditioned GMRES solve, as hand-written sm_100a kernels behind the C ABI of
`include/nsb.h` (`csrc/`), plus the host-side mirror of the reference's `NavierStokes`
interface (`problem.py`) that the tests and `bench.py` drive.  No CPU fallback exists.
"""
from .engine import Engine, HostDofs, HostMesh  # noqa: F401
from .problem import NavierStokes  # noqa: F401
from .quadrature import gauss_simplex  # noqa: F401

__all__ = ["Engine", "HostDofs", "HostMesh", "NavierStokes", "gauss_simplex"]
