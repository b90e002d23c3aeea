"""The host side of setup (nsb_set_mesh + nsb_finalize_setup: sparsity patterns, scatter map of the step assembly,
SELL / block-SELL storage, the ILU orderings and their packed factors) without a GPU: nsb_debug_setup_fingerprint
runs the very same code on a handle without device state and hashes every array that would be uploaded.  The
hashes are pinned in tests/golden/setup_fingerprints.json (generated before the setup code was parallelised, kept
bit-identical since), so a change of the host code that alters any device data structure shows up here, on the
CPU, and not only as a wrong answer on the GPU."""
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from navierstokes_project_nm4pde_b200 import HostDofs, HostMesh, _lib
from navierstokes_project_nm4pde_b200._lib import dptr, iptr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "setup_fingerprints.json")
CASES = {"cyl3d(1,3)": lambda: HostMesh.cylinder3d(1, 3), "cyl3d(2,8)": lambda: HostMesh.cylinder3d(2, 8),
         "cyl2d(1)": lambda: HostMesh.cylinder2d(1), "cyl2d(8)": lambda: HostMesh.cylinder2d(8),
         "cube(5)": lambda: HostMesh.cube(5)}
# (ilu_ordering, ilu_ordering_schur): natural, point multicolour, block multicolour, the two mixed pairs bench.py
# uses (3D at 19.9 M DoF: 2 / 1; 2D: 1 / 2) and the subdomain ordering
ORDERINGS = ((0, -1), (1, -1), (2, -1), (2, 1), (1, 2), (3, -1))


def fingerprint(mesh, o1, o2, n_owned=None):
    d = HostDofs(mesh)
    cc, cd = d.cell_coords(copy=False), d.cell_dofs(copy=False)
    cap = 4096
    out, n = np.zeros(cap, np.uint64), C.c_int32(0)
    nu_o, np_o = (d.n_u, d.n_p) if n_owned is None else n_owned
    rc = _lib.lib().nsb_debug_setup_fingerprint(d.dim, d.n_cells, dptr(cc), iptr(cd), d.n_u, d.n_p, nu_o, np_o, o1, o2,
                                                out.ctypes.data_as(C.POINTER(C.c_uint64)), cap, C.byref(n))
    assert rc == 0 and 0 < n.value <= cap
    return [hashlib.sha256(out[: n.value].tobytes()).hexdigest()[:16], int(n.value)]


def owned_prefix(mesh):
    """A subdomain-like input: the first two thirds of the nodes / pressure DoFs are owned rows, the rest only occur
    as (ghost) columns -- the layout nsb_set_mesh gets from every rank of a multi-rank run."""
    d = HostDofs(mesh)
    return d.dim * (2 * d.n_nodes // 3), 2 * d.n_p // 3


@pytest.mark.parametrize("name", sorted(CASES))
def test_setup_structures_match_the_pinned_fingerprints(name):
    with open(GOLDEN) as f:
        golden = json.load(f)
    mesh = CASES[name]()
    for o1, o2 in ORDERINGS:
        assert fingerprint(mesh, o1, o2) == golden[f"{name}/{o1}/{o2}"], (name, o1, o2)
        if o1 != 3:  # (the subdomain ordering needs support points of owned rows only: single-rank option)
            assert fingerprint(mesh, o1, o2, owned_prefix(mesh)) == golden[f"{name}/{o1}/{o2}/ghosts"], (name, o1, o2)


# real subdomains: (mesh, ranks); every rank's local problem from distributed.build_local_problem
LOCAL_CASES = (("cyl3d(2,8)", 2), ("cyl2d(8)", 3))
LOCAL_ORDERINGS = ((1, -1), (2, 1), (1, 2))


def fingerprint_local(mesh, nranks, rank, o1, o2):
    """What rank `rank` of `nranks` hands to nsb_set_mesh (owned + ghost cells, owned DoFs first)."""
    from navierstokes_project_nm4pde_b200.distributed import build_local_problem

    d = HostDofs(mesh)
    loc = build_local_problem(d.dim, d.cell_dofs(copy=False), d.cell_coords(copy=False), d.n_nodes, d.n_p,
                              mesh.partition(nranks), nranks, rank)
    cc = np.ascontiguousarray(loc["cell_coords"], dtype=np.float64)
    cd = np.ascontiguousarray(loc["cell_dofs"], dtype=np.int32)
    cap = 4096
    out, n = np.zeros(cap, np.uint64), C.c_int32(0)
    rc = _lib.lib().nsb_debug_setup_fingerprint(d.dim, cd.shape[0], dptr(cc), iptr(cd), d.dim * loc["node_gid"].size,
                                                loc["p_gid"].size, d.dim * loc["n_nodes_owned"], loc["n_p_owned"], o1, o2,
                                                out.ctypes.data_as(C.POINTER(C.c_uint64)), cap, C.byref(n))
    assert rc == 0 and 0 < n.value <= cap
    return [hashlib.sha256(out[: n.value].tobytes()).hexdigest()[:16], int(n.value)]


@pytest.mark.parametrize("name,nranks", LOCAL_CASES)
def test_subdomain_structures_match_the_pinned_fingerprints(name, nranks):
    """The same pin for what every rank of a multi-rank run builds (block-Jacobi ILU of the owned rows, ghost
    columns dropped from the factors, ghost rows absent from the scatter map)."""
    with open(GOLDEN) as f:
        golden = json.load(f)
    mesh = CASES[name]()
    for rank in range(nranks):
        for o1, o2 in LOCAL_ORDERINGS:
            assert fingerprint_local(mesh, nranks, rank, o1, o2) == golden[f"local/{name}/{nranks}/{rank}/{o1}/{o2}"]


def test_setup_structures_do_not_depend_on_the_thread_count():
    """The parallel host code must build the same arrays on 1 and on 3 threads (OpenMP schedules differ)."""
    code = ("import json, sys; sys.path[:0] = [%r, %r]; from test_setup_fingerprint import *; m = CASES['cyl3d(2,8)']();"
            "print(json.dumps([fingerprint(m, o1, o2) for o1, o2 in ORDERINGS]))" % (ROOT, os.path.join(ROOT, "tests")))
    outs = []
    for nt in ("1", "3"):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT,
                           env=dict(os.environ, OMP_NUM_THREADS=nt))
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert outs[0] == outs[1]


def test_dry_run_is_not_a_cpu_fallback():
    """The dry run leaves no state behind: creating an engine without a device still fails loudly."""
    fingerprint(CASES["cyl2d(1)"](), 1, -1)
    h = C.c_void_p()
    if _lib.lib().nsb_device_count() > 0:
        pytest.skip("a GPU is present")
    assert _lib.lib().nsb_create(C.byref(h), 2, 0, 1, 0, None) != 0


def test_bad_arguments():
    n = C.c_int32(0)
    assert _lib.lib().nsb_debug_setup_fingerprint(4, 1, None, None, 3, 1, 3, 1, 0, -1, None, 0, C.byref(n)) != 0
    m = CASES["cyl2d(1)"]()
    d = HostDofs(m)
    bad = d.cell_dofs().copy()
    bad[0, 0] += 1  # velocity DoFs no longer node-interleaved
    rc = _lib.lib().nsb_debug_setup_fingerprint(2, d.n_cells, dptr(d.cell_coords(copy=False)), iptr(bad), d.n_u, d.n_p, d.n_u,
                                                d.n_p, 0, -1, None, 0, C.byref(n))
    assert rc != 0
