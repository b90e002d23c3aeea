"""Setup fingerprints (nsb_debug_setup_fingerprint) for several meshes / orderings, with timing.
python scripts/setup_fingerprint.py [out.json] [--big | --huge]   (NSB_VERBOSE=2: phase timings; NSB_TREE: another checkout)"""
import ctypes as C, hashlib, json, os, sys, time
sys.path.insert(0, __import__("os").environ.get("NSB_TREE", __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
import numpy as np
from navierstokes_project_nm4pde_b200 import HostMesh, HostDofs, _lib
from navierstokes_project_nm4pde_b200._lib import dptr, iptr

def fingerprint(mesh, o1, o2, ghosts=False):
    d = HostDofs(mesh)
    nuo, npo = (d.dim * (2 * d.n_nodes // 3), 2 * d.n_p // 3) if ghosts else (d.n_u, d.n_p)
    cc, cd = d.cell_coords(copy=False), d.cell_dofs(copy=False)
    cap = 4096
    out = np.zeros(cap, np.uint64); n = C.c_int32(0)
    t0 = time.perf_counter()
    rc = _lib.lib().nsb_debug_setup_fingerprint(d.dim, d.n_cells, dptr(cc), iptr(cd), d.n_u, d.n_p, nuo, npo, o1, o2,
                                                out.ctypes.data_as(C.POINTER(C.c_uint64)), cap, iptr(np.zeros(1, np.int32)) if False else C.byref(n))
    dt = time.perf_counter() - t0
    assert rc == 0, rc
    assert n.value <= cap
    return hashlib.sha256(out[: n.value].tobytes()).hexdigest()[:16], n.value, dt

cases = {"cyl3d(1,3)": lambda: HostMesh.cylinder3d(1, 3), "cyl3d(2,8)": lambda: HostMesh.cylinder3d(2, 8),
         "cyl2d(1)": lambda: HostMesh.cylinder2d(1), "cyl2d(8)": lambda: HostMesh.cylinder2d(8), "cube(5)": lambda: HostMesh.cube(5)}
if "--big" in sys.argv:
    cases["cyl3d(4,16)"] = lambda: HostMesh.cylinder3d(4, 16)
    cases["cyl2d(16)"] = lambda: HostMesh.cylinder2d(16)
if "--huge" in sys.argv:
    cases = {"cyl3d(8,40)": lambda: HostMesh.cylinder3d(8, 40)}
res = {}
for k, f in cases.items():
    m = f()
    for (o1, o2) in ((0, -1), (1, -1), (2, -1), (2, 1), (1, 2), (3, -1)):
        if "--huge" in sys.argv and (o1, o2) not in ((2, 1), (1, -1)):
            continue
        fp, n, dt = fingerprint(m, o1, o2)
        res[f"{k}/{o1}/{o2}"] = [fp, n]
        print(f"{k} ordering {o1}/{o2}: {fp} ({n} uploads) {dt:.2f}s", flush=True)
        if o1 != 3 and "--huge" not in sys.argv:
            fp, n, dt = fingerprint(m, o1, o2, True)
            res[f"{k}/{o1}/{o2}/g"] = [fp, n]
out = [a for a in sys.argv[1:] if not a.startswith("--")]
if out:
    json.dump(res, open(out[0], "w"), indent=1, sort_keys=True)
