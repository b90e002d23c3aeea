"""CPU-only checks of the boundary and the host logic: the C-ABI library loads and exports every
symbol include/nsb.h declares, refuses to compute without a GPU (no CPU fallback), and the host
prerequisites (mesh generators, .msh I/O, DoF numbering, boundary lists) behave like the
reference's setup() (NavierStokes2D.cpp:2-157)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from navierstokes_project_nm4pde_b200 import Engine, HostDofs, HostMesh, NavierStokes, _lib
from oracle import ns_ref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R_SEED = 20240607


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nsb.h")).read()
    names = sorted(set(re.findall(r"\b(ns[bh]_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 60
    lib = _lib.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback():
    """Without a usable sm_100 device nsb_create fails loudly; nothing routes to the oracle."""
    if Engine.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.NsbError) as ei:
        Engine(2)
    assert "no CPU fallback" in str(ei.value)
    src = "".join(open(os.path.join(ROOT, "navierstokes_project_nm4pde_b200", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "navierstokes_project_nm4pde_b200")) if f.endswith(".py"))
    assert "oracle" not in src.replace("CPU oracle lives", "")  # the product never imports the checker


@pytest.mark.parametrize("gen,dim,ids", [(lambda: HostMesh.cylinder2d(1), 2, {0, 1, 2, 3}),
                                         (lambda: HostMesh.cylinder3d(1, 3), 3, {0, 1, 2, 3}),
                                         (lambda: HostMesh.cube(3), 3, {0, 1, 2, 3, 4, 5})])
def test_generators_and_numbering(gen, dim, ids):
    """Boundary ids as in the .geo scripts (Cylinder2D.geo:40-43, Cylinder3D.geo:126-129,
    mesh-cube.geo:16-21); DoF numbering = distribute_dofs + component_wise as restated by the
    oracle's independent Python numbering."""
    m = gen()
    assert m.dim == dim and set(np.unique(m.bface_ids).tolist()) == ids
    d = HostDofs(m)
    num = R.number_dofs(dim, m.vertices, m.cells)
    assert np.array_equal(num["cell_dofs"], d.cell_dofs())
    assert d.n_u == num["n_u"] and d.n_p == num["n_p"] == m.n_vertices
    # positive orientation of every cell (FEValues would otherwise see negative JxW)
    X = m.vertices[m.cells]
    J = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))
    assert (np.linalg.det(J) > 0).all()


@pytest.mark.parametrize("gen", [lambda: HostMesh.cylinder2d(2), lambda: HostMesh.cylinder3d(1, 4), lambda: HostMesh.cube(4),
                                 lambda: HostMesh.box(2, [3, 5], [0, 0], [1, 2])])
def test_boundary_faces_against_a_brute_force_search(gen):
    """Mesh::build_boundary (faces bucketed by their smallest vertex, buckets searched in parallel): the faces that
    belong to exactly one cell, in (cell, local face) order, vertices ascending -- against numpy's unique()."""
    m = gen()
    dim, cells = m.dim, m.cells
    faces, owner = [], []
    for c in range(cells.shape[0]):
        for f in range(dim + 1):  # local face f = the face opposite to local vertex f
            faces.append(sorted(int(v) for k, v in enumerate(cells[c]) if k != f))
            owner.append(c)
    faces, owner = np.array(faces), np.array(owner)
    _, inv, cnt = np.unique(faces, axis=0, return_inverse=True, return_counts=True)
    single = cnt[inv.ravel()] == 1
    assert np.array_equal(m.bfaces, faces[single]) and np.array_equal(m.bface_cells, owner[single])
    # a closed surface: every (dim-2)-dimensional sub-face of the boundary is shared by exactly two boundary faces
    sub = np.concatenate([np.delete(m.bfaces, k, axis=1) for k in range(dim)])
    _, c2 = np.unique(np.sort(sub, axis=1), axis=0, return_counts=True)
    assert (c2 == 2).all()


def test_msh_roundtrip(tmp_path):
    """GridIn::read_msh replacement: Gmsh v2 ASCII written and read back keeps cells, vertices and ids."""
    m = HostMesh.cylinder2d(1)
    path = str(tmp_path / "c.msh")
    m.write_msh(path)
    r = HostMesh.read_msh(path)
    assert np.array_equal(m.cells, r.cells) and np.allclose(m.vertices, r.vertices, rtol=0, atol=1e-15)
    key = lambda mm: sorted((tuple(sorted(f.tolist())), int(i)) for f, i in zip(mm.bfaces, mm.bface_ids))  # noqa: E731
    assert key(m) == key(r)


def test_dirichlet_lists_follow_the_reference_order():
    """interpolate_boundary_values is called for the inlet first and the walls second, the second
    call overwriting shared nodes with zero (NavierStokes2D.cpp:328-353)."""
    p = NavierStokes(HostMesh.cylinder2d(1), "2d", T=1.0, deltat=0.01, test_case=3)
    p.setup_host()
    v = p.dirichlet_values(1.0).reshape(-1, 2)
    xyz = p._dir_xyz
    inlet = np.isclose(xyz[:, 0], 0.0)
    corner = inlet & (np.isclose(xyz[:, 1], 0.0) | np.isclose(xyz[:, 1], 0.41))
    assert (v[~inlet] == 0).all() and (v[corner] == 0).all() and (v[inlet & ~corner, 0] > 0).all()
    assert (v[:, 1] == 0).all()
    assert len(np.unique(p._dir_rows)) == len(p._dir_rows)


@pytest.mark.parametrize("gen,dim", [(lambda: HostMesh.cylinder2d(1), 2), (lambda: HostMesh.box(3, (3, 2, 2), (0, 0, 0), (1.0, 0.41, 0.41)), 3)])
def test_point_value_reproduces_p2_p1_fields(gen, dim):
    """VectorTools::point_value replacement (compute_pressure_difference, NavierStokes2D.cpp:862-936):
    a quadratic velocity and a linear pressure are reproduced exactly anywhere inside the mesh; a
    point outside (here: inside the cylinder hole / outside the box) is reported as not available."""
    m = gen()
    d = HostDofs(m)
    X, P = d.node_xyz, d.p_xyz
    quad = lambda x: 1.0 + x[..., 0] * x[..., 1] - 2.0 * x[..., 0] ** 2 + (x[..., -1] ** 2 if dim == 3 else 0.0)  # noqa: E731
    lin = lambda x: 0.3 - 1.7 * x[..., 0] + 0.9 * x[..., 1] + (0.5 * x[..., -1] if dim == 3 else 0.0)  # noqa: E731
    u = np.stack([(k + 1) * quad(X) for k in range(dim)], axis=1).ravel()
    sol = np.concatenate([u, lin(P)])
    pts = [[0.45, 0.2], [0.55, 0.2], [1.234, 0.111]] if dim == 2 else [[0.45, 0.2, 0.205], [0.9, 0.4, 0.01]]
    for x in pts:
        x = np.array(x)
        v = d.point_value(sol, x)
        assert v is not None
        assert np.allclose(v[:dim], [(k + 1) * quad(x) for k in range(dim)], rtol=0, atol=1e-12)
        assert abs(v[dim] - lin(x)) < 1e-12
    outside = [0.2, 0.2] if dim == 2 else [1.5, 0.2, 0.2]
    assert d.point_value(sol, np.array(outside)) is None


def test_vtu_writer(tmp_path):
    """DataOut::write_vtu stand-in: a well-formed UnstructuredGrid with the mesh's points / cells and
    the vertex values of velocity and pressure."""
    import xml.etree.ElementTree as ET

    m = HostMesh.cylinder2d(1)
    d = HostDofs(m)
    sol = np.concatenate([np.tile([1.5, -0.5], d.n_nodes), 2.0 + d.p_xyz[:, 0]])
    path = tmp_path / "output-navier-stokes-2D_001.vtu"
    d.write_vtu(sol, path)
    root = ET.parse(path).getroot()
    piece = root.find("UnstructuredGrid/Piece")
    assert int(piece.get("NumberOfPoints")) == m.n_vertices and int(piece.get("NumberOfCells")) == m.n_cells
    arrays = {a.get("Name"): np.array(a.text.split(), float) for a in piece.iter("DataArray") if a.get("Name")}
    assert np.array_equal(arrays["connectivity"].astype(int).reshape(-1, 3), m.cells)
    assert np.allclose(arrays["velocity"].reshape(-1, 3), [1.5, -0.5, 0.0])
    pts = np.array(piece.find("Points/DataArray").text.split(), float).reshape(-1, 3)
    assert np.allclose(pts[:, :2], m.vertices) and np.allclose(arrays["pressure"], 2.0 + m.vertices[:, 0])


MSH41 = """$MeshFormat
4.1 0 8
$EndMeshFormat
$Entities
4 4 1 0
1 0 0 0 0
2 1 0 0 0
3 1 1 0 0
4 0 1 0 0
1 0 0 0 1 0 0 1 0 2 1 -2
2 1 0 0 1 1 0 1 1 2 2 -3
3 0 1 0 1 1 0 1 2 2 3 -4
4 0 0 0 0 1 0 1 3 2 4 -1
1 0 0 0 1 1 0 1 9 4 1 2 3 4
$EndEntities
$Nodes
1 4 1 4
2 1 0 4
1
2
3
4
0 0 0
1 0 0
1 1 0
0 1 0
$EndNodes
$Elements
5 6 1 6
1 1 1 1
1 1 2
1 2 1 1
2 2 3
1 3 1 1
3 3 4
1 4 1 1
4 4 1
2 1 2 2
5 1 2 3
6 1 3 4
$EndElements
"""


def test_msh_v41_reader_and_malformed_files(tmp_path):
    """GridIn::read_msh accepts Gmsh 4.1 ASCII: physical ids of the curves arrive through $Entities.
    Unreadable input is refused (NULL handle -> NsbError), never half-read."""
    p = tmp_path / "square41.msh"
    p.write_text(MSH41)
    m = HostMesh.read_msh(str(p))
    assert m.dim == 2 and m.n_cells == 2 and m.n_vertices == 4
    ids = {tuple(sorted(f.tolist())): int(i) for f, i in zip(m.bfaces, m.bface_ids)}
    assert ids == {(0, 1): 0, (1, 2): 1, (2, 3): 2, (0, 3): 3}
    X = m.vertices[m.cells]
    assert (np.linalg.det(np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))) > 0).all()
    d = HostDofs(m)
    assert (d.n_nodes, d.n_p, d.dpc) == (9, 4, 15)
    for bad, text in (("empty.msh", ""), ("binary.msh", "$MeshFormat\n2.2 1 8\n$EndMeshFormat\n"),
                      ("nocells.msh", "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n1\n1 0 0 0\n$EndNodes\n$Elements\n0\n$EndElements\n")):
        q = tmp_path / bad
        q.write_text(text)
        with pytest.raises(_lib.NsbError):
            HostMesh.read_msh(str(q))
    with pytest.raises(_lib.NsbError):
        HostMesh.read_msh(str(tmp_path / "does_not_exist.msh"))


def _hole_measure(m, dim):
    """Area (2D) / volume (3D) enclosed by the boundary faces with id 3, by the divergence theorem."""
    F = m.bfaces[m.bface_ids == 3]
    X = m.vertices
    if dim == 2:
        c = X[F].mean(axis=(0, 1))
        a = 0.0
        for e in F:
            p, q = X[e[0]] - c, X[e[1]] - c
            a += 0.5 * abs(p[0] * q[1] - p[1] * q[0])
        return a
    # the obstacle is a right prism (polygon x [0, H]) whose caps lie on the channel walls, so id 3 only
    # holds its lateral surface: the cones from the centroid to the lateral faces make up 2/3 of it
    c = X[F].mean(axis=(0, 1))
    v = 0.0
    for t in F:
        p, q, r = X[t[0]] - c, X[t[1]] - c, X[t[2]] - c
        v += abs(np.dot(p, np.cross(q, r))) / 6.0
    return 1.5 * v


@pytest.mark.parametrize("gen,dim", [(lambda: HostMesh.cylinder2d(1), 2), (lambda: HostMesh.cylinder3d(1, 3), 3)])
def test_boundary_forces_exact_for_linear_fields(gen, dim):
    """compute_forces face integrals (NavierStokes2D.cpp:752-859, NavierStokes3D.cpp:744-840) on the
    obstacle (id 3).  With the reference's n = -(outward normal of the fluid cell) the integral of
    -p n over the closed obstacle surface is -grad(p) * |hole| for a linear pressure; a constant
    velocity gradient contributes nu G (closed integral of n) = 0 in the 2D formula."""
    m = gen()
    d = HostDofs(m)
    X, P = d.node_xyz, d.p_xyz
    beta, gamma = 0.7, -1.3
    p = 2.0 + beta * P[:, 0] + gamma * P[:, 1]
    if dim == 3:  # the cylinder spans the whole channel in z: its end caps lie on the walls (id 2), not on id 3
        u = np.zeros(d.n_u)
    else:
        u = np.stack([0.3 * X[:, 0] - 0.2 * X[:, 1], 0.5 * X[:, 0] + 0.1 * X[:, 1]], axis=1).ravel()
    drag, lift = d.boundary_forces(np.concatenate([u, p]), 3, nu=1e-3)
    hole = _hole_measure(m, dim)
    assert hole > 0
    if dim == 2:
        assert abs(drag + beta * hole) < 1e-12 and abs(lift + gamma * hole) < 1e-12
    else:
        # side surface only: the closed-surface identity holds for the x and y components because the
        # missing caps have normals along z
        assert abs(drag + beta * hole) < 1e-10 and abs(lift + gamma * hole) < 1e-10


@pytest.mark.parametrize("gen,dim,bs,leaf", [(lambda: HostMesh.cylinder3d(1, 3), 3, 3, 300), (lambda: HostMesh.cylinder2d(2), 2, 2, 200),
                                             (lambda: HostMesh.cylinder3d(1, 3), 3, 1, 150), (lambda: HostMesh.cube(4), 3, 3, 64)])
def test_subdomain_ilu_storage_cpu(gen, dim, bs, leaf):
    """ilu_ordering = 3 without a GPU (nsb_debug_sd_check): the multi-level subdomain ordering is a
    permutation whose colours are independent sets and whose part rows only couple with their own part
    or with rows of other levels, and the packed per-part streams (rounds, slice headers, 16-bit local
    columns, rings) reproduce plain forward / backward substitution through a host emulation of the
    device kernels."""
    import ctypes as C

    import scipy.sparse as sp

    from navierstokes_project_nm4pde_b200 import _lib
    from navierstokes_project_nm4pde_b200._lib import dptr, iptr

    d = HostDofs(gen())
    cd = d.cell_dofs(copy=False)
    nv = dim + 1
    if bs == 1:  # pressure graph squared ~ pattern of B D^-1 B^T
        ids, n, xyz = cd[:, [v * (dim + 1) + dim for v in range(nv)]] - d.n_u, d.n_p, d.p_xyz
    else:
        ne = 3 if dim == 2 else 6
        cols = [v * (dim + 1) for v in range(nv)] + [nv * (dim + 1) + e * dim for e in range(ne)]
        ids, n, xyz = cd[:, cols] // dim, d.n_nodes, d.node_xyz
    k = ids.shape[1]
    A = sp.csr_matrix((np.ones(ids.size * k), (np.repeat(ids, k, axis=1).ravel(), np.tile(ids, (1, k)).ravel())), shape=(n, n))
    if bs == 1:
        A = (A @ A).tocsr()
    A.sum_duplicates(); A.sort_indices()
    rp, ci = A.indptr.astype(np.int32), A.indices.astype(np.int32)
    err, stats, order = C.c_double(0), np.zeros(8, np.int32), np.zeros(n, np.int32)
    xyz = np.ascontiguousarray(xyz)
    for levels in ([leaf, 0, 0], [leaf, leaf // 3, leaf // 3]):  # one level; three levels (separators cut into parts again)
        lv = np.array(levels, np.int32)
        rc = _lib.lib().nsb_debug_sd_check(n, iptr(rp), iptr(ci), dptr(xyz), dim, iptr(lv), 0, bs, C.byref(err), iptr(stats),
                                           iptr(order))
        assert rc == 0
        assert sorted(order.tolist()) == list(range(n))
        assert err.value < 1e-12, err.value
        assert stats[0] >= 2 and 0 < stats[1] <= n and stats[3] <= 65535
        assert stats[6] == (1 if levels[1] == 0 else 3) or stats[1] == n


def _graph_of(gen, dim, bs):
    """Node graph of F_s (bs = dim) or the two-ring pressure graph of B D^-1 B^T (bs = 1) of a generated mesh."""
    import scipy.sparse as sp

    d = HostDofs(gen())
    cd = d.cell_dofs(copy=False)
    nv = dim + 1
    if bs == 1:
        ids, n = cd[:, [v * (dim + 1) + dim for v in range(nv)]] - d.n_u, d.n_p
    else:
        ne = 3 if dim == 2 else 6
        cols = [v * (dim + 1) for v in range(nv)] + [nv * (dim + 1) + e * dim for e in range(ne)]
        ids, n = cd[:, cols] // dim, d.n_nodes
    k = ids.shape[1]
    A = sp.csr_matrix((np.ones(ids.size * k), (np.repeat(ids, k, axis=1).ravel(), np.tile(ids, (1, k)).ravel())), shape=(n, n))
    if bs == 1:
        A = (A @ A).tocsr()
    A.sum_duplicates(); A.sort_indices()
    return n, A.indptr.astype(np.int32), A.indices.astype(np.int32)


@pytest.mark.parametrize("gen,dim,bs", [(lambda: HostMesh.cylinder3d(1, 3), 3, 3), (lambda: HostMesh.cylinder2d(2), 2, 2),
                                        (lambda: HostMesh.cylinder3d(1, 3), 3, 1), (lambda: HostMesh.cube(4), 3, 3),
                                        (lambda: HostMesh.cylinder3d(2, 8), 3, 3)])
def test_block_multicolour_ilu_storage_cpu(gen, dim, bs):
    """ilu_ordering = 2 without a GPU (nsb_debug_bsell_check): the block multicolour ordering is a permutation whose
    blocks of one colour do not touch, and the packed block storage of both factors (passes of eight rows, staged-list
    indices, (pass, slot) -> row table, in-block cursors) reproduces plain forward / backward substitution through a
    host emulation of the sweep kernel -- with the whole outside list staged, with a capacity of 8 rows (the rest
    gathered by factor row, as the kernel does beyond NSB_BSELL_XCAP) and without staging (velocity block)."""
    import ctypes as C

    from navierstokes_project_nm4pde_b200 import _lib
    from navierstokes_project_nm4pde_b200._lib import iptr

    n, rp, ci = _graph_of(gen, dim, bs)
    err, stats, order = C.c_double(0), np.zeros(4, np.int32), np.zeros(n, np.int32)
    for xcap in (65535, 8, 0):
        rc = _lib.lib().nsb_debug_bsell_check(n, iptr(rp), iptr(ci), bs, xcap, C.byref(err), iptr(stats), iptr(order))
        assert rc == 0
        assert sorted(order.tolist()) == list(range(n))
        assert err.value < 1e-12, err.value
        assert stats[0] == (n + 31) // 32 and 2 <= stats[1] <= 64 and 0 < stats[2] <= 65535 and stats[3] > 0


@pytest.mark.parametrize("gen,dim,bs", [(lambda: HostMesh.cylinder3d(1, 3), 3, 3), (lambda: HostMesh.cylinder2d(2), 2, 2),
                                        (lambda: HostMesh.cylinder3d(1, 3), 3, 1), (lambda: HostMesh.cylinder3d(2, 8), 3, 3)])
def test_point_multicolour_sell_storage_cpu(gen, dim, bs):
    """ilu_ordering = 1 and the SpMV format without a GPU (nsb_debug_sell_check): colours are independent sets, and the
    SELL-32 copies of L, U (colours as row ranges) and of the whole matrix reproduce plain substitution and a CSR
    product through a host emulation of k_sell3 -- one and four lanes per row, small and large sort windows."""
    import ctypes as C

    from navierstokes_project_nm4pde_b200 import _lib
    from navierstokes_project_nm4pde_b200._lib import iptr

    n, rp, ci = _graph_of(gen, dim, bs)
    err, stats, order = C.c_double(0), np.zeros(3, np.int32), np.zeros(n, np.int32)
    for lanes, window in ((1, 4096), (4, 4096), (1, 64), (4, 7)):
        rc = _lib.lib().nsb_debug_sell_check(n, iptr(rp), iptr(ci), bs, lanes, window, C.byref(err), iptr(stats), iptr(order))
        assert rc == 0
        assert sorted(order.tolist()) == list(range(n))
        assert err.value < 1e-12, (lanes, window, err.value)
        assert 2 <= stats[0] <= 128 and stats[1] >= n * lanes // 32 and 1000 <= stats[2] < 4000


def test_driver_rendezvous_without_gpu(tmp_path):
    """The launcher contract of the C++ drivers (csrc/host/rendezvous.hpp, scripts/nsb_launch.sh): three
    processes find each other over TCP and all-gather blobs of 1 B .. 100 kB; no GPU is touched."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "navierstokes_project_nm4pde_b200", "csrc", "host", "bin", "navier_stokes3D")
    if not os.path.exists(exe):
        pytest.fail(f"{exe} is not built (make -C navierstokes_project_nm4pde_b200/csrc drivers)")
    r = subprocess.run([os.path.join(root, "scripts", "nsb_launch.sh"), "3", exe], capture_output=True, text=True,
                       env=dict(os.environ, NSB_RDV_SELFTEST="1"), cwd=tmp_path, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "rendezvous ok: rank 0 of 3" in r.stdout
    for k in (1, 2):
        assert f"rendezvous ok: rank {k} of 3" in (tmp_path / f"rank{k}.log").read_text()


@pytest.mark.parametrize("gen", [lambda: HostMesh.cylinder2d(1), lambda: HostMesh.cylinder3d(1, 3)])
def test_find_cell_matches_point_value(gen):
    """nsh_dofs_find_cell (the cell VectorTools::point_value evaluates in; lets one rank of a multi-rank run evaluate
    only points of cells it owns): barycentric coordinates reproduce the point, a linear field is interpolated exactly
    from them, points inside the obstacle / outside the channel are reported as -1."""
    m = gen()
    d = HostDofs(m)
    dim = m.dim
    L = _lib.lib()
    rng = np.random.default_rng(R_SEED)
    cc = d.cell_coords()
    cells = rng.integers(0, d.n_cells, 20)
    for c in cells:
        w = rng.uniform(0.05, 1.0, dim + 1)
        w /= w.sum()
        x = np.ascontiguousarray(w @ cc[c])
        lam = np.zeros(4)
        k = L.nsh_dofs_find_cell(d.h, _lib.dptr(x), _lib.dptr(lam))
        assert k >= 0
        assert np.allclose(lam[: dim + 1] @ cc[k], x, atol=1e-12) and lam[: dim + 1].min() >= -1e-10
        # a linear pressure field p = 1 + a.x through point_value
        a = np.arange(1, dim + 1, dtype=float)
        sol = np.zeros(d.N)
        sol[d.n_u:] = 1.0 + d.p_xyz @ a
        assert abs(d.point_value(sol, x)[dim] - (1.0 + x @ a)) < 1e-12
    inside_obstacle = np.array([0.2, 0.2, 0.2][:dim]) if dim == 2 else np.array([0.5, 0.2, 0.2])
    for x in (inside_obstacle, np.full(dim, -1.0)):
        lam = np.zeros(4)
        assert L.nsh_dofs_find_cell(d.h, _lib.dptr(np.ascontiguousarray(x)), _lib.dptr(lam)) == -1
        assert d.point_value(np.zeros(d.N), x) is None
