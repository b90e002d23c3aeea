"""Thin object wrappers over the C ABI: host meshes / DoFs (`nsh_*`) and the device engine (`nsb_*`).

These mirror, call for call, what the reference's `NavierStokes` methods would invoke through the
C ABI; they add no logic of their own beyond numpy <-> pointer plumbing.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import NsbError, NsbParams, dptr, iptr

BLK = {"F": 0, "Bt": 1, "B": 2, "Mp": 3, "S": 4}
MAT = {"system": 0, "mass": 1, "stiffness": 2, "convection": 3}
VARIANT = {"2d": 0, "3d": 1, "conv": 2}
PRECOND = {"yosida": 0, "simple": 1, "ayosida": 2, "asimple": 3}


class HostMesh:
    """Simplex mesh with boundary ids (replaces Triangulation + GridIn::read_msh)."""

    def __init__(self, handle):
        if not handle:
            raise NsbError(-6, "mesh could not be created / read")
        self.L = _lib.lib()
        self.h = C.c_void_p(handle)

    @classmethod
    def cylinder2d(cls, s=1):
        return cls(_lib.lib().nsh_mesh_cylinder2d(s))

    @classmethod
    def cylinder3d(cls, s=1, nz=6):
        return cls(_lib.lib().nsh_mesh_cylinder3d(s, nz))

    @classmethod
    def cube(cls, n):
        return cls(_lib.lib().nsh_mesh_cube(n))

    @classmethod
    def box(cls, dim, n, lo, hi):
        n = list(n) + [1] * (3 - len(n))
        lo = np.ascontiguousarray(list(lo) + [0.0] * (3 - len(lo)), dtype=np.float64)
        hi = np.ascontiguousarray(list(hi) + [0.0] * (3 - len(hi)), dtype=np.float64)
        return cls(_lib.lib().nsh_mesh_box(dim, n[0], n[1], n[2], dptr(lo), dptr(hi)))

    @classmethod
    def read_msh(cls, path):
        return cls(_lib.lib().nsh_mesh_read_msh(str(path).encode()))

    def write_msh(self, path):
        rc = self.L.nsh_mesh_write_msh(self.h, str(path).encode())
        if rc:
            raise NsbError(rc, f"cannot write {path}")

    def reorder_cells(self, mode, block=512):
        rc = self.L.nsh_mesh_reorder_cells(self.h, mode, block)
        if rc:
            raise NsbError(rc, "reorder_cells")

    @property
    def dim(self):
        return self.L.nsh_mesh_dim(self.h)

    @property
    def n_cells(self):
        return self.L.nsh_mesh_n_cells(self.h)

    @property
    def n_vertices(self):
        return self.L.nsh_mesh_n_vertices(self.h)

    @property
    def vertices(self):
        return np.ctypeslib.as_array(self.L.nsh_mesh_vertices(self.h), shape=(self.n_vertices, self.dim)).copy()

    @property
    def cells(self):
        return np.ctypeslib.as_array(self.L.nsh_mesh_cells(self.h), shape=(self.n_cells, self.dim + 1)).copy()

    @property
    def bfaces(self):
        n = self.L.nsh_mesh_n_bfaces(self.h)
        return np.ctypeslib.as_array(self.L.nsh_mesh_bfaces(self.h), shape=(n, self.dim)).copy()

    @property
    def bface_ids(self):
        n = self.L.nsh_mesh_n_bfaces(self.h)
        return np.ctypeslib.as_array(self.L.nsh_mesh_bface_ids(self.h), shape=(n,)).copy()

    @property
    def bface_cells(self):
        n = self.L.nsh_mesh_n_bfaces(self.h)
        return np.ctypeslib.as_array(self.L.nsh_mesh_bface_cells(self.h), shape=(n,)).copy()

    def partition(self, nparts):
        part = np.zeros(self.n_cells, np.int32)
        rc = self.L.nsh_partition_cells(self.h, nparts, iptr(part))
        if rc:
            raise NsbError(rc, "partition")
        return part

    def close(self):
        if self.h:
            self.L.nsh_mesh_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HostDofs:
    """P2-P1 DoF numbering in the reference's order (replaces DoFHandler + component_wise)."""

    def __init__(self, mesh: HostMesh):
        self.L = _lib.lib()
        self.mesh = mesh
        self.h = C.c_void_p(self.L.nsh_dofs_create(mesh.h))
        self.dim = mesh.dim
        self.n_nodes = self.L.nsh_dofs_n_nodes(self.h)
        self.n_p = self.L.nsh_dofs_n_p(self.h)
        self.n_u = self.dim * self.n_nodes
        self.N = self.n_u + self.n_p
        self.dpc = self.L.nsh_dofs_per_cell(self.h)
        self.n_cells = mesh.n_cells

    def cell_dofs(self, copy=True):
        a = np.ctypeslib.as_array(self.L.nsh_dofs_cell_dofs(self.h), shape=(self.n_cells, self.dpc))
        return a.copy() if copy else a

    def cell_coords(self, copy=True):
        a = np.ctypeslib.as_array(self.L.nsh_dofs_cell_coords(self.h), shape=(self.n_cells, self.dim + 1, self.dim))
        return a.copy() if copy else a

    @property
    def node_xyz(self):
        return np.ctypeslib.as_array(self.L.nsh_dofs_node_xyz(self.h), shape=(self.n_nodes, self.dim)).copy()

    @property
    def p_xyz(self):
        return np.ctypeslib.as_array(self.L.nsh_dofs_p_xyz(self.h), shape=(self.n_p, self.dim)).copy()

    def boundary_nodes(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        n = self.L.nsh_dofs_boundary_nodes(self.h, self.mesh.h, iptr(ids), len(ids), None)
        out = np.zeros(max(n, 1), np.int32)
        self.L.nsh_dofs_boundary_nodes(self.h, self.mesh.h, iptr(ids), len(ids), iptr(out))
        return out[:n]

    def point_value(self, solution, x):
        """VectorTools::point_value: (u_0..u_{dim-1}, p) at x, or None when no cell contains x."""
        sol = np.ascontiguousarray(solution, dtype=np.float64)
        xx = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros(self.dim + 1)
        rc = self.L.nsh_dofs_point_value(self.h, dptr(sol), dptr(xx), dptr(out))
        return out if rc == 0 else None

    def boundary_forces(self, solution, boundary_id, nu, rho=1.0):
        """(drag, lift) face integrals of NavierStokes::compute_forces over the faces with boundary_id."""
        sol = np.ascontiguousarray(solution, dtype=np.float64)
        out = np.zeros(2)
        rc = self.L.nsh_boundary_forces(self.mesh.h, self.h, dptr(sol), boundary_id, float(nu), float(rho), dptr(out))
        if rc:
            raise NsbError(rc, "boundary_forces")
        return out

    def write_vtu(self, solution, path):
        """DataOut::write_vtu stand-in: linear cells, vertex values of velocity and pressure."""
        sol = np.ascontiguousarray(solution, dtype=np.float64)
        rc = self.L.nsh_write_vtu(self.mesh.h, self.h, dptr(sol), str(path).encode())
        if rc:
            raise NsbError(rc, f"cannot write {path}")

    def boundary_faces(self, bid):
        n = self.L.nsh_dofs_boundary_faces(self.h, self.mesh.h, bid, None, None)
        fc, fl = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
        self.L.nsh_dofs_boundary_faces(self.h, self.mesh.h, bid, iptr(fc), iptr(fl))
        return fc[:n], fl[:n]

    def close(self):
        if self.h:
            self.L.nsh_dofs_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One device engine handle (`nsb_handle`): one per GPU / rank."""

    def __init__(self, dim, device=0, nranks=1, rank=0, unique_id=None):
        self.L = _lib.lib()
        self.dim = dim
        self.h = C.c_void_p()
        uid = C.c_char_p(unique_id) if unique_id is not None else None
        rc = self.L.nsb_create(C.byref(self.h), dim, device, nranks, rank, uid)
        if rc:
            msg = self.L.nsb_last_error(None).decode()
            self.h = None
            raise NsbError(rc, msg)
        self.params = NsbParams()
        self.L.nsb_default_params(C.byref(self.params), 0 if dim == 2 else 1)

    @staticmethod
    def device_count():
        return _lib.lib().nsb_device_count()

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        rc = _lib.lib().nsb_get_unique_id(buf)
        if rc:
            raise NsbError(rc, _lib.lib().nsb_last_error(None).decode())
        return buf.raw

    def _ck(self, rc):
        if rc:
            raise NsbError(rc, self.L.nsb_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.L.nsb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup
    def set_mesh(self, cell_coords, cell_dofs, n_u, n_p, n_u_owned=None, n_p_owned=None):
        cc = np.ascontiguousarray(cell_coords, dtype=np.float64)
        cd = np.ascontiguousarray(cell_dofs, dtype=np.int32)
        self.n_u, self.n_p = int(n_u), int(n_p)
        self.n_u_owned = self.n_u if n_u_owned is None else int(n_u_owned)
        self.n_p_owned = self.n_p if n_p_owned is None else int(n_p_owned)
        self.N = self.n_u + self.n_p
        self._ck(self.L.nsb_set_mesh(self.h, cd.shape[0], dptr(cc), iptr(cd), self.n_u, self.n_p, self.n_u_owned,
                                     self.n_p_owned))

    def set_quadrature(self, xi, w):
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        self._ck(self.L.nsb_set_quadrature(self.h, len(w), dptr(xi), dptr(w)))

    def default_params(self, variant):
        self.L.nsb_default_params(C.byref(self.params), VARIANT[variant] if isinstance(variant, str) else variant)
        return self.params

    def set_params(self, **kw):
        for k, v in kw.items():
            if k == "variant" and isinstance(v, str):
                v = VARIANT[v]
            if k == "precond_type" and isinstance(v, str):
                v = PRECOND[v]
            if not hasattr(self.params, k):
                raise AttributeError(k)
            setattr(self.params, k, v)
        self._ck(self.L.nsb_set_params(self.h, C.byref(self.params)))

    def finalize(self):
        self._ck(self.L.nsb_finalize_setup(self.h))

    def set_halo(self, nb_ranks, send_node_ptr, send_node_idx, recv_node_cnt, send_p_ptr, send_p_idx, recv_p_cnt):
        arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in
                (nb_ranks, send_node_ptr, send_node_idx, recv_node_cnt, send_p_ptr, send_p_idx, recv_p_cnt)]
        arrs = [a if a.size else np.zeros(1, np.int32) for a in arrs]
        self._ck(self.L.nsb_set_halo(self.h, len(nb_ranks), *[iptr(a) for a in arrs]))

    def p2p_export(self):
        """64-byte CUDA IPC handle of this rank's mailbox (after set_halo)."""
        buf = C.create_string_buffer(64)
        self._ck(self.L.nsb_p2p_export(self.h, buf))
        return buf.raw

    def p2p_attach(self, handles):
        """handles: the nranks 64-byte handles in rank order; switches the transport to peer memory."""
        blob = b"".join(handles)
        self._ck(self.L.nsb_p2p_attach(self.h, C.c_char_p(blob)))

    def pattern(self, blk):
        b = BLK[blk] if isinstance(blk, str) else blk
        nr, nnz = C.c_int32(0), C.c_int64(0)
        self._ck(self.L.nsb_get_pattern_size(self.h, b, C.byref(nr), C.byref(nnz)))
        rp = np.zeros(nr.value + 1, np.int32)
        ci = np.zeros(max(nnz.value, 1), np.int32)
        self._ck(self.L.nsb_get_pattern(self.h, b, iptr(rp), iptr(ci)))
        return rp, ci[: nnz.value]

    def check_pattern(self, blk, rowptr, colind):
        rp = np.ascontiguousarray(rowptr, dtype=np.int32)
        ci = np.ascontiguousarray(colind, dtype=np.int32)
        self._ck(self.L.nsb_check_pattern(self.h, BLK[blk], iptr(rp), iptr(ci)))

    # ---- boundary data / state
    def set_dirichlet(self, rows):
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        self._n_dir = len(rows)
        self._ck(self.L.nsb_set_dirichlet(self.h, len(rows), iptr(rows if rows.size else np.zeros(1, np.int32))))

    def set_dirichlet_values(self, vals):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        assert len(vals) == self._n_dir
        self._ck(self.L.nsb_set_dirichlet_values(self.h, dptr(vals if vals.size else np.zeros(1))))

    def set_neumann_rhs(self, rhs_u):
        if rhs_u is None:
            self._ck(self.L.nsb_set_neumann_rhs(self.h, None))
        else:
            r = np.ascontiguousarray(rhs_u, dtype=np.float64)
            self._ck(self.L.nsb_set_neumann_rhs(self.h, dptr(r)))

    def set_force_faces(self, face_cell, face_local, xi, w):
        """Obstacle faces of locally owned cells + the face rule for compute_forces (once, after set_mesh)."""
        fc = np.ascontiguousarray(face_cell, dtype=np.int32)
        fl = np.ascontiguousarray(face_local, dtype=np.int32)
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        z = np.zeros(1, np.int32)
        self._ck(self.L.nsb_set_force_faces(self.h, len(fc), iptr(fc if fc.size else z), iptr(fl if fl.size else z),
                                            len(w), dptr(xi), dptr(w)))

    def compute_forces(self, rho=1.0):
        """(drag, lift) integrals of NavierStokes::compute_forces of the current solution, summed over ranks."""
        out = np.zeros(2)
        self._ck(self.L.nsb_compute_forces(self.h, float(rho), dptr(out)))
        return out

    def set_solution(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.size == self.N
        self._ck(self.L.nsb_set_solution(self.h, dptr(x)))

    def get_solution(self):
        x = np.zeros(self.N)
        self._ck(self.L.nsb_get_solution(self.h, dptr(x)))
        return x

    def get_rhs(self):
        x = np.zeros(self.N)
        self._ck(self.L.nsb_get_rhs(self.h, dptr(x)))
        return x

    # ---- hot path
    def assemble_first(self, time=0.0):
        self._ck(self.L.nsb_assemble_first(self.h, time))

    def assemble_step(self, time=0.0):
        self._ck(self.L.nsb_assemble_step(self.h, time))

    def solve_step(self):
        its, tp, ts = C.c_int32(0), C.c_double(0), C.c_double(0)
        self._ck(self.L.nsb_solve_step(self.h, C.byref(its), C.byref(tp), C.byref(ts)))
        return its.value, tp.value, ts.value

    def step_host(self, first, time, dirichlet_values, solution_out):
        its = C.c_int32(0)
        dv = dptr(dirichlet_values) if dirichlet_values is not None and dirichlet_values.size else None
        so = dptr(solution_out) if solution_out is not None else None
        self._ck(self.L.nsb_step_host(self.h, 1 if first else 0, time, dv, so, C.byref(its)))
        return its.value

    # ---- parity harness
    def matrix_values(self, mat, blk):
        rp, ci = self.pattern(blk)
        v = np.zeros(max(len(ci), 1))
        self._ck(self.L.nsb_get_matrix_values(self.h, MAT[mat], BLK[blk], dptr(v)))
        return rp, ci, v[: len(ci)]

    def matrix(self, mat, blk):
        import scipy.sparse as sp

        rp, ci, v = self.matrix_values(mat, blk)
        ncols = {"F": self.n_u, "Bt": self.n_p, "B": self.n_u, "Mp": self.n_p, "S": self.n_p}[blk]
        return sp.csr_matrix((v, ci, rp), shape=(len(rp) - 1, ncols))

    def system_vmult(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.N)
        self._ck(self.L.nsb_op_system_vmult(self.h, dptr(x), dptr(y)))
        return y

    def block_vmult(self, blk, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        nout = self.n_u_owned if blk in ("F", "Bt") else self.n_p_owned
        y = np.zeros(nout)
        self._ck(self.L.nsb_op_block_vmult(self.h, BLK[blk], dptr(x), dptr(y)))
        return y

    def precond_init(self):
        self._ck(self.L.nsb_op_precond_init(self.h))

    def ilu_apply(self, which, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        self._ck(self.L.nsb_op_ilu_apply(self.h, which, dptr(x), dptr(y)))
        return y

    def precond_vmult(self, src, dst_in=None):
        src = np.ascontiguousarray(src, dtype=np.float64)
        dst = np.zeros(self.N)
        di = None if dst_in is None else dptr(np.ascontiguousarray(dst_in, dtype=np.float64))
        self._ck(self.L.nsb_op_precond_vmult(self.h, dptr(src), di, dptr(dst)))
        return dst

    def ilu_order(self, which):
        n = (self.n_u_owned // self.dim) if which == 0 else self.n_p_owned
        out = np.zeros(n, np.int32)
        self._ck(self.L.nsb_get_ilu_order(self.h, which, iptr(out)))
        return out

    def schur(self):
        import scipy.sparse as sp

        rp, ci = self.pattern("S")
        v = np.zeros(max(len(ci), 1))
        self._ck(self.L.nsb_get_schur_values(self.h, dptr(v)))
        return sp.csr_matrix((v[: len(ci)], ci, rp), shape=(self.n_p_owned, self.n_p))

    # ---- measurement
    def stat(self, name):
        return self.L.nsb_stat(self.h, name.encode())

    def bench_kernel(self, which, iters=10, flush_l2=True):
        ms, nbytes = C.c_double(0), C.c_double(0)
        self._ck(self.L.nsb_bench_kernel(self.h, which.encode(), iters, 1 if flush_l2 else 0, C.byref(ms), C.byref(nbytes)))
        return ms.value, nbytes.value

    def timer_start(self):
        self._ck(self.L.nsb_timer_mark(self.h, 0))

    def timer_stop_ms(self):
        self._ck(self.L.nsb_timer_mark(self.h, 1))
        ms = C.c_double(0)
        self._ck(self.L.nsb_timer_elapsed_ms(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self, reset=False):
        return int(self.L.nsb_launch_count(self.h, 1 if reset else 0))
