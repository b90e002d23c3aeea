"""ctypes binding of libnsb.so (C ABI declared in include/nsb.h).

The library is built in-tree by `__graft_entry__.build()` (or `make -C csrc`).  There is no
fallback: if the shared object is missing or no CUDA device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnsb.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int32)


class NsbParams(C.Structure):
    _fields_ = [
        ("nu", C.c_double), ("deltat", C.c_double), ("variant", C.c_int32), ("precond_type", C.c_int32),
        ("gmres_tmp", C.c_int32), ("outer_maxit", C.c_int32), ("outer_tol", C.c_double),
        ("inner_maxit", C.c_int32), ("inner_rtol", C.c_double), ("alpha_simple", C.c_double),
        ("alpha_asimple", C.c_double), ("dirichlet_mode", C.c_int32), ("assembly_kernel", C.c_int32),
        ("sptrsv_kernel", C.c_int32), ("ilu_ordering", C.c_int32), ("orthogonalisation", C.c_int32), ("ilu_ordering_schur", C.c_int32), ("reserved", C.c_int32 * 5),
    ]


# every symbol declared in include/nsb.h: name -> (restype, argtypes)
_H = C.c_void_p
SIGNATURES = {
    "nsb_default_params": (C.c_int, [C.POINTER(NsbParams), C.c_int]),
    "nsb_create": (C.c_int, [C.POINTER(_H), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "nsb_destroy": (C.c_int, [_H]),
    "nsb_last_error": (C.c_char_p, [_H]),
    "nsb_get_unique_id": (C.c_int, [C.c_void_p]),
    "nsb_device_count": (C.c_int, []),
    "nsb_set_mesh": (C.c_int, [_H, C.c_int32, c_double_p, c_int_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "nsb_set_quadrature": (C.c_int, [_H, C.c_int32, c_double_p, c_double_p]),
    "nsb_set_params": (C.c_int, [_H, C.POINTER(NsbParams)]),
    "nsb_finalize_setup": (C.c_int, [_H]),
    "nsb_check_pattern": (C.c_int, [_H, C.c_int, c_int_p, c_int_p]),
    "nsb_get_pattern_size": (C.c_int, [_H, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "nsb_get_pattern": (C.c_int, [_H, C.c_int, c_int_p, c_int_p]),
    "nsb_set_halo": (C.c_int, [_H, C.c_int32, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "nsb_p2p_export": (C.c_int, [_H, C.c_void_p]),
    "nsb_p2p_attach": (C.c_int, [_H, C.c_void_p]),
    "nsb_set_dirichlet": (C.c_int, [_H, C.c_int32, c_int_p]),
    "nsb_set_dirichlet_values": (C.c_int, [_H, c_double_p]),
    "nsb_set_neumann_rhs": (C.c_int, [_H, c_double_p]),
    "nsb_set_force_faces": (C.c_int, [_H, C.c_int32, c_int_p, c_int_p, C.c_int32, c_double_p, c_double_p]),
    "nsb_compute_forces": (C.c_int, [_H, C.c_double, c_double_p]),
    "nsb_set_solution": (C.c_int, [_H, c_double_p]),
    "nsb_get_solution": (C.c_int, [_H, c_double_p]),
    "nsb_allreduce_sum": (C.c_int, [_H, c_double_p, C.c_int32]),
    "nsb_assemble_first": (C.c_int, [_H, C.c_double]),
    "nsb_assemble_step": (C.c_int, [_H, C.c_double]),
    "nsb_solve_step": (C.c_int, [_H, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "nsb_step_host": (C.c_int, [_H, C.c_int, C.c_double, c_double_p, c_double_p, C.POINTER(C.c_int32)]),
    "nsb_get_matrix_values": (C.c_int, [_H, C.c_int, C.c_int, c_double_p]),
    "nsb_get_rhs": (C.c_int, [_H, c_double_p]),
    "nsb_op_system_vmult": (C.c_int, [_H, c_double_p, c_double_p]),
    "nsb_op_block_vmult": (C.c_int, [_H, C.c_int, c_double_p, c_double_p]),
    "nsb_op_precond_init": (C.c_int, [_H]),
    "nsb_op_ilu_apply": (C.c_int, [_H, C.c_int, c_double_p, c_double_p]),
    "nsb_op_precond_vmult": (C.c_int, [_H, c_double_p, c_double_p, c_double_p]),
    "nsb_get_schur_values": (C.c_int, [_H, c_double_p]),
    "nsb_get_ilu_order": (C.c_int, [_H, C.c_int, c_int_p]),
    "nsb_stat": (C.c_double, [_H, C.c_char_p]),
    "nsb_bench_kernel": (C.c_int, [_H, C.c_char_p, C.c_int, C.c_int, c_double_p, c_double_p]),
    "nsb_launch_count": (C.c_int64, [_H, C.c_int]),
    "nsb_timer_mark": (C.c_int, [_H, C.c_int]),
    "nsb_timer_elapsed_ms": (C.c_int, [_H, c_double_p]),
    "nsb_debug_sd_check": (C.c_int, [C.c_int32, c_int_p, c_int_p, c_double_p, C.c_int32, c_int_p, C.c_int32, C.c_int32,
                                     c_double_p, c_int_p, c_int_p]),
    "nsb_debug_bsell_check": (C.c_int, [C.c_int32, c_int_p, c_int_p, C.c_int32, C.c_int32, c_double_p, c_int_p, c_int_p]),
    "nsb_debug_sell_check": (C.c_int, [C.c_int32, c_int_p, c_int_p, C.c_int32, C.c_int32, C.c_int32, c_double_p, c_int_p, c_int_p]),
    "nsb_debug_setup_fingerprint": (C.c_int, [C.c_int32, C.c_int32, c_double_p, c_int_p, C.c_int32, C.c_int32, C.c_int32,
                                              C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.c_int32, c_int_p]),
    # host prerequisites
    "nsh_mesh_cylinder2d": (_H, [C.c_int]),
    "nsh_mesh_cylinder3d": (_H, [C.c_int, C.c_int]),
    "nsh_mesh_cube": (_H, [C.c_int]),
    "nsh_mesh_box": (_H, [C.c_int, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p]),
    "nsh_mesh_read_msh": (_H, [C.c_char_p]),
    "nsh_mesh_write_msh": (C.c_int, [_H, C.c_char_p]),
    "nsh_mesh_free": (None, [_H]),
    "nsh_mesh_dim": (C.c_int, [_H]),
    "nsh_mesh_n_vertices": (C.c_int32, [_H]),
    "nsh_mesh_n_cells": (C.c_int32, [_H]),
    "nsh_mesh_n_bfaces": (C.c_int32, [_H]),
    "nsh_mesh_vertices": (c_double_p, [_H]),
    "nsh_mesh_cells": (c_int_p, [_H]),
    "nsh_mesh_bfaces": (c_int_p, [_H]),
    "nsh_mesh_bface_ids": (c_int_p, [_H]),
    "nsh_mesh_bface_cells": (c_int_p, [_H]),
    "nsh_mesh_reorder_cells": (C.c_int, [_H, C.c_int, C.c_int]),
    "nsh_dofs_create": (_H, [_H]),
    "nsh_dofs_free": (None, [_H]),
    "nsh_dofs_n_nodes": (C.c_int32, [_H]),
    "nsh_dofs_n_p": (C.c_int32, [_H]),
    "nsh_dofs_per_cell": (C.c_int32, [_H]),
    "nsh_dofs_cell_dofs": (c_int_p, [_H]),
    "nsh_dofs_node_xyz": (c_double_p, [_H]),
    "nsh_dofs_p_xyz": (c_double_p, [_H]),
    "nsh_dofs_cell_coords": (c_double_p, [_H]),
    "nsh_dofs_boundary_nodes": (C.c_int32, [_H, _H, c_int_p, C.c_int32, c_int_p]),
    "nsh_dofs_boundary_faces": (C.c_int32, [_H, _H, C.c_int32, c_int_p, c_int_p]),
    "nsh_dofs_point_value": (C.c_int, [_H, c_double_p, c_double_p, c_double_p]),
    "nsh_boundary_forces": (C.c_int, [_H, _H, c_double_p, C.c_int32, C.c_double, C.c_double, c_double_p]),
    "nsh_dofs_find_cell": (C.c_int32, [_H, c_double_p, c_double_p]),
    "nsh_write_vtu": (C.c_int, [_H, _H, c_double_p, C.c_char_p]),
    "nsh_partition_cells": (C.c_int, [_H, C.c_int, c_int_p]),
    "nsh_local_create": (_H, [_H, _H, C.c_int32, C.c_int32]),
    "nsh_local_free": (None, [_H]),
    "nsh_local_n_cells": (C.c_int32, [_H]),
    "nsh_local_n_nodes": (C.c_int32, [_H]),
    "nsh_local_n_p": (C.c_int32, [_H]),
    "nsh_local_n_nodes_owned": (C.c_int32, [_H]),
    "nsh_local_n_p_owned": (C.c_int32, [_H]),
    "nsh_local_cells": (c_int_p, [_H]),
    "nsh_local_cell_part": (c_int_p, [_H]),
    "nsh_local_cell_dofs": (c_int_p, [_H]),
    "nsh_local_cell_coords": (c_double_p, [_H]),
    "nsh_local_node_gid": (c_int_p, [_H]),
    "nsh_local_p_gid": (c_int_p, [_H]),
    "nsh_local_g2l_node": (c_int_p, [_H]),
    "nsh_local_g2l_cell": (c_int_p, [_H]),
    "nsh_local_halo": (C.c_int32, [_H] + [C.POINTER(c_int_p)] * 7),
}

_lib = None


class NsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nsb error {code}: {msg}")
        self.code = code


def lib():
    """Load libnsb.so; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(the engine has no CPU fallback)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)
