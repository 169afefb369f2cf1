"""ctypes binding of the C ABI in include/nbgpu.h (nbots_b200/lib/libnbgpu.so).

This is plumbing for the Python tests and bench.py; the product is the shared
library.  Loading fails loudly when the CUDA library has not been built -- there
is no CPU fallback (compute calls return NBGPU_ERR_CUDA without a device).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.environ.get("NBGPU_LIB_DIR") or os.path.join(_HERE, "lib")   # (diagnostic builds live in other dirs)
LIB_PATH = os.path.join(LIB_DIR, "libnbgpu.so")
SHIM_PATH = os.path.join(LIB_DIR, "libnbots_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "nbgpu.h")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)
vpp = C.POINTER(C.c_void_p)

OK, NOT_CONVERGED, DISTORTED_ELEMENT = 0, 1, 1
ERR_CUDA, ERR_ARG, ERR_PATTERN, ERR_NOMEM, ERR_COMM = 10, 11, 12, 13, 14
ASSEMBLY_GATHER, ASSEMBLY_ATOMIC, ASSEMBLY_COLOR = 0, 1, 2


class ElemTables(C.Structure):
    _fields_ = [("N_nodes", C.c_uint32), ("N_gp", C.c_uint32), ("gp_weight", C.c_double * 4),
                ("Ni", C.c_double * 16), ("dNi_dpsi", C.c_double * 16), ("dNi_deta", C.c_double * 16)]


class AssemblyParams(C.Structure):
    _fields_ = [("D", C.c_double * 4), ("density", C.c_double), ("D_void", C.c_double * 4),
                ("density_void", C.c_double), ("thickness", C.c_double), ("self_weight", C.c_int32),
                ("gravity", C.c_double * 2), ("mode", C.c_int32)]


class BCond(C.Structure):
    """One boundary condition, as one nb_bcond_push call (include/nbgpu.h nbgpu_bcond_t)."""
    _fields_ = [("kind", C.c_int32), ("where", C.c_int32), ("id", C.c_uint32), ("mask", C.c_int32 * 2),
                ("val", C.c_double * 2), ("fval", C.c_void_p)]


BCFUNC = C.CFUNCTYPE(None, f64p, C.c_double, f64p)


class MeshDesc(C.Structure):
    """nbgpu_mesh_desc_t: flat view of a mesh (include/nbgpu.h)."""
    _fields_ = [("N_nod", C.c_uint32), ("nod", f64p), ("N_elems", C.c_uint32), ("npe", C.c_uint32),
                ("adj", u32p), ("N_edg", C.c_uint32), ("edg", u32p), ("N_vtx", C.c_uint32),
                ("vtx", u32p), ("N_sgm", C.c_uint32), ("sgm_sizes", u32p), ("sgm_nodes", u32p)]

    @classmethod
    def of(cls, m):
        p = lambda a, t: a.ctypes.data_as(t)  # noqa: E731
        d = cls(m.n_nod, p(m.nod, f64p), m.n_elems, m.npe, p(m.adj, u32p), m.n_edg, p(m.edg, u32p),
                m.vtx.size, p(m.vtx, u32p), m.sgm_sizes.size, p(m.sgm_sizes, u32p), p(m.sgm_nodes, u32p))
        d._keep = m
        return d

# name -> (restype, argtypes); every symbol declared in include/nbgpu.h
SIGNATURES = {
    "nbgpu_init": (C.c_int, [C.c_int]),
    "nbgpu_finalize": (C.c_int, []),
    "nbgpu_device_count": (C.c_int, []),
    "nbgpu_sync": (C.c_int, []),
    "nbgpu_last_error": (C.c_char_p, []),
    "nbgpu_stream": (C.c_void_p, []),
    "nbgpu_launch_count": (C.c_uint64, []),
    "nbgpu_device_info": (C.c_int, [C.c_char_p, C.c_size_t]),
    "nbgpu_malloc": (C.c_int, [vpp, C.c_size_t]),
    "nbgpu_free": (C.c_int, [C.c_void_p]),
    "nbgpu_memset": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t]),
    "nbgpu_copy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "nbgpu_copy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "nbgpu_copy_d2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "nbgpu_host_alloc": (C.c_int, [vpp, C.c_size_t]),
    "nbgpu_host_free": (C.c_int, [C.c_void_p]),
    "nbgpu_timer_start": (C.c_int, []),
    "nbgpu_timer_stop": (C.c_int, [C.POINTER(C.c_float)]),
    "nbgpu_matrix_create_from_rows": (C.c_int, [C.c_uint32, u32p, C.c_void_p, C.c_void_p, vpp]),
    "nbgpu_matrix_create_from_csr": (C.c_int, [C.c_uint32, u32p, u32p, f64p, vpp]),
    "nbgpu_matrix_destroy": (C.c_int, [C.c_void_p]),
    "nbgpu_matrix_info": (C.c_int, [C.c_void_p, u32p, u64p, u32p, u64p]),
    "nbgpu_matrix_layout": (C.c_int, [C.c_void_p, u32p, u32p, u32p, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "nbgpu_matrix_set_values_rows": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nbgpu_matrix_set_values_csr": (C.c_int, [C.c_void_p, f64p]),
    "nbgpu_matrix_get_values_rows": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nbgpu_matrix_get_values_csr": (C.c_int, [C.c_void_p, f64p]),
    "nbgpu_matrix_get_pattern_csr": (C.c_int, [C.c_void_p, u32p, u32p]),
    "nbgpu_matrix_reset": (C.c_int, [C.c_void_p]),
    "nbgpu_spmv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nbgpu_spmv_host": (C.c_int, [C.c_void_p, f64p, f64p]),
    "nbgpu_pcg_jacobi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, u32p, f64p]),
    "nbgpu_cg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, u32p, f64p]),
    "nbgpu_pcg_jacobi_host": (C.c_int, [C.c_void_p, f64p, f64p, C.c_uint32, C.c_double, u32p, f64p]),
    "nbgpu_cg_host": (C.c_int, [C.c_void_p, f64p, f64p, C.c_uint32, C.c_double, u32p, f64p]),
    "nbgpu_set_reduction_order": (C.c_int, [C.c_int]),
    "nbgpu_krylov_profile": (C.c_int, [C.c_int]),
    "nbgpu_krylov_profile_get": (C.c_int, [f64p, u32p]),
    "nbgpu_pattern_from_mesh": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, u32p, C.c_uint32, u32p, C.c_uint32,
                                          u32p, u32p, u64p]),
    "nbgpu_mesh_create": (C.c_int, [C.c_uint32, f64p, C.c_uint32, C.c_uint32, u32p, vpp]),
    "nbgpu_mesh_destroy": (C.c_int, [C.c_void_p]),
    "nbgpu_matrix_create_from_mesh": (C.c_int, [C.c_void_p, C.c_uint32, u32p, vpp]),
    "nbgpu_mesh_coloring": (C.c_int, [C.c_void_p, u32p, u8p]),
    "nbgpu_elem_tables_default": (C.c_int, [C.c_uint32, C.POINTER(ElemTables)]),
    "nbgpu_constitutive_matrix": (C.c_int, [C.c_double, C.c_double, C.c_int, f64p]),
    "nbgpu_assemble_elasticity2d": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(ElemTables),
                                              C.POINTER(AssemblyParams), u8p, f64p, C.c_void_p, u32p]),
    "nbgpu_assemble_elasticity2d_damage": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(ElemTables),
                                                     C.POINTER(AssemblyParams), u8p, f64p, C.c_void_p, u32p]),
    "nbgpu_assemble_lumped_mass": (C.c_int, [C.c_void_p, C.POINTER(ElemTables), C.c_double, C.c_double, C.c_double,
                                             u8p, C.c_void_p, u32p]),
    "nbgpu_vector_add_entries": (C.c_int, [C.c_void_p, C.c_uint32, u32p, f64p]),
    "nbgpu_apply_dirichlet": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, u32p, f64p]),
    "nbgpu_dirichlet_create": (C.c_int, [C.c_uint32, C.c_uint32, u32p, f64p, vpp]),
    "nbgpu_dirichlet_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nbgpu_dirichlet_destroy": (C.c_int, [C.c_void_p]),
    "nbgpu_vector_add_entries_dev": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "nbgpu_compute_strain": (C.c_int, [C.c_void_p, C.POINTER(ElemTables), C.c_void_p, C.c_void_p]),
    "nbgpu_stress_from_strain": (C.c_int, [C.c_uint32, C.c_uint32, f64p, f64p, u8p, C.c_void_p, C.c_void_p]),
    "nbgpu_gp_to_nodes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]),
    "nbgpu_von_mises": (C.c_int, [C.c_uint64, C.c_void_p, C.c_void_p]),
    "nbgpu_main_stress": (C.c_int, [C.c_uint64, C.c_void_p, C.c_void_p]),
    "nbgpu_fem_session_create": (C.c_int, [C.c_void_p, C.c_void_p, f64p, C.c_double, C.c_uint32, u32p, f64p, C.c_uint32,
                                           u32p, f64p, C.c_int, f64p, C.c_double, C.c_int, vpp]),
    "nbgpu_fem_session_step": (C.c_int, [C.c_void_p, u8p, f64p, C.c_int, C.c_uint32, C.c_double, C.c_void_p]),
    "nbgpu_fem_session_results": (C.c_int, [C.c_void_p, f64p, f64p]),
    "nbgpu_fem_session_destroy": (C.c_int, [C.c_void_p]),
    "nbgpu_matrix_create_local": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, u32p, u32p, f64p, vpp]),
    "nbgpu_dist_ext_layout": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, u32p, u32p, u32p]),
    "nbgpu_dist_plan_layout": (C.c_int, [C.c_void_p, u32p, u32p, u32p, u32p]),
    "nbgpu_dist_plan_visit_order": (C.c_int, [C.c_void_p, C.c_uint32, u32p, u32p, u32p]),
    "nbgpu_dist_connect_local": (C.c_int, [C.c_void_p, vpp, C.POINTER(C.c_int)]),
    "nbgpu_set_pcg_mode": (C.c_int, [C.c_int]),
    "nbgpu_thread_bind_device": (C.c_int, [C.c_int]),
    "nbgpu_dist_plan_create": (C.c_int, [C.c_int, C.c_int, u32p, u32p, u32p, vpp]),
    "nbgpu_dist_plan_destroy": (C.c_int, [C.c_void_p]),
    "nbgpu_dist_plan_info": (C.c_int, [C.c_void_p, u32p, u32p, u64p, u32p]),
    "nbgpu_dist_plan_halo_ids": (C.c_int, [C.c_void_p, u32p]),
    "nbgpu_dist_plan_local_cols": (C.c_int, [C.c_void_p, u32p]),
    "nbgpu_dist_plan_set_sends": (C.c_int, [C.c_void_p, u32p, u32p, u32p]),
    "nbgpu_dist_create": (C.c_int, [C.c_int, C.c_int, C.c_size_t, C.c_void_p, vpp]),
    "nbgpu_dist_connect": (C.c_int, [C.c_void_p, C.c_void_p, u64p]),
    "nbgpu_dist_destroy": (C.c_int, [C.c_void_p]),
    "nbgpu_dist_error": (C.c_int, [C.c_void_p]),
    "nbgpu_dist_pcg_jacobi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                        C.c_double, u32p, f64p]),
    "nbgpu_dist_cg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_double,
                                u32p, f64p]),
    "nbgpu_dist_input_vector": (C.c_void_p, [C.c_void_p, C.c_void_p]),
    "nbgpu_dist_spmv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nbgpu_devices_from_env": (C.c_int, []),
    "nbgpu_fem_static_elasticity2d_lists_multi": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, f64p, C.c_double, C.c_uint32,
                                                            u32p, f64p, C.c_uint32, u32p, f64p, C.c_int, f64p,
                                                            C.c_double, u8p, C.c_double, f64p, f64p, C.c_void_p]),
    "nbgpu_solve_rows_multi": (C.c_int, [C.c_int, C.c_int, C.c_uint32, u32p, C.c_void_p, C.c_void_p, f64p, f64p,
                                         C.c_uint32, C.c_double, u32p, f64p]),
    "nbgpu_partition_nodes": (C.c_int, [C.c_uint32, C.c_int, C.c_uint32, u32p]),
    "nbgpu_dist_fem_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, u32p, C.c_void_p, f64p, C.c_double, C.c_uint32,
                                        u32p, f64p, C.c_uint32, u32p, f64p, C.c_int, f64p, C.c_double, C.c_void_p,
                                        vpp]),
    "nbgpu_dist_fem_destroy": (C.c_int, [C.c_void_p]),
    "nbgpu_dist_plan_from_mesh": (C.c_int, [C.c_void_p, C.c_int, C.c_int, u32p, u32p, vpp]),
    "nbgpu_dist_plan_sends": (C.c_int, [C.c_void_p, u32p, u32p, u32p]),
    "nbgpu_dist_fem_info": (C.c_int, [C.c_void_p, u32p, u64p, u32p, u32p, u64p, f64p]),
    "nbgpu_dist_fem_connect": (C.c_int, [C.c_void_p, C.c_void_p, u64p]),
    "nbgpu_dist_fem_connect_local": (C.c_int, [C.c_void_p, vpp, C.POINTER(C.c_int)]),
    "nbgpu_dist_fem_assemble": (C.c_int, [C.c_void_p, u8p, f64p, u32p]),
    "nbgpu_dist_fem_solve": (C.c_int, [C.c_void_p, C.c_int, C.c_uint32, C.c_double, u32p, f64p]),
    "nbgpu_dist_fem_results": (C.c_int, [C.c_void_p, f64p]),
    "nbgpu_dist_fem_matrix": (C.c_void_p, [C.c_void_p]),
    "nbgpu_dist_fem_plan": (C.c_void_p, [C.c_void_p]),
    "nbgpu_dist_fem_dist": (C.c_void_p, [C.c_void_p]),
    "nbgpu_dist_fem_rhs": (C.c_void_p, [C.c_void_p]),
    "nbgpu_dist_fem_solution": (C.c_void_p, [C.c_void_p]),
}

_lib = None


class NbgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nbgpu error {code}: {msg}")
        self.code = code


def lib():
    """The loaded libnbgpu.so.  Raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                               "(make -C nbots_b200/csrc); nbots_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            if not hasattr(L, name):
                continue   # optional groups (driver / comm) are bound by their own modules
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(code, ok=(OK,)):
    if code not in ok:
        raise NbgpuError(code, lib().nbgpu_last_error().decode())
    return code
