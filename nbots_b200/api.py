"""Thin object wrappers over the C ABI for the Python tests and bench.py.

Every method is one C call (see include/nbgpu.h); nothing is computed in Python.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from . import capi
from .capi import check, lib, f64p, u32p, u64p, u8p


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


class DeviceBuffer:
    """A device allocation made by the library (nbgpu_malloc)."""

    def __init__(self, nbytes: int):
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        check(lib().nbgpu_malloc(C.byref(p), self.nbytes))
        self.ptr = p.value

    @classmethod
    def from_host(cls, a: np.ndarray):
        a = np.ascontiguousarray(a)
        buf = cls(a.nbytes)
        check(lib().nbgpu_copy_h2d(buf.ptr, a.ctypes.data, a.nbytes))
        return buf

    @classmethod
    def zeros(cls, n: int, dtype=np.float64):
        buf = cls(n * np.dtype(dtype).itemsize)
        check(lib().nbgpu_memset(buf.ptr, 0, buf.nbytes))
        return buf

    def upload(self, a: np.ndarray):
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        check(lib().nbgpu_copy_h2d(self.ptr, a.ctypes.data, a.nbytes))

    def to_host(self, dtype=np.float64, count=None) -> np.ndarray:
        n = self.nbytes // np.dtype(dtype).itemsize if count is None else count
        out = np.empty(n, dtype=dtype)
        check(lib().nbgpu_copy_d2h(out.ctypes.data, self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            lib().nbgpu_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """numpy view over pinned host memory (nbgpu_host_alloc)."""

    def __init__(self, n: int, dtype=np.float64):
        self.nbytes = n * np.dtype(dtype).itemsize
        p = C.c_void_p()
        check(lib().nbgpu_host_alloc(C.byref(p), self.nbytes))
        self.ptr = p.value
        buf = (C.c_char * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=n)

    def free(self):
        if self.ptr:
            self.array = None
            lib().nbgpu_host_free(self.ptr)
            self.ptr = None


class Matrix:
    """Device-resident nb_sparse_t (SELL-32)."""

    def __init__(self, handle):
        self.h = handle
        N = C.c_uint32(); nnz = C.c_uint64(); ns = C.c_uint32(); st = C.c_uint64()
        check(lib().nbgpu_matrix_info(self.h, C.byref(N), C.byref(nnz), C.byref(ns), C.byref(st)))
        self.N, self.nnz, self.n_slices, self.stored = N.value, nnz.value, ns.value, st.value
        sg = C.c_uint32(); uw = C.c_uint32(); mw = C.c_uint32(); bl = C.c_int(); i16 = C.c_int()
        check(lib().nbgpu_matrix_layout(self.h, C.byref(sg), C.byref(uw), C.byref(mw), C.byref(bl), C.byref(i16)))
        self.sigma, self.uniform_width, self.max_width, self.blocked = sg.value, uw.value, mw.value, bool(bl.value)
        self.idx16 = bool(i16.value)

    @classmethod
    def from_csr(cls, rows_size, cols, vals=None):
        rows_size, cols = _u32(rows_size), _u32(cols)
        vals = None if vals is None else _f64(vals)
        h = C.c_void_p()
        check(lib().nbgpu_matrix_create_from_csr(rows_size.size, _ptr(rows_size, u32p), _ptr(cols, u32p),
                                                 _ptr(vals, f64p), C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_row_pointers(cls, N, rows_size_ptr, rows_index_ptr, rows_values_ptr):
        """From the raw jagged arrays of a genuine nb_sparse_t (addresses)."""
        h = C.c_void_p()
        check(lib().nbgpu_matrix_create_from_rows(N, C.cast(rows_size_ptr, u32p), rows_index_ptr,
                                                  rows_values_ptr, C.byref(h)))
        return cls(h.value)

    def set_values_csr(self, vals):
        vals = _f64(vals)
        assert vals.size == self.nnz
        check(lib().nbgpu_matrix_set_values_csr(self.h, _ptr(vals, f64p)))

    def values_csr(self) -> np.ndarray:
        out = np.empty(self.nnz, dtype=np.float64)
        check(lib().nbgpu_matrix_get_values_csr(self.h, _ptr(out, f64p)))
        return out

    def pattern_csr(self):
        rs = np.empty(self.N, dtype=np.uint32)
        cols = np.empty(self.nnz, dtype=np.uint32)
        check(lib().nbgpu_matrix_get_pattern_csr(self.h, _ptr(rs, u32p), _ptr(cols, u32p)))
        return rs, cols

    def reset(self):
        check(lib().nbgpu_matrix_reset(self.h))

    # -- SpMV ------------------------------------------------------------
    def spmv(self, d_in: DeviceBuffer, d_out: DeviceBuffer):
        check(lib().nbgpu_spmv(self.h, d_in.ptr, d_out.ptr))

    def spmv_host(self, x) -> np.ndarray:
        x = _f64(x)
        y = np.empty(self.N)
        check(lib().nbgpu_spmv_host(self.h, _ptr(x, f64p), _ptr(y, f64p)))
        return y

    # -- Krylov ----------------------------------------------------------
    def _solve_dev(self, fn, d_b, d_x, max_iter, tol):
        it = C.c_uint32(0); res = C.c_double(0)
        st = fn(self.h, d_b.ptr, d_x.ptr, self.N if max_iter is None else max_iter, tol, C.byref(it), C.byref(res))
        check(st, ok=(capi.OK, capi.NOT_CONVERGED))
        return st, it.value, res.value

    def pcg_jacobi(self, d_b, d_x, max_iter=None, tol=1e-8):
        return self._solve_dev(lib().nbgpu_pcg_jacobi, d_b, d_x, max_iter, tol)

    def cg(self, d_b, d_x, max_iter=None, tol=1e-8):
        return self._solve_dev(lib().nbgpu_cg, d_b, d_x, max_iter, tol)

    def _solve_host(self, fn, b, x0, max_iter, tol):
        b = _f64(b)
        x = np.zeros(self.N) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        it = C.c_uint32(0); res = C.c_double(0)
        st = fn(self.h, _ptr(b, f64p), _ptr(x, f64p), self.N if max_iter is None else max_iter, tol,
                C.byref(it), C.byref(res))
        check(st, ok=(capi.OK, capi.NOT_CONVERGED))
        return st, x, it.value, res.value

    def pcg_jacobi_host(self, b, x0=None, max_iter=None, tol=1e-8):
        return self._solve_host(lib().nbgpu_pcg_jacobi_host, b, x0, max_iter, tol)

    def cg_host(self, b, x0=None, max_iter=None, tol=1e-8):
        return self._solve_host(lib().nbgpu_cg_host, b, x0, max_iter, tol)

    # -- boundary conditions ----------------------------------------------
    def apply_dirichlet(self, d_F: DeviceBuffer, dofs, values):
        dofs, values = _u32(dofs), _f64(values)
        check(lib().nbgpu_apply_dirichlet(self.h, d_F.ptr, dofs.size, _ptr(dofs, u32p), _ptr(values, f64p)))

    def destroy(self):
        if self.h:
            lib().nbgpu_matrix_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def pattern_from_mesh(m, vars_per_node=2, use_edges=True):
    """Host pattern builder (nbgpu_pattern_from_mesh) -> (rows_size, cols)."""
    L = lib()
    rs = np.empty(m.n_nod * vars_per_node, dtype=np.uint32)
    nnz = C.c_uint64(0)
    edg = m.edg if use_edges else None
    n_edg = m.n_edg if use_edges else 0
    check(L.nbgpu_pattern_from_mesh(m.n_nod, m.n_elems, m.npe, _ptr(m.adj, u32p), n_edg, _ptr(edg, u32p),
                                    vars_per_node, _ptr(rs, u32p), None, C.byref(nnz)))
    cols = np.empty(nnz.value, dtype=np.uint32)
    check(L.nbgpu_pattern_from_mesh(m.n_nod, m.n_elems, m.npe, _ptr(m.adj, u32p), n_edg, _ptr(edg, u32p),
                                    vars_per_node, _ptr(rs, u32p), _ptr(cols, u32p), C.byref(nnz)))
    return rs, cols


def elem_tables(npe: int) -> capi.ElemTables:
    t = capi.ElemTables()
    check(lib().nbgpu_elem_tables_default(npe, C.byref(t)))
    return t


def constitutive_matrix(E, nu, analysis=0) -> np.ndarray:
    D = np.zeros(4)
    check(lib().nbgpu_constitutive_matrix(E, nu, analysis, _ptr(D, f64p)))
    return D


class Mesh:
    """Device-resident FEM mesh (nbgpu_mesh_create)."""

    def __init__(self, m):
        self.m = m
        self.tables = elem_tables(m.npe)
        h = C.c_void_p()
        check(lib().nbgpu_mesh_create(m.n_nod, _ptr(m.nod, f64p), m.n_elems, m.npe, _ptr(m.adj, u32p), C.byref(h)))
        self.h = h.value

    def create_matrix(self, with_edges=True):
        """nbgpu_matrix_create_from_mesh: the pattern built on the device -> Matrix, or None when the mesh does
        not qualify for the device path (then: pattern_from_mesh + Matrix.from_csr)."""
        h = C.c_void_p()
        m = self.m
        edg = m.edg if with_edges else None
        check(lib().nbgpu_matrix_create_from_mesh(self.h, m.n_edg if with_edges else 0, _ptr(edg, u32p), C.byref(h)))
        return Matrix(h.value) if h.value else None

    def coloring(self):
        """(n_colors, colour of every element) of the device-built element colouring."""
        n = C.c_uint32(0)
        colors = np.zeros(max(1, self.m.n_elems), dtype=np.uint8)
        check(lib().nbgpu_mesh_coloring(self.h, C.byref(n), _ptr(colors, u8p)))
        return n.value, colors[:self.m.n_elems]

    def assemble(self, K: Matrix, d_F: DeviceBuffer, E, nu, density=0.0, self_weight=False, gravity=(0.0, 0.0),
                 analysis=0, thickness=1.0, enabled=None, elem_scale=None, mode=capi.ASSEMBLY_GATHER, gp_damage=None):
        """pipeline_assemble_system; with gp_damage the damage driver's loop (static_damage2D.c:474-569)."""
        p = capi.AssemblyParams()
        D = constitutive_matrix(E, nu, analysis)
        for k in range(4):
            p.D[k] = D[k]
            p.D_void[k] = 1e-6
        p.density, p.density_void, p.thickness = density, 1e-6, thickness
        p.self_weight = int(self_weight)
        p.gravity[0], p.gravity[1] = gravity
        p.mode = mode
        en = None if enabled is None else np.ascontiguousarray(enabled, dtype=np.uint8)
        sc = None if elem_scale is None else _f64(elem_scale)
        bad = C.c_uint32(0)
        if gp_damage is not None:
            assert elem_scale is None
            st = lib().nbgpu_assemble_elasticity2d_damage(K.h, self.h, C.byref(self.tables), C.byref(p), _ptr(en, u8p),
                                                          _ptr(_f64(gp_damage), f64p), d_F.ptr, C.byref(bad))
            check(st, ok=(capi.OK, capi.DISTORTED_ELEMENT))
            return st, bad.value
        st = lib().nbgpu_assemble_elasticity2d(K.h, self.h, C.byref(self.tables), C.byref(p), _ptr(en, u8p),
                                               _ptr(sc, f64p), d_F.ptr, C.byref(bad))
        check(st, ok=(capi.OK, capi.DISTORTED_ELEMENT))
        return st, bad.value

    def lumped_mass(self, d_M: DeviceBuffer, density, thickness=1.0, enabled=None):
        """The M vector of pipeline_assemble_system (pipeline.c:216-222, :256-259); returns (status, first bad element)."""
        en = None if enabled is None else np.ascontiguousarray(enabled, dtype=np.uint8)
        bad = C.c_uint32(0)
        st = lib().nbgpu_assemble_lumped_mass(self.h, C.byref(self.tables), density, 1e-6, thickness, _ptr(en, u8p),
                                              d_M.ptr, C.byref(bad))
        check(st, ok=(capi.OK, capi.DISTORTED_ELEMENT))
        return st, bad.value

    def compute_strain(self, d_disp: DeviceBuffer, d_strain: DeviceBuffer):
        check(lib().nbgpu_compute_strain(self.h, C.byref(self.tables), d_disp.ptr, d_strain.ptr))

    def gp_to_nodes(self, n_comp: int, d_gp_values: DeviceBuffer, d_nodal: DeviceBuffer) -> int:
        """nb_fem_interpolate_from_gpoints_to_nodes; returns 0, or 1 for a distorted element."""
        return check(lib().nbgpu_gp_to_nodes(self.h, C.byref(self.tables), n_comp, d_gp_values.ptr, d_nodal.ptr),
                     ok=(capi.OK, capi.DISTORTED_ELEMENT))

    def destroy(self):
        if self.h:
            lib().nbgpu_mesh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def vector_add_entries(d_F: DeviceBuffer, dofs, adds):
    dofs, adds = _u32(dofs), _f64(adds)
    check(lib().nbgpu_vector_add_entries(d_F.ptr, dofs.size, _ptr(dofs, u32p), _ptr(adds, f64p)))


def stress_from_strain(n_elems, n_gp, D, d_strain: DeviceBuffer, d_stress: DeviceBuffer, enabled=None):
    D = _f64(D)
    Dv = np.full(4, 1e-6)
    en = None if enabled is None else np.ascontiguousarray(enabled, dtype=np.uint8)
    check(lib().nbgpu_stress_from_strain(n_elems, n_gp, _ptr(D, f64p), _ptr(Dv, f64p), _ptr(en, u8p), d_strain.ptr,
                                         d_stress.ptr))


def von_mises(n_points, d_stress: DeviceBuffer, d_vm: DeviceBuffer):
    check(lib().nbgpu_von_mises(n_points, d_stress.ptr, d_vm.ptr))


def main_stress(n_points, d_stress: DeviceBuffer, d_main: DeviceBuffer):
    check(lib().nbgpu_main_stress(n_points, d_stress.ptr, d_main.ptr))


def timer_start():
    check(lib().nbgpu_timer_start())


def timer_stop() -> float:
    ms = C.c_float(0)
    check(lib().nbgpu_timer_stop(C.byref(ms)))
    return ms.value


def sync():
    check(lib().nbgpu_sync())


def launch_count() -> int:
    return int(lib().nbgpu_launch_count())


def device_info() -> dict:
    import json
    buf = C.create_string_buffer(1024)
    check(lib().nbgpu_device_info(buf, 1024))
    return json.loads(buf.value.decode())
