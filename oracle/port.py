"""ctypes client of oracle/libnboracle.so (plain-C restatement, oracle/port/*.c).

TEST INFRASTRUCTURE: checker and CPU baseline only, never on the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnboracle.so")

u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)


class BC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("where", C.c_int32), ("id", C.c_uint32), ("mask", C.c_int32 * 2),
                ("fn", C.c_int32), ("val", C.c_double * 2)]


_lib = None


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libnboracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.nbo_graph_nodes_by_elems.restype = C.c_uint64
        L.nbo_graph_nodes_by_elems.argtypes = [C.c_uint32, C.c_uint32, u32p, C.c_uint32, C.c_uint32, u32p, u32p, u32p]
        L.nbo_sparse_pattern.restype = C.c_uint64
        L.nbo_sparse_pattern.argtypes = [C.c_uint32, u32p, u32p, C.c_uint32, u32p, u32p]
        L.nbo_spmv.argtypes = [C.c_uint32, u64p, u32p, f64p, f64p, f64p, C.c_uint32]
        for n in ("nbo_pcg_jacobi", "nbo_cg"):
            f = getattr(L, n)
            f.restype = C.c_int
            f.argtypes = [C.c_uint32, u64p, u32p, f64p, f64p, f64p, C.c_uint32, C.c_double, u32p, f64p, C.c_uint32]
        L.nbo_dirichlet.argtypes = [u64p, u32p, f64p, f64p, C.c_uint32, C.c_double]
        L.nbo_elem_tables.argtypes = [C.c_int, u32p, u32p, f64p, f64p, f64p, f64p]
        L.nbo_constitutive.argtypes = [C.c_double, C.c_double, C.c_int, f64p]
        L.nbo_assemble.restype = C.c_int
        L.nbo_assemble.argtypes = [C.c_uint32, f64p, C.c_uint32, C.c_int, u32p, C.c_double, C.c_double, C.c_double,
                                   C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, u8p, u64p, u32p, f64p, f64p]
        L.nbo_assemble_damage.restype = C.c_int
        L.nbo_assemble_damage.argtypes = [C.c_uint32, f64p, C.c_uint32, C.c_int, u32p, C.c_double, C.c_double,
                                          C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, u8p, f64p,
                                          u64p, u32p, f64p, f64p]
        L.nbo_lumped_mass.restype = C.c_int
        L.nbo_lumped_mass.argtypes = [C.c_uint32, f64p, C.c_uint32, C.c_int, u32p, C.c_double, C.c_double, u8p, f64p]
        L.nbo_assemble_scaled.restype = C.c_int
        L.nbo_assemble_scaled.argtypes = [C.c_uint32, f64p, C.c_uint32, C.c_int, u32p, C.c_double, C.c_double,
                                          C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, u8p, f64p,
                                          u64p, u32p, f64p, f64p]
        L.nbo_set_bconditions.argtypes = [f64p, u32p, u32p, u32p, C.c_uint32, C.POINTER(BC), C.c_double, u64p, u32p,
                                          f64p, f64p]
        L.nbo_compute_strain.restype = C.c_int
        L.nbo_compute_strain.argtypes = [f64p, C.c_uint32, C.c_int, u32p, f64p, f64p]
        L.nbo_stress_from_strain.argtypes = [C.c_uint32, C.c_int, C.c_double, C.c_double, C.c_int, f64p, u8p, f64p]
        L.nbo_kirsch_stress.argtypes = [C.c_double, C.c_double, f64p]
        L.nbo_gp_to_nodes.restype = C.c_int
        L.nbo_gp_to_nodes.argtypes = [C.c_uint32, f64p, C.c_uint32, C.c_int, u32p, C.c_uint32, f64p, f64p]
        L.nbo_vm_stress.restype = C.c_double
        L.nbo_vm_stress.argtypes = [C.c_double, C.c_double, C.c_double]
        L.nbo_main_stress.argtypes = [C.c_double, C.c_double, C.c_double, f64p]
        _lib = L
    return _lib


def row_ptr_of(rows_size):
    rp = np.zeros(rows_size.size + 1, dtype=np.uint64)
    np.cumsum(rows_size, out=rp[1:])
    return rp


def graph_nodes_by_elems(m):
    L = lib()
    n_adj = np.zeros(m.n_nod, dtype=np.uint32)
    tot = L.nbo_graph_nodes_by_elems(m.n_nod, m.n_edg, _p(m.edg, u32p), m.n_elems, m.npe, _p(m.adj, u32p),
                                     _p(n_adj, u32p), None)
    adj = np.zeros(tot, dtype=np.uint32)
    L.nbo_graph_nodes_by_elems(m.n_nod, m.n_edg, _p(m.edg, u32p), m.n_elems, m.npe, _p(m.adj, u32p),
                               _p(n_adj, u32p), _p(adj, u32p))
    return n_adj, adj


def sparse_pattern(n_adj, adj_flat, vars_per_node):
    L = lib()
    n_adj = np.ascontiguousarray(n_adj, dtype=np.uint32)
    adj_flat = np.ascontiguousarray(adj_flat, dtype=np.uint32)
    rs = np.zeros(n_adj.size * vars_per_node, dtype=np.uint32)
    nnz = L.nbo_sparse_pattern(n_adj.size, _p(n_adj, u32p), _p(adj_flat, u32p), vars_per_node, _p(rs, u32p), None)
    cols = np.zeros(nnz, dtype=np.uint32)
    L.nbo_sparse_pattern(n_adj.size, _p(n_adj, u32p), _p(adj_flat, u32p), vars_per_node, _p(rs, u32p), _p(cols, u32p))
    return rs, cols


def pattern_from_mesh(m, vars_per_node=2):
    return sparse_pattern(*graph_nodes_by_elems(m), vars_per_node)


class Csr:
    """Flat CSR matrix of the port (rows_size/cols as the reference's rows, vals f64)."""

    def __init__(self, rows_size, cols, vals=None):
        self.rows_size = np.ascontiguousarray(rows_size, dtype=np.uint32)
        self.cols = np.ascontiguousarray(cols, dtype=np.uint32)
        self.vals = np.zeros(self.cols.size) if vals is None else np.array(vals, dtype=np.float64, copy=True)
        self.row_ptr = row_ptr_of(self.rows_size)

    @property
    def N(self):
        return self.rows_size.size

    @property
    def nnz(self):
        return self.cols.size

    def copy(self):
        return Csr(self.rows_size, self.cols, self.vals)

    def spmv(self, x, threads=1):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.N)
        lib().nbo_spmv(self.N, _p(self.row_ptr, u64p), _p(self.cols, u32p), _p(self.vals, f64p), _p(x, f64p),
                       _p(y, f64p), threads)
        return y

    def _solve(self, fn, b, x0, max_iter, tol, threads):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.zeros(self.N) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
        it = C.c_uint32(0)
        res = C.c_double(0)
        st = fn(self.N, _p(self.row_ptr, u64p), _p(self.cols, u32p), _p(self.vals, f64p), _p(b, f64p), _p(x, f64p),
                self.N if max_iter is None else max_iter, tol, C.byref(it), C.byref(res), threads)
        return st, x, it.value, res.value

    def pcg_jacobi(self, b, x0=None, max_iter=None, tol=1e-8, threads=1):
        return self._solve(lib().nbo_pcg_jacobi, b, x0, max_iter, tol, threads)

    def cg(self, b, x0=None, max_iter=None, tol=1e-8, threads=1):
        return self._solve(lib().nbo_cg, b, x0, max_iter, tol, threads)

    def dirichlet(self, rhs, idx, value):
        lib().nbo_dirichlet(_p(self.row_ptr, u64p), _p(self.cols, u32p), _p(self.vals, f64p), _p(rhs, f64p), int(idx),
                            float(value))


def elem_tables(elem_type):
    n = C.c_uint32(); g = C.c_uint32()
    w = np.zeros(4); Ni = np.zeros(16); dpsi = np.zeros(16); deta = np.zeros(16)
    lib().nbo_elem_tables(elem_type, C.byref(n), C.byref(g), _p(w, f64p), _p(Ni, f64p), _p(dpsi, f64p), _p(deta, f64p))
    k = n.value * g.value
    return n.value, g.value, w[:g.value].copy(), Ni[:k].copy(), dpsi[:k].copy(), deta[:k].copy()


def constitutive(E, nu, analysis):
    D = np.zeros(4)
    lib().nbo_constitutive(E, nu, analysis, _p(D, f64p))
    return D


def assemble(K: Csr, m, E, nu, density=0.0, self_weight=False, gravity=(0.0, 0.0), analysis=0, thickness=1.0,
             enabled=None, elem_scale=None, gp_damage=None):
    F = np.zeros(K.N)
    en = None if enabled is None else np.ascontiguousarray(enabled, dtype=np.uint8)
    if gp_damage is not None:      # the damage driver's loop (static_damage2D.c:474-569)
        dm = np.ascontiguousarray(gp_damage, dtype=np.float64)
        st = lib().nbo_assemble_damage(m.n_nod, _p(m.nod, f64p), m.n_elems, m.kind, _p(m.adj, u32p), E, nu, density,
                                       int(self_weight), gravity[0], gravity[1], analysis, thickness, _p(en, u8p),
                                       _p(dm, f64p), _p(K.row_ptr, u64p), _p(K.cols, u32p), _p(K.vals, f64p),
                                       _p(F, f64p))
        return st, F
    if elem_scale is not None:
        sc = np.ascontiguousarray(elem_scale, dtype=np.float64)
        st = lib().nbo_assemble_scaled(m.n_nod, _p(m.nod, f64p), m.n_elems, m.kind, _p(m.adj, u32p), E, nu, density,
                                       int(self_weight), gravity[0], gravity[1], analysis, thickness, _p(en, u8p),
                                       _p(sc, f64p), _p(K.row_ptr, u64p), _p(K.cols, u32p), _p(K.vals, f64p),
                                       _p(F, f64p))
        return st, F
    st = lib().nbo_assemble(m.n_nod, _p(m.nod, f64p), m.n_elems, m.kind, _p(m.adj, u32p), E, nu, density,
                            int(self_weight), gravity[0], gravity[1], analysis, thickness, _p(en, u8p),
                            _p(K.row_ptr, u64p), _p(K.cols, u32p), _p(K.vals, f64p), _p(F, f64p))
    return st, F


def lumped_mass(m, density, thickness=1.0, enabled=None):
    """M of pipeline_assemble_system(K, M != NULL, ...) (pipeline.c:216-222, :256-259)."""
    M = np.zeros(2 * m.n_nod)
    en = None if enabled is None else np.ascontiguousarray(enabled, dtype=np.uint8)
    st = lib().nbo_lumped_mass(m.n_nod, _p(m.nod, f64p), m.n_elems, m.kind, _p(m.adj, u32p), density, thickness,
                               _p(en, u8p), _p(M, f64p))
    return st, M


def make_bcs(records):
    """records: iterable of (kind 'dirichlet'|'neumann', where 'vtx'|'sgm', id, mask(2), val(2)[, fn])."""
    arr = (BC * max(len(records), 1))()
    for k, r in enumerate(records):
        arr[k].kind = 0 if r[0] == "dirichlet" else 1
        arr[k].where = 0 if r[1] == "vtx" else 1
        arr[k].id = r[2]
        arr[k].mask[0], arr[k].mask[1] = int(r[3][0]), int(r[3][1])
        arr[k].val[0], arr[k].val[1] = float(r[4][0]), float(r[4][1])
        arr[k].fn = r[5] if len(r) > 5 else 0
    return arr, len(records)


def set_bconditions(m, K: Csr, F, records, factor=1.0):
    arr, n = make_bcs(records)
    lib().nbo_set_bconditions(_p(m.nod, f64p), _p(m.vtx, u32p), _p(m.sgm_sizes, u32p), _p(m.sgm_nodes, u32p), n, arr,
                              factor, _p(K.row_ptr, u64p), _p(K.cols, u32p), _p(K.vals, f64p), _p(F, f64p))


def compute_strain(m, disp):
    ngp = 4 if m.kind else 1
    disp = np.ascontiguousarray(disp, dtype=np.float64)
    strain = np.zeros(3 * ngp * m.n_elems)
    st = lib().nbo_compute_strain(_p(m.nod, f64p), m.n_elems, m.kind, _p(m.adj, u32p), _p(disp, f64p),
                                  _p(strain, f64p))
    return st, strain


def gp_to_nodes(m, n_comp, gp_values):
    gp_values = np.ascontiguousarray(gp_values, dtype=np.float64)
    out = np.zeros(m.n_nod * n_comp)
    st = lib().nbo_gp_to_nodes(m.n_nod, _p(m.nod, f64p), m.n_elems, m.kind, _p(m.adj, u32p), n_comp,
                               _p(gp_values, f64p), _p(out, f64p))
    return st, out


def vm_stress(stress):
    s = np.asarray(stress, dtype=np.float64).reshape(-1, 3)
    return np.array([lib().nbo_vm_stress(a, b, c) for a, b, c in s])


def main_stress(stress):
    s = np.asarray(stress, dtype=np.float64).reshape(-1, 3)
    out = np.zeros((s.shape[0], 2))
    tmp = np.zeros(2)
    for i, (a, b, c) in enumerate(s):
        lib().nbo_main_stress(a, b, c, _p(tmp, f64p))
        out[i] = tmp
    return out.ravel()


def stress_from_strain(n_elems, elem_type, E, nu, analysis, strain, enabled=None):
    strain = np.ascontiguousarray(strain, dtype=np.float64)
    stress = np.zeros_like(strain)
    en = None if enabled is None else np.ascontiguousarray(enabled, dtype=np.uint8)
    lib().nbo_stress_from_strain(n_elems, elem_type, E, nu, analysis, _p(strain, f64p), _p(en, u8p), _p(stress, f64p))
    return stress
