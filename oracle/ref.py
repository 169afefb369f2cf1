"""ctypes client of oracle/_ref/libnbots_ref.so (the UNMODIFIED reference + oracle/ref_harness.c).

TEST INFRASTRUCTURE: checker and CPU baseline only, never on the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libnbots_ref.so")

_lib = None

u32p = C.POINTER(C.c_uint32)
f64p = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _p(a, t):
    if a is None:
        return None
    return a.ctypes.data_as(t)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)
        L.refh_mesh_create.restype = C.c_void_p
        L.refh_mesh_create.argtypes = [C.c_int, C.c_uint32, f64p, C.c_uint32, u32p, C.c_uint32, u32p,
                                       C.c_uint32, u32p, C.c_uint32, u32p, u32p]
        L.refh_mesh_from_model.restype = C.c_void_p
        L.refh_mesh_from_model.argtypes = [C.c_uint32, f64p, C.c_uint32, u32p, C.c_uint32, f64p, C.c_uint32]
        L.refh_mesh_counts.argtypes = [C.c_void_p, u32p]
        L.refh_mesh_export.argtypes = [C.c_void_p, f64p, u32p, u32p, u32p, u32p, u32p]
        L.refh_mesh_destroy.argtypes = [C.c_void_p]
        L.refh_mesh_ptr.restype = C.c_void_p
        L.refh_mesh_ptr.argtypes = [C.c_void_p]
        L.refh_sparse_from_mesh.restype = C.c_void_p
        L.refh_sparse_from_mesh.argtypes = [C.c_void_p]
        L.refh_graph_from_mesh.restype = C.c_uint64
        L.refh_graph_from_mesh.argtypes = [C.c_void_p, u32p, u32p]
        L.refh_sparse_from_graph.restype = C.c_void_p
        L.refh_sparse_from_graph.argtypes = [C.c_uint32, u32p, u32p, C.c_uint32]
        L.refh_sparse_N.restype = C.c_uint32
        L.refh_sparse_N.argtypes = [C.c_void_p]
        L.refh_sparse_nnz.restype = C.c_uint64
        L.refh_sparse_nnz.argtypes = [C.c_void_p]
        L.refh_sparse_export.argtypes = [C.c_void_p, u32p, u32p, f64p]
        L.refh_sparse_import_values.argtypes = [C.c_void_p, f64p]
        L.refh_sparse_destroy.argtypes = [C.c_void_p]
        L.refh_spmv.argtypes = [C.c_void_p, f64p, f64p, C.c_uint32]
        for name in ("refh_pcg_jacobi", "refh_cg"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, f64p, f64p, C.c_uint32, C.c_double, u32p, f64p, C.c_uint32]
        L.refh_dirichlet.argtypes = [C.c_void_p, f64p, C.c_uint32, C.c_double]
        L.refh_elem_tables.argtypes = [C.c_int, u32p, u32p, f64p, f64p, f64p, f64p]
        L.refh_constitutive.argtypes = [C.c_double, C.c_double, C.c_int, f64p]
        L.refh_assemble.restype = C.c_int
        L.refh_assemble.argtypes = [C.c_void_p, f64p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                                    C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, u8p]
        L.refh_assemble_mass.restype = C.c_int
        L.refh_assemble_mass.argtypes = [C.c_void_p, f64p, f64p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                                         C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, u8p]
        L.refh_assemble_damage.restype = None
        L.refh_assemble_damage.argtypes = [C.c_void_p, f64p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double,
                                           C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, f64p, u8p]
        L.refh_bcond_create.restype = C.c_void_p
        L.refh_bcond_destroy.argtypes = [C.c_void_p]
        L.refh_bcond_push.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int,
                                      C.c_double, C.c_double]
        L.refh_bcond_push_kirsch.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        L.refh_kirsch_stress.argtypes = [C.c_double, C.c_double, f64p]
        L.refh_set_bconditions.argtypes = [C.c_void_p, C.c_void_p, f64p, C.c_void_p, C.c_double]
        L.refh_fem_static.restype = C.c_int
        L.refh_fem_static.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_void_p,
                                      C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, u8p, f64p, f64p]
        L.refh_compute_strain.argtypes = [C.c_void_p, C.c_int, f64p, f64p]
        L.refh_sparse_save.argtypes = [C.c_void_p, C.c_char_p]
        L.refh_sparse_save_mat4.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.refh_mat4_save_vec.argtypes = [C.c_char_p, C.c_char_p, f64p, C.c_uint32]
        L.refh_mesh_save_vtk.restype = C.c_int
        L.refh_mesh_save_vtk.argtypes = [C.c_void_p, C.c_char_p]
        L.refh_gp_to_nodes.restype = C.c_int
        L.refh_gp_to_nodes.argtypes = [C.c_void_p, C.c_int, C.c_uint32, f64p, f64p]
        L.refh_vm_stress.argtypes = [C.c_uint32, f64p, f64p]
        L.refh_main_stress.argtypes = [C.c_uint32, f64p, f64p]
        L.refh_stress_from_strain.argtypes = [C.c_uint32, C.c_int, C.c_double, C.c_double, C.c_int, f64p,
                                              u8p, f64p]
        L.refh_mesh_save_nbt.restype = C.c_int
        L.refh_mesh_save_nbt.argtypes = [C.c_void_p, C.c_char_p]
        L.refh_mesh_read_type_nbt.restype = C.c_int
        L.refh_mesh_read_type_nbt.argtypes = [C.c_char_p, C.POINTER(C.c_int)]
        L.refh_mesh_read_nbt.restype = C.c_int
        L.refh_mesh_read_nbt.argtypes = [C.c_void_p, C.c_char_p]
        L.refh_inv_power.restype = C.c_int
        L.refh_inv_power.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, f64p, f64p, C.POINTER(C.c_int),
                                     C.c_double, C.c_uint32]
        _lib = L
    return _lib


class RefMesh:
    """A genuine reference ``nb_mesh2D_t`` built from flat arrays (or by the reference mesher)."""

    def __init__(self, handle, kind):
        self.h = handle
        self.kind = kind

    @classmethod
    def from_arrays(cls, m):
        L = lib()
        h = L.refh_mesh_create(m.kind, m.n_nod, _p(m.nod, f64p), m.n_edg, _p(m.edg, u32p), m.n_elems,
                               _p(m.adj, u32p), m.vtx.size, _p(m.vtx, u32p), m.sgm_sizes.size,
                               _p(m.sgm_sizes, u32p), _p(m.sgm_nodes, u32p))
        return cls(h, m.kind)

    @classmethod
    def from_model(cls, vertex, edge, holes, max_vtx):
        L = lib()
        vertex = np.ascontiguousarray(vertex, dtype=np.float64)
        edge = np.ascontiguousarray(edge, dtype=np.uint32)
        holes = np.ascontiguousarray(holes, dtype=np.float64)
        h = L.refh_mesh_from_model(vertex.size // 2, _p(vertex, f64p), edge.size // 2, _p(edge, u32p),
                                   holes.size // 2, _p(holes, f64p) if holes.size else None, max_vtx)
        return cls(h, 0)

    def export(self):
        """-> nbots_b200.meshgen.Mesh2D with the mesh's flat arrays."""
        from nbots_b200.meshgen import Mesh2D
        L = lib()
        cnt = np.zeros(6, dtype=np.uint32)
        L.refh_mesh_counts(self.h, _p(cnt, u32p))
        npe = 4 if self.kind else 3
        nod = np.zeros(2 * cnt[0]); edg = np.zeros(2 * cnt[1], dtype=np.uint32)
        adj = np.zeros(npe * cnt[2], dtype=np.uint32); vtx = np.zeros(cnt[3], dtype=np.uint32)
        ss = np.zeros(cnt[4], dtype=np.uint32); sn = np.zeros(cnt[5], dtype=np.uint32)
        L.refh_mesh_export(self.h, _p(nod, f64p), _p(edg, u32p), _p(adj, u32p), _p(vtx, u32p),
                           _p(ss, u32p), _p(sn, u32p))
        return Mesh2D(kind=self.kind, nod=nod, edg=edg, adj=adj, vtx=vtx, sgm_sizes=ss, sgm_nodes=sn)

    def graph(self):
        L = lib()
        cnt = np.zeros(6, dtype=np.uint32)
        L.refh_mesh_counts(self.h, _p(cnt, u32p))
        n_adj = np.zeros(cnt[0], dtype=np.uint32)
        tot = L.refh_graph_from_mesh(self.h, _p(n_adj, u32p), None)
        adj = np.zeros(tot, dtype=np.uint32)
        L.refh_graph_from_mesh(self.h, _p(n_adj, u32p), _p(adj, u32p))
        return n_adj, adj

    def close(self):
        if self.h:
            lib().refh_mesh_destroy(self.h)
            self.h = None


class RefSparse:
    """A genuine reference ``nb_sparse_t``."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def from_mesh(cls, mesh: RefMesh):
        return cls(lib().refh_sparse_from_mesh(mesh.h))

    @classmethod
    def from_graph(cls, n_adj, adj_flat, vars_per_node):
        n_adj = np.ascontiguousarray(n_adj, dtype=np.uint32)
        adj_flat = np.ascontiguousarray(adj_flat, dtype=np.uint32)
        return cls(lib().refh_sparse_from_graph(n_adj.size, _p(n_adj, u32p), _p(adj_flat, u32p), vars_per_node))

    @classmethod
    def from_csr(cls, rows_size, cols, vals):
        """Build through nb_sparse_create from the pattern's own graph, then import values."""
        rows_size = np.asarray(rows_size, dtype=np.uint32)
        cols = np.asarray(cols, dtype=np.uint32)
        row_of = np.repeat(np.arange(rows_size.size, dtype=np.uint32), rows_size)
        keep = cols != row_of
        A = cls.from_graph(rows_size - 1, cols[keep], 1)
        rs, cc, _ = A.export()
        assert np.array_equal(rs, rows_size) and np.array_equal(cc, cols)
        A.set_values(vals)
        return A

    @property
    def N(self):
        return lib().refh_sparse_N(self.h)

    @property
    def nnz(self):
        return lib().refh_sparse_nnz(self.h)

    def export(self):
        L = lib()
        rs = np.zeros(self.N, dtype=np.uint32)
        cols = np.zeros(self.nnz, dtype=np.uint32)
        vals = np.zeros(self.nnz, dtype=np.float64)
        L.refh_sparse_export(self.h, _p(rs, u32p), _p(cols, u32p), _p(vals, f64p))
        return rs, cols, vals

    def set_values(self, vals):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        assert vals.size == self.nnz
        lib().refh_sparse_import_values(self.h, _p(vals, f64p))

    def spmv(self, x, threads=1):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.N)
        lib().refh_spmv(self.h, _p(x, f64p), _p(y, f64p), threads)
        return y

    def _solve(self, fn, b, x0, max_iter, tol, threads):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.array(x0, dtype=np.float64, copy=True) if x0 is not None else np.zeros(self.N)
        it = C.c_uint32(0)
        res = C.c_double(0)
        st = fn(self.h, _p(b, f64p), _p(x, f64p), max_iter, tol, C.byref(it), C.byref(res), threads)
        return st, x, it.value, res.value

    def pcg_jacobi(self, b, x0=None, max_iter=None, tol=1e-8, threads=1):
        return self._solve(lib().refh_pcg_jacobi, b, x0, self.N if max_iter is None else max_iter, tol, threads)

    def cg(self, b, x0=None, max_iter=None, tol=1e-8, threads=1):
        return self._solve(lib().refh_cg, b, x0, self.N if max_iter is None else max_iter, tol, threads)

    def dirichlet(self, rhs, idx, value):
        lib().refh_dirichlet(self.h, _p(rhs, f64p), int(idx), float(value))

    def close(self):
        if self.h:
            lib().refh_sparse_destroy(self.h)
            self.h = None


class RefBcond:
    def __init__(self):
        self.h = lib().refh_bcond_create()

    def push(self, kind, where, ident, mask, val):
        """kind 'dirichlet'|'neumann', where 'vtx'|'sgm'."""
        lib().refh_bcond_push(self.h, 0 if kind == "dirichlet" else 1, 0 if where == "vtx" else 1, ident,
                              int(mask[0]), int(mask[1]), float(val[0]), float(val[1]))

    def push_kirsch(self, sgm, which):
        lib().refh_bcond_push_kirsch(self.h, sgm, which)

    def close(self):
        if self.h:
            lib().refh_bcond_destroy(self.h)
            self.h = None


def elem_tables(elem_type):
    n = C.c_uint32(); g = C.c_uint32()
    w = np.zeros(4); Ni = np.zeros(16); dpsi = np.zeros(16); deta = np.zeros(16)
    lib().refh_elem_tables(elem_type, C.byref(n), C.byref(g), _p(w, f64p), _p(Ni, f64p), _p(dpsi, f64p),
                           _p(deta, f64p))
    k = n.value * g.value
    return n.value, g.value, w[:g.value].copy(), Ni[:k].copy(), dpsi[:k].copy(), deta[:k].copy()


def constitutive(E, nu, analysis):
    D = np.zeros(4)
    lib().refh_constitutive(E, nu, analysis, _p(D, f64p))
    return D


def assemble(K: RefSparse, mesh: RefMesh, elem_type, E, nu, density=0.0, self_weight=False, gravity=(0.0, 0.0),
             analysis=0, thickness=1.0, enabled=None):
    F = np.zeros(K.N)
    en = None if enabled is None else np.ascontiguousarray(enabled, dtype=np.uint8)
    st = lib().refh_assemble(K.h, _p(F, f64p), mesh.h, elem_type, E, nu, density, int(self_weight),
                             gravity[0], gravity[1], analysis, thickness, _p(en, u8p))
    return st, F


def assemble_with_mass(K: RefSparse, mesh: RefMesh, elem_type, E, nu, density=0.0, self_weight=False,
                       gravity=(0.0, 0.0), analysis=0, thickness=1.0, enabled=None):
    """pipeline_assemble_system with M != NULL: returns (status, F, M)."""
    F = np.zeros(K.N)
    M = np.zeros(K.N)
    en = None if enabled is None else np.ascontiguousarray(enabled, dtype=np.uint8)
    st = lib().refh_assemble_mass(K.h, _p(M, f64p), _p(F, f64p), mesh.h, elem_type, E, nu, density, int(self_weight),
                                  gravity[0], gravity[1], analysis, thickness, _p(en, u8p))
    return st, F, M


def assemble_damage(K: RefSparse, mesh: RefMesh, elem_type, E, nu, gp_damage, density=0.0, self_weight=False,
                    gravity=(0.0, 0.0), analysis=0, thickness=1.0, enabled=None):
    """DMG_pipeline_assemble_system (static_damage2D.c:474-569): D of every Gauss point scaled by 1 - damage."""
    F = np.zeros(K.N)
    dm = None if gp_damage is None else np.ascontiguousarray(gp_damage, dtype=np.float64)
    en = None if enabled is None else np.ascontiguousarray(enabled, dtype=np.uint8)
    lib().refh_assemble_damage(K.h, _p(F, f64p), mesh.h, elem_type, E, nu, density, int(self_weight), gravity[0],
                               gravity[1], analysis, thickness, _p(dm, f64p), _p(en, u8p))
    return F


def set_bconditions(mesh: RefMesh, K: RefSparse, F, bc: RefBcond, factor=1.0):
    lib().refh_set_bconditions(mesh.h, K.h, _p(F, f64p), bc.h, factor)


def fem_static(mesh: RefMesh, elem_type, E, nu, bc: RefBcond, density=0.0, self_weight=False,
               gravity=(0.0, 0.0), analysis=0, thickness=1.0, enabled=None, n_nod=None, n_elems=None, n_gp=None):
    disp = np.zeros(2 * n_nod)
    strain = np.zeros(3 * n_gp * n_elems)
    en = None if enabled is None else np.ascontiguousarray(enabled, dtype=np.uint8)
    st = lib().refh_fem_static(mesh.h, elem_type, E, nu, density, bc.h, int(self_weight), gravity[0], gravity[1],
                               analysis, thickness, _p(en, u8p), _p(disp, f64p), _p(strain, f64p))
    return st, disp, strain


def compute_strain(mesh: RefMesh, elem_type, disp, n_elems, n_gp):
    disp = np.ascontiguousarray(disp, dtype=np.float64)
    strain = np.zeros(3 * n_gp * n_elems)
    lib().refh_compute_strain(mesh.h, elem_type, _p(disp, f64p), _p(strain, f64p))
    return strain


def gp_to_nodes(mesh: RefMesh, elem_type, n_nod, n_comp, gp_values):
    gp_values = np.ascontiguousarray(gp_values, dtype=np.float64)
    out = np.zeros(n_nod * n_comp)
    st = lib().refh_gp_to_nodes(mesh.h, elem_type, n_comp, _p(gp_values, f64p), _p(out, f64p))
    return st, out


def vm_stress(stress):
    stress = np.ascontiguousarray(stress, dtype=np.float64)
    vm = np.zeros(stress.size // 3)
    lib().refh_vm_stress(vm.size, _p(stress, f64p), _p(vm, f64p))
    return vm


def main_stress(stress):
    stress = np.ascontiguousarray(stress, dtype=np.float64)
    out = np.zeros(2 * (stress.size // 3))
    lib().refh_main_stress(stress.size // 3, _p(stress, f64p), _p(out, f64p))
    return out


def stress_from_strain(n_elems, elem_type, E, nu, analysis, strain, enabled=None):
    strain = np.ascontiguousarray(strain, dtype=np.float64)
    stress = np.zeros_like(strain)
    en = None if enabled is None else np.ascontiguousarray(enabled, dtype=np.uint8)
    lib().refh_stress_from_strain(n_elems, elem_type, E, nu, analysis, _p(strain, f64p), _p(en, u8p),
                                  _p(stress, f64p))
    return stress


def inv_power(A: "RefSparse", h, mu=0.0, tolerance=1e-8, use_jacobi=True, threads=1):
    """nb_sparse_eigen_ipower (eigen/inv_power.c:13-130) with the Krylov solvers inside -> (status, vecs[h][N], vals, its)."""
    vecs = np.zeros((h, A.N)); vals = np.zeros(h); its = (C.c_int * h)()
    st = lib().refh_inv_power(A.h, int(use_jacobi), h, float(mu), _p(vecs, f64p), _p(vals, f64p), its, tolerance, threads)
    return st, vecs, vals, list(its)
