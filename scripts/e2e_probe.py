"""GPU box: where does the time of the reference-facing (host-buffer) solve go?"""
import os, sys, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nbots_b200 import api, capi, meshgen
import bench

L = capi.lib(); capi.check(L.nbgpu_init(0))
m = bench.workload_mesh(1)
rs, cols = api.pattern_from_mesh(m)
K = api.Matrix.from_csr(rs, cols)
mesh = api.Mesh(m)
d_F = api.DeviceBuffer.zeros(K.N)
mesh.assemble(K, d_F, 1.0, 0.3)
from util import flatten_bcs
neu_dof, neu_add, dir_dof, dir_val = flatten_bcs(m, bench.workload_bcs())
api.vector_add_entries(d_F, neu_dof, neu_add)
K.apply_dirichlet(d_F, dir_dof, dir_val)
b = d_F.to_host(); tol = 1e-8 * np.linalg.norm(b); N = K.N
vals = K.values_csr()
A_host, keep = bench.host_nb_sparse(rs, cols, vals)

def t(f, n=3):
    f(); api.sync(); t0 = time.perf_counter()
    for _ in range(n): f()
    api.sync(); return (time.perf_counter() - t0) / n * 1e3

print("create_from_csr  %.1f ms" % t(lambda: api.Matrix.from_csr(rs, cols, vals).destroy()))
rows_c = keep[1].ctypes.data_as(C.POINTER(C.c_void_p)); rows_v = keep[0].ctypes.data_as(C.POINTER(C.c_void_p))
def mk():
    M = api.Matrix.from_row_pointers(N, rs.ctypes.data, rows_c, rows_v); M.destroy()
print("create_from_rows %.1f ms" % t(mk))
M = api.Matrix.from_csr(rs, cols, vals)
print("pcg host-buffers  %.1f ms" % t(lambda: M.pcg_jacobi_host(b, tol=tol)))
d_x = api.DeviceBuffer.zeros(N)
def res():
    capi.check(L.nbgpu_memset(d_x.ptr, 0, N * 8)); M.pcg_jacobi(d_F, d_x, tol=tol)
print("pcg resident      %.1f ms" % t(res))
shim = C.CDLL(capi.SHIM_PATH)
fn = shim.nb_sparse_solve_CG_precond_Jacobi
fn.restype = C.c_int
fn.argtypes = [C.POINTER(bench.NbSparse), capi.f64p, capi.f64p, C.c_uint32, C.c_double, capi.u32p, capi.f64p, C.c_uint32]
x = np.zeros(N)
def e2e():
    x[:] = 0; it = C.c_uint32(0); r = C.c_double(0)
    fn(C.byref(A_host), b.ctypes.data_as(capi.f64p), x.ctypes.data_as(capi.f64p), N, tol, C.byref(it), C.byref(r), 1)
print("shim e2e          %.1f ms" % t(e2e))
