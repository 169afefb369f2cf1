"""GPU box: time the standalone SpMV and a fixed-iteration PCG on Q1 (and optionally L-size Laplacians)
under the current NBGPU_* environment; prints one line per configuration."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nbots_b200 import api, capi, meshgen
import ctypes as C

L = capi.lib(); capi.check(L.nbgpu_init(0))
which = sys.argv[1] if len(sys.argv) > 1 else "q1"
if which[0] in "qt":
    nx, ny = {"1": (1000, 500), "4": (2000, 1000), "16": (4000, 2000)}[which[1:]]
    # t<n>: triangles with random diagonals -- ragged rows like an unstructured mesh
    m = meshgen.structured_mesh(nx, ny, 2.0, 1.0) if which[0] == "q" else \
        meshgen.structured_mesh(nx, ny, 2.0, 1.0, kind=0, diagonal_seed=1)
    rs, cols = api.pattern_from_mesh(m)
    K = api.Matrix.from_csr(rs, cols)
    mesh = api.Mesh(m)
    d_F = api.DeviceBuffer.zeros(K.N)
    mesh.assemble(K, d_F, 1.0, 0.3)
    left = m.segment(3)
    dofs = np.concatenate([2 * left, 2 * left + 1]).astype(np.uint32)
    K.apply_dirichlet(d_F, dofs, np.zeros(dofs.size))
    b = meshgen.uniform_rhs(K.N)
else:
    n = int(which[1:])
    rs, cols, vals = meshgen.laplacian9_csr(n)
    K = api.Matrix.from_csr(rs, cols, vals)
    b = meshgen.uniform_rhs(K.N)
N, nnz = K.N, K.nnz
d_b = api.DeviceBuffer.from_host(b)
d_x = api.DeviceBuffer.zeros(N); d_y = api.DeviceBuffer.zeros(N)
bytes_spmv = 12 * nnz + 20 * N + 4
for _ in range(5):
    K.spmv(d_b, d_y)
api.sync(); api.timer_start()
reps = 50
for _ in range(reps):
    K.spmv(d_b, d_y)
ms = api.timer_stop() / reps
capi.check(L.nbgpu_krylov_profile(1))
capi.check(L.nbgpu_memset(d_x.ptr, 0, N * 8))
api.timer_start()
st, it, res = K.pcg_jacobi(d_b, d_x, max_iter=200, tol=0.0)
ms_pcg = api.timer_stop()
ms3 = np.zeros(3); n_prof = C.c_uint32(0)
capi.check(L.nbgpu_krylov_profile_get(ms3.ctypes.data_as(capi.f64p), C.byref(n_prof)))
per = ms3 / n_prof.value
capi.check(L.nbgpu_krylov_profile(0))
capi.check(L.nbgpu_memset(d_x.ptr, 0, N * 8))
api.timer_start()
st, it, res = K.pcg_jacobi(d_b, d_x, max_iter=400, tol=0.0)
ms_pcg = api.timer_stop()
env = {k: v for k, v in os.environ.items() if k.startswith("NBGPU_")}
print(f"{which} N={N} nnz={nnz} sigma={K.sigma} stored/nnz={K.stored/nnz:.3f} env={env} | spmv {ms*1e3:.1f} us = {bytes_spmv/ms/1e6:.0f} GB/s ({bytes_spmv/ms/1e6/6551.7*100:.1f}%) | "
      f"K1 {per[0]*1e3:.1f} us ({bytes_spmv/per[0]/1e6:.0f} GB/s) K2 {per[1]*1e3:.1f} K3 {per[2]*1e3:.1f} | "
      f"pcg {ms_pcg/it*1e3:.1f} us/iter = {N*it/ms_pcg/1e6:.2f} GDOFit/s")
