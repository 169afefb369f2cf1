"""Single GPU, diagnostic build (see scripts/dist_timeline.py): phase stamps of the one-GPU classic solver, to set
beside the row-partitioned one's (world = 1).  Medians over iterations 50..450, ns."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from nbots_b200 import api, capi, meshgen, multigpu
from util import flatten_bcs
import bench as B
L = capi.lib(); capi.check(L.nbgpu_init(0))
m = meshgen.structured_mesh(B.NX, B.NY_PER_GPU, 2.0, 1.0, kind=1)
_, _, A, d_b = multigpu._single_gpu_system(m, flatten_bcs(m, B.workload_bcs()))
N = A.N
d_x = api.DeviceBuffer.zeros(N)
for rep in range(2):
    capi.check(L.nbgpu_memset(d_x.ptr, 0, N * 8)); api.sync()
    it = C.c_uint32(0); res = C.c_double(0)
    api.timer_start()
    L.nbgpu_pcg_jacobi(A.h, d_b.ptr, d_x.ptr, 500, C.c_double(0.0), C.byref(it), C.byref(res))
    ms = api.timer_stop()
tl = np.zeros(512 * 10, dtype=np.uint64)
L.nbgpu_krylov_timeline.argtypes = [C.c_void_p]
capi.check(L.nbgpu_krylov_timeline(tl.ctypes.data))
t = tl.reshape(512, 10).astype(np.int64)[50:450]
nxt = tl.reshape(512, 10).astype(np.int64)[51:451]
med = lambda a: float(np.median(a))
print(json.dumps({"us_per_iter": round(ms * 1e3 / 500, 2),
                  "K1 start -> CTA0 rows done": med(t[:, 9] - t[:, 0]),
                  "K1 CTA0 rows done -> K2 CTA0 past wait": med(t[:, 3] - t[:, 9]),
                  "K2 sums K1's partials": med(t[:, 4] - t[:, 3]),
                  "K2 summed -> K3 CTA0 past wait": med(t[:, 6] - t[:, 4]),
                  "K3 sums K2's partials": med(t[:, 7] - t[:, 6]),
                  "K3 summed -> CTA0 done": med(t[:, 8] - t[:, 7]),
                  "K3 CTA0 done -> next K1 start": med(nxt[:, 0] - t[:, 8]),
                  "iteration": med(nxt[:, 0] - t[:, 0])}))
