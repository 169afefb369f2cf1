"""One rank of the row-partitioned path by itself (world = 1): its fixed overhead against the single-GPU path.
    python scripts/dist_one.py q4 [iters]      -- also the target of the ncu comparison of the two K1 variants"""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nbots_b200 import api, capi, meshgen, multigpu
from ab_pcg import SIZES, BCS
from util import flatten_bcs
L = capi.lib(); capi.check(L.nbgpu_init(0))
name = sys.argv[1] if len(sys.argv) > 1 else "q4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
nx, ny = SIZES[name]
m = meshgen.structured_mesh(nx, ny, 2.0, 2.0 * ny / nx, kind=1)
neu_dof, neu_add, dir_dof, dir_val = flatten_bcs(m, BCS)
ns = np.array([0, m.n_nod], dtype=np.uint32)
fem = multigpu.DistFem(m, 0, 1, ns, api.constitutive_matrix(1.0, 0.3), neu_dof, neu_add, dir_dof, dir_val, lambda o: [o])
fem.assemble()
best = 1e30
for rep in range(2):
    capi.check(L.nbgpu_memset(fem.d_x.ptr, 0, fem.N_loc * 8))
    api.sync(); api.timer_start()
    it = C.c_uint32(0); res = C.c_double(0)
    L.nbgpu_dist_pcg_jacobi(fem.dist, fem.plan, fem.A.h, fem.d_b.ptr, fem.d_x.ptr, iters, 0.0, C.byref(it), C.byref(res))
    best = min(best, api.timer_stop())
print(json.dumps({"workload": name, "path": "dist world=1", "us_per_iter": round(best * 1e3 / iters, 2)}))
