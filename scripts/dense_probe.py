"""Does a fixed iteration budget from x0 = 0 time the same work as a full solve?  (Most of p, g are exactly 0
until the load's influence has crossed the mesh: one grid line per iteration.)"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from nbots_b200 import api, capi
from ab_pcg import build, SIZES
L = capi.lib(); capi.check(L.nbgpu_init(0))
name = sys.argv[1]
K, d_F = build(*SIZES[name]); N = K.N
b = d_F.to_host(); tol = 1e-8 * float(np.linalg.norm(b))
rng = np.random.default_rng(1)
x_dense = rng.standard_normal(N)
d_x = api.DeviceBuffer.zeros(N)
for mode in (0, 1):
    capi.check(L.nbgpu_set_pcg_mode(mode))
    for label, x0, mi, t in (("x0=0, 600 its", None, 600, 0.0), ("x0=dense random, 600 its", x_dense, 600, 0.0),
                             ("x0=0, full solve", None, N, tol)):
        for rep in range(2):
            if x0 is None:
                capi.check(L.nbgpu_memset(d_x.ptr, 0, N * 8))
            else:
                d_x.upload(x0)
            api.timer_start(); st, it, res = K.pcg_jacobi(d_F, d_x, max_iter=mi, tol=t); ms = api.timer_stop()
        print(json.dumps({"workload": name, "mode": "fused" if mode else "classic", "case": label, "iterations": it,
                          "us_per_iter": round(ms * 1e3 / it, 2)}), flush=True)
d_in = api.DeviceBuffer.from_host(x_dense); d_out = api.DeviceBuffer.zeros(N)
for _ in range(5): K.spmv(d_in, d_out)
api.timer_start()
for _ in range(50): K.spmv(d_in, d_out)
print(json.dumps({"workload": name, "spmv_us": round(api.timer_stop() * 1e3 / 50, 2)}))
