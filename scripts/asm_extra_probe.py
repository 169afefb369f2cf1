"""Call times (CUDA events around the C-ABI call, best of 5) of the plain GATHER assembly, the damage driver's loop
(incl. the upload of the per-Gauss-point damage field) and the lumped mass vector at Q1:  python scripts/asm_extra_probe.py"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from nbots_b200 import api, capi, meshgen
L = capi.lib(); capi.check(L.nbgpu_init(0))
m = meshgen.structured_mesh(1000, 500, 2.0, 1.0, kind=1)
mesh = api.Mesh(m)
K = mesh.create_matrix()
d_F = api.DeviceBuffer.zeros(K.N); d_M = api.DeviceBuffer.zeros(K.N)
dmg = np.random.default_rng(1).random(m.n_elems * 4) * 0.9


def best_of(fn, n=5):
    best = 1e30
    for _ in range(n):
        api.sync(); api.timer_start(); fn(); best = min(best, api.timer_stop())
    return round(best, 4)


out = {"workload": "Q1 (500 000 quads)",
       "gather_ms": best_of(lambda: mesh.assemble(K, d_F, 1.0, 0.3)),
       "damage_loop_ms_incl_16MB_upload": best_of(lambda: mesh.assemble(K, d_F, 1.0, 0.3, gp_damage=dmg)),
       "lumped_mass_ms": best_of(lambda: mesh.lumped_mass(d_M, 7.85, 1.0))}
M = d_M.to_host()
out["mass_total_over_expected"] = float(M[0::2].sum() / (7.85 * 2.0))
print(json.dumps(out))
