"""Assembly timings on the device (CUDA events, best of 5) per schedule: python scripts/asm_probe.py q1 [q16]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from nbots_b200 import api, capi, meshgen
from ab_pcg import SIZES
L = capi.lib(); capi.check(L.nbgpu_init(0))
for name in sys.argv[1:] or ["q1"]:
    nx, ny = SIZES[name]
    m = meshgen.structured_mesh(nx, ny, 2.0, 2.0 * ny / nx, kind=1)
    rs, cols = api.pattern_from_mesh(m)
    K = api.Matrix.from_csr(rs, cols); mesh = api.Mesh(m); d_F = api.DeviceBuffer.zeros(K.N)
    ref_vals = None
    for label, mode in (("gather", capi.ASSEMBLY_GATHER), ("atomic", capi.ASSEMBLY_ATOMIC), ("color", capi.ASSEMBLY_COLOR)):
        best = 1e30
        for rep in range(5):
            api.sync(); api.timer_start()
            st, _ = mesh.assemble(K, d_F, 1.0, 0.3, thickness=1.0, mode=mode)
            best = min(best, api.timer_stop())
        vals = K.values_csr()
        if ref_vals is None:
            ref_vals = vals
        bytes_alg = 8 * K.nnz + 16 * m.n_nod + 16 * m.n_elems
        print(json.dumps({"workload": name, "schedule": label + (" (rows)" if os.environ.get("NBGPU_ASSEMBLY_ROWS") and mode == 0 else ""),
                          "ms": round(best, 4), "elems_per_s": m.n_elems / (best * 1e-3),
                          "GBps_algorithmic": round(bytes_alg / (best * 1e-3) / 1e9, 1),
                          "bit_identical_to_gather": bool(np.array_equal(vals, ref_vals)),
                          "rel_to_gather": float(np.linalg.norm(vals - ref_vals) / np.linalg.norm(ref_vals))}), flush=True)
    K.destroy(); mesh.destroy(); d_F.free()
