"""Warm timings of mesh upload, device pattern + matrix creation, host pattern path: python scripts/pattern_probe.py q1 q16"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
from nbots_b200 import api, capi, meshgen
from ab_pcg import SIZES
L = capi.lib(); capi.check(L.nbgpu_init(0))
for name in sys.argv[1:] or ["q1"]:
    nx, ny = SIZES[name]
    m = meshgen.structured_mesh(nx, ny, 2.0, 2.0 * ny / nx, kind=1)
    out = {"workload": name, "N_dof": 2 * m.n_nod}
    for rep in range(3):
        api.sync(); t0 = time.perf_counter(); mesh = api.Mesh(m); api.sync(); out["mesh_upload_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
        t0 = time.perf_counter(); K = mesh.create_matrix(); api.sync(); out["device_pattern_and_matrix_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
        n_colors, _ = (0, None)
        if rep == 2:
            t0 = time.perf_counter(); n_colors, colors = mesh.coloring(); out["device_colouring_ms"] = round((time.perf_counter() - t0) * 1e3, 2); out["n_colors"] = n_colors
        K.destroy(); mesh.destroy()
    t0 = time.perf_counter(); rs, cols = api.pattern_from_mesh(m); out["host_pattern_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
    for rep in range(2):
        t0 = time.perf_counter(); K = api.Matrix.from_csr(rs, cols); api.sync(); out["host_path_matrix_upload_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
        K.destroy()
    print(json.dumps(out), flush=True)
