#!/usr/bin/env python
"""A/B of the two PCG formulations on one GPU: CLASSIC (3 kernels, 2 reductions) vs FUSED (2 kernels, 1 reduction).

  python scripts/ab_pcg.py [q1|q4|q16 ...]     one JSON line per (workload, mode)
Full solve at Q1 (iteration counts, field agreement), fixed 600-iteration budgets at Q4 / Q16, per-kernel CUDA-event
times from nbgpu_krylov_profile."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from nbots_b200 import api, capi, meshgen  # noqa: E402
from util import flatten_bcs  # noqa: E402

SIZES = {"q1": (1000, 500), "q4": (2000, 1000), "q16": (4000, 2000), "q025": (500, 250)}
BCS = [("dirichlet", "sgm", 3, (1, 1), (0.0, 0.0)), ("neumann", "sgm", 1, (1, 1), (0.0, -1.0))]


def build(nx, ny):
    m = meshgen.structured_mesh(nx, ny, 2.0, 2.0 * ny / nx, kind=1)
    rs, cols = api.pattern_from_mesh(m)
    K = api.Matrix.from_csr(rs, cols)
    mesh = api.Mesh(m)
    d_F = api.DeviceBuffer.zeros(K.N)
    st, _ = mesh.assemble(K, d_F, 1.0, 0.3, thickness=1.0)
    assert st == 0
    neu_dof, neu_add, dir_dof, dir_val = flatten_bcs(m, BCS)
    api.vector_add_entries(d_F, neu_dof, neu_add)
    K.apply_dirichlet(d_F, dir_dof, dir_val)
    return K, d_F


def main():
    L = capi.lib()
    capi.check(L.nbgpu_init(0))
    for name in (sys.argv[1:] or ["q1"]):
        nx, ny = SIZES[name]
        K, d_F = build(nx, ny)
        N = K.N
        b = d_F.to_host()
        tol = 1e-8 * float(np.linalg.norm(b))
        full = name in ("q1", "q025")
        xs = {}
        for mode in (0, 1):
            capi.check(L.nbgpu_set_pcg_mode(mode))
            d_x = api.DeviceBuffer.zeros(N)
            best, it = 1e30, 0
            for rep in range(3 if full else 2):
                capi.check(L.nbgpu_memset(d_x.ptr, 0, N * 8))
                api.timer_start()
                st, it, res = K.pcg_jacobi(d_F, d_x, max_iter=N if full else 600, tol=tol if full else 0.0)
                best = min(best, api.timer_stop())
            xs[mode] = d_x.to_host()
            capi.check(L.nbgpu_krylov_profile(1))
            capi.check(L.nbgpu_memset(d_x.ptr, 0, N * 8))
            K.pcg_jacobi(d_F, d_x, max_iter=256, tol=0.0)
            capi.check(L.nbgpu_krylov_profile(0))
            ms3 = np.zeros(3); n_prof = C.c_uint32(0)
            capi.check(L.nbgpu_krylov_profile_get(ms3.ctypes.data_as(capi.f64p), C.byref(n_prof)))
            per = ms3 / max(1, n_prof.value) * 1e3
            r = K.spmv_host(xs[mode]) - b
            print(json.dumps({"workload": name, "N": N, "mode": "fused" if mode else "classic", "status": st,
                              "iterations": it, "ms": round(best, 3), "us_per_iter": round(best * 1e3 / it, 3),
                              "kernel_us": [round(float(v), 2) for v in per],
                              "true_rel_residual": float(np.linalg.norm(r) / np.linalg.norm(b))}), flush=True)
        d = float(np.linalg.norm(xs[0] - xs[1]) / np.linalg.norm(xs[0]))
        print(json.dumps({"workload": name, "rel_l2_fused_vs_classic": d}), flush=True)
        capi.check(L.nbgpu_set_pcg_mode(-1))
        K.destroy(); d_F.free()


if __name__ == "__main__":
    main()
