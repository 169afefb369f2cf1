"""BASELINE.json configs[4]: SIMP-style loop -- repeated device assembly (per-element stiffness E_e = rho_e^3 E)
+ boundary conditions + warm-started Jacobi-PCG on a 2000x1000-quad cantilever (4M DOF), K and x never leave
the device.  The density update is a deterministic stand-in for an optimiser (no sensitivities are computed);
what is measured is the call pattern.  One JSON line per outer iteration, then a summary."""
import ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nbots_b200 import api, capi, meshgen
from util import flatten_bcs

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2000, 1000)
outer = int(sys.argv[3]) if len(sys.argv) > 3 else 10
L = capi.lib(); capi.check(L.nbgpu_init(0))
m = meshgen.structured_mesh(nx, ny, 2.0, 2.0 * ny / nx, kind=1)
recs = [("dirichlet", "sgm", 3, (1, 1), (0.0, 0.0)), ("neumann", "sgm", 1, (1, 1), (0.0, -1.0))]
neu_dof, neu_add, dir_dof, dir_val = flatten_bcs(m, recs)
D = api.constitutive_matrix(1.0, 0.3, 0)


class Desc(C.Structure):
    _fields_ = [("N_nod", C.c_uint32), ("nod", capi.f64p), ("N_elems", C.c_uint32), ("npe", C.c_uint32),
                ("adj", capi.u32p), ("N_edg", C.c_uint32), ("edg", capi.u32p), ("N_vtx", C.c_uint32),
                ("vtx", capi.u32p), ("N_sgm", C.c_uint32), ("sgm_sizes", capi.u32p), ("sgm_nodes", capi.u32p)]


class Report(C.Structure):
    _fields_ = [("N", C.c_uint32), ("nnz", C.c_uint64), ("iters", C.c_uint32), ("status", C.c_int32),
                ("residual", C.c_double), ("ms", C.c_double * 6)]


p = lambda a, t: a.ctypes.data_as(t)  # noqa: E731
d = Desc(m.n_nod, p(m.nod, capi.f64p), m.n_elems, m.npe, p(m.adj, capi.u32p), 0, None, m.vtx.size, p(m.vtx, capi.u32p),
         m.sgm_sizes.size, p(m.sgm_sizes, capi.u32p), p(m.sgm_nodes, capi.u32p))
S = C.c_void_p()
t0 = time.perf_counter()
capi.check(L.nbgpu_fem_session_create(C.byref(d), None, p(D, capi.f64p), 0.0, neu_dof.size, p(neu_dof, capi.u32p),
                                      p(neu_add, capi.f64p), dir_dof.size, p(dir_dof, capi.u32p), p(dir_val, capi.f64p),
                                      0, None, 1.0, capi.ASSEMBLY_GATHER, C.byref(S)))
t_create = time.perf_counter() - t0
rho = np.full(m.n_elems, 0.5)
cx = (np.arange(m.n_elems) % nx + 0.5) / nx
cy = (np.arange(m.n_elems) // nx + 0.5) / ny
tot = dict(asm=0.0, solve=0.0, iters=0)
bnorm = None
for k in range(outer):
    scale = rho ** 3
    rep = Report()
    tol = 1e-8 * bnorm if bnorm else 1e-12
    st = L.nbgpu_fem_session_step(S, None, p(scale, capi.f64p), 1, 0 if bnorm else 1, tol, C.byref(rep))
    if bnorm is None:      # first call only measured |b| (one iteration); now solve for real
        bnorm = float(np.sqrt(np.dot(neu_add, neu_add)))   # loads only on free dofs: |b| = |Neumann adds| here
        rep = Report()
        st = L.nbgpu_fem_session_step(S, None, p(scale, capi.f64p), 0, 0, 1e-8 * bnorm, C.byref(rep))
    assert st == 0, L.nbgpu_last_error()
    ms = list(rep.ms)
    tot["asm"] += ms[2]; tot["solve"] += ms[4]; tot["iters"] += rep.iters
    print(json.dumps({"outer": k, "assembly_ms": round(ms[2], 3), "bcond_ms": round(ms[3], 3), "solve_ms": round(ms[4], 2),
                      "pcg_iters": rep.iters, "status": rep.status,
                      "elem_per_s": m.n_elems / (ms[2] * 1e-3), "dof_iter_per_s": rep.N * rep.iters / (ms[4] * 1e-3)}))
    # stand-in density update: material drifts towards a diagonal band, volume roughly kept
    target = np.clip(1.2 - 3.0 * np.abs(cy - (1.0 - cx)), 0.05, 1.0)
    rho = np.clip(0.7 * rho + 0.3 * target, 0.05, 1.0)
print(json.dumps({"config": "configs[4] SIMP-style loop", "mesh": f"{nx}x{ny}", "N_dof": 2 * m.n_nod, "outer_iterations": outer,
                  "session_create_s": round(t_create, 2), "assembly_elem_per_s": m.n_elems * outer / (tot["asm"] * 1e-3),
                  "pcg_dof_iter_per_s": 2 * m.n_nod * tot["iters"] / (tot["solve"] * 1e-3), "total_pcg_iters": tot["iters"]}))
capi.check(L.nbgpu_fem_session_destroy(S))
