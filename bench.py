#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200: Jacobi-PCG DOF*iter/s (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W]        this implementation (CUDA, sm_100a)
  python bench.py --impl reference ...                       the reference's own CPU path, host cores

Workload at N=1: BASELINE.json configs[1] -- synthetic structured-quad cantilever, 1000x500 quads,
1 003 002 DOF, plane stress, assembled ON THE DEVICE, Jacobi-PCG from x0 = 0 to |g| <= 1e-8 |b|
(SURVEY.md §8d input 2).  For N>1 the mesh grows with N (1000 x 500N quads, slabs of grid lines per
rank: weak scaling, ~1M DOF per GPU).

One "step" = one complete solve.  `value` = DOF * iterations / device time with everything resident in
HBM; `e2e` = the same solve through the reference-facing entry point
nb_sparse_solve_CG_precond_Jacobi (libnbots_b200.so) with a host-resident nb_sparse_t in the reference's own
memory layout (two heap blocks per row, sparse_struct.c:23-31) and host vectors, all copies inside the timed
region.  `roofline` is the SpMV+dot kernel of the solver (the dominant kernel) timed live with CUDA events on
the library's stream.

Beside the headline the line carries
  `target`  the north-star configuration: Q16 (4000x2000 quads, 16 012 002 DOF, "plane strain" flag) split into
            N slabs, a fixed budget of 1000 iterations: us per iteration, K1 CUDA-event time, algorithmic and
            physically moved GB/s per GPU, strong-scaling efficiency against one GPU measured in the same run;
            at N = 1 also the L64 SpMV (8192^2 9-point Laplacian, 50 repetitions);
  `parity`  (N > 1) rank-local rows against the single-GPU matrix (checksums), 50 iterations against the
            single-GPU path and the CPU reference, full-solve iteration count against the single-GPU count;
            the run fails (exit code 1) when one of them does not hold.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from nbots_b200 import meshgen  # noqa: E402

METRIC = "pcg_dof_iter_per_s"
UNIT = "DOF*iter/s"
NX, NY_PER_GPU = 1000, 500
E_MOD, POISSON, THICKNESS = 1.0, 0.3, 1.0
REL_TOL = 1e-8
Q16 = (4000, 2000)
TARGET_ITERS = 1000
L64_N = 8192


def workload_mesh(n_gpus):
    nx, ny = NX, NY_PER_GPU * n_gpus
    return meshgen.structured_mesh(nx, ny, 2.0, 1.0 * n_gpus, kind=1)


def workload_name(n_gpus):
    if n_gpus == 1:
        return ("Q1: structured-quad cantilever 1000x500, plane stress, Jacobi-PCG to 1e-8*|b| "
                "(BASELINE.json configs[1])")
    return (f"Q1 x {n_gpus}: structured-quad cantilever {NX}x{NY_PER_GPU * n_gpus} in {n_gpus} slabs of grid lines, "
            "plane stress, Jacobi-PCG to 1e-8*|b| (BASELINE.json configs[1] per GPU)")


def workload_bcs():
    # clamp side x=0 (segment 3), total traction (0,-1) on side x=L (segment 1)
    return [("dirichlet", "sgm", 3, (1, 1), (0.0, 0.0)), ("neumann", "sgm", 1, (1, 1), (0.0, -1.0))]


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 6 and r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# host-side nb_sparse_t (struct nb_sparse_s, sparse_struct.h:6-11) for the e2e call, in the memory layout
# nb_sparse_allocate gives it (sparse_struct.c:23-31): two heap blocks per row.
class NbSparse(C.Structure):
    _fields_ = [("rows_values", C.POINTER(C.c_void_p)), ("rows_index", C.POINTER(C.c_void_p)),
                ("rows_size", C.POINTER(C.c_uint32)), ("N", C.c_uint32)]


def host_nb_sparse(shim, rows_size, cols, vals):
    shim.nbshim_sparse_from_csr.restype = C.POINTER(NbSparse)
    shim.nbshim_sparse_from_csr.argtypes = [C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                            C.POINTER(C.c_double)]
    shim.nbshim_sparse_free.argtypes = [C.POINTER(NbSparse)]
    A = shim.nbshim_sparse_from_csr(rows_size.size, rows_size.ctypes.data_as(C.POINTER(C.c_uint32)),
                                    cols.ctypes.data_as(C.POINTER(C.c_uint32)),
                                    vals.ctypes.data_as(C.POINTER(C.c_double)))
    assert A, "out of host memory"
    return A


def kernel_profile(L, solve256):
    """CUDA events around each solver kernel of 256 iterations (nbgpu_krylov_profile) -> us per launch [K1, K2, K3]."""
    from nbots_b200 import capi
    capi.check(L.nbgpu_krylov_profile(1))
    solve256()
    capi.check(L.nbgpu_krylov_profile(0))
    ms3 = np.zeros(3); n_prof = C.c_uint32(0)
    capi.check(L.nbgpu_krylov_profile_get(ms3.ctypes.data_as(capi.f64p), C.byref(n_prof)))
    return ms3 / max(1, n_prof.value) * 1e3, int(n_prof.value)


def matrix_bytes_moved(K):
    """HBM bytes one pass over the stored matrix moves: 8-byte values + the column-id stream of the layout in use."""
    per_entry_ids = (0.5 if K.idx16 else 1.0) if K.blocked else (2.0 if K.idx16 else 4.0)
    return int(K.stored * (8.0 + per_entry_ids))


def target_single_gpu(L, peak):
    """`target` at N = 1: Q16 on one GPU (fixed iteration budget) and the L64 SpMV."""
    from nbots_b200 import api, capi
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import flatten_bcs
    out = {}
    t0 = time.perf_counter()
    nx, ny = Q16
    m = meshgen.structured_mesh(nx, ny, 2.0, 1.0, kind=1)
    mesh = api.Mesh(m)
    api.sync()
    # pattern + matrix on the device.  The first build of this size also grows the library's
    # allocation pool by ~4 GB (cudaMalloc); the second shows the builder itself.
    t_pat_first = time.perf_counter()
    K = mesh.create_matrix()
    api.sync()
    t_pat_first = time.perf_counter() - t_pat_first
    assert K is not None
    K.destroy()
    t_pat = time.perf_counter()
    K = mesh.create_matrix()
    api.sync()
    t_pat = time.perf_counter() - t_pat
    assert K is not None
    d_F = api.DeviceBuffer.zeros(K.N)
    api.sync(); api.timer_start()
    st, _ = mesh.assemble(K, d_F, E_MOD, POISSON, analysis=1, thickness=THICKNESS)   # "plane strain" flag (reference: plane-stress D)
    ms_asm = api.timer_stop()
    assert st == 0
    neu_dof, neu_add, dir_dof, dir_val = flatten_bcs(m, workload_bcs())
    api.vector_add_entries(d_F, neu_dof, neu_add)
    K.apply_dirichlet(d_F, dir_dof, dir_val)
    N, nnz = K.N, K.nnz
    d_x = api.DeviceBuffer.zeros(N)
    setup_s = time.perf_counter() - t0
    best = 1e30
    for rep in range(2):
        capi.check(L.nbgpu_memset(d_x.ptr, 0, N * 8))
        api.timer_start()
        st, it, res = K.pcg_jacobi(d_F, d_x, max_iter=TARGET_ITERS, tol=0.0)
        best = min(best, api.timer_stop())
    assert it == TARGET_ITERS

    def solve256():
        capi.check(L.nbgpu_memset(d_x.ptr, 0, N * 8))
        K.pcg_jacobi(d_F, d_x, max_iter=256, tol=0.0)
    per, n_prof = kernel_profile(L, solve256)
    us_it = best * 1e3 / it
    alg_k1 = 12 * nnz + 20 * N + 4
    phys_k1 = matrix_bytes_moved(K) + 16 * N
    out["q16"] = {"workload": "Q16: structured-quad cantilever 4000x2000, 'plane strain' flag (reference semantics: "
                              "plane-stress D), assembled on the device, Jacobi-PCG, fixed budget from x0 = 0 "
                              "(BASELINE.json configs[3] on one GPU)",
                  "N_dof": int(N), "nnz": int(nnz), "iterations": int(it), "us_per_iteration": round(us_it, 2),
                  "dof_iter_per_s": N * it / (best * 1e-3),
                  "kernel_us": [round(float(v), 2) for v in per],
                  "k1_GBps_algorithmic": round(alg_k1 / (per[0] * 1e-6) / 1e9, 1),
                  "k1_GBps_physical": round(phys_k1 / (per[0] * 1e-6) / 1e9, 1),
                  "k1_frac_of_peak_physical": round(phys_k1 / (per[0] * 1e-6) / 1e9 / peak, 4),
                  "iteration_GBps_algorithmic": round((12 * nnz + 108 * N) / (us_it * 1e-6) / 1e9, 1),
                  "layout": {"blocked": K.blocked, "idx16": K.idx16, "uniform_width": K.uniform_width},
                  "assembly_ms_on_device": round(ms_asm, 3), "pattern_and_matrix_on_device_ms": round(t_pat * 1e3, 2),
                  "pattern_and_matrix_first_call_ms": round(t_pat_first * 1e3, 2), "setup_s": round(setup_s, 2)}
    K.destroy(); mesh.destroy(); d_F.free(); d_x.free()
    del m

    # ---- L64: 8192^2 9-point Laplacian, SpMV timed over 50 repetitions
    t0 = time.perf_counter()
    n = L64_N
    rs = np.empty(n * n, dtype=np.uint32)
    cols_l, vals_l = [], []
    chunk = 1 << 23
    for a in range(0, n * n, chunk):
        b = min(n * n, a + chunk)
        s_, c_, v_ = meshgen.laplacian9_csr(n, a, b)
        rs[a:b] = s_; cols_l.append(c_); vals_l.append(v_)
    cols = np.concatenate(cols_l); del cols_l
    vals = np.concatenate(vals_l); del vals_l
    A = api.Matrix.from_csr(rs, cols, vals)
    del cols, vals
    N, nnz = A.N, A.nnz
    x = meshgen.uniform_rhs(N, seed=12345)
    d_in = api.DeviceBuffer.from_host(x); d_out = api.DeviceBuffer.zeros(N)
    gen_s = time.perf_counter() - t0
    for _ in range(5):
        A.spmv(d_in, d_out)
    api.sync(); api.timer_start()
    for _ in range(50):
        A.spmv(d_in, d_out)
    ms = api.timer_stop() / 50
    alg = 12 * nnz + 20 * N + 4
    phys = matrix_bytes_moved(A) + 16 * N
    # parity at full size: a checksum of the product against the closed form of the stencil
    # (row sums of A are 0 in the interior, so A * ones is nonzero on the boundary only)
    ones = api.DeviceBuffer.from_host(np.ones(N)); A.spmv(ones, d_out)
    y = d_out.to_host().reshape(n, n)
    want = np.zeros((n, n)); want[0, :] += 3; want[-1, :] += 3; want[:, 0] += 3; want[:, -1] += 3
    want[0, 0] -= 1; want[0, -1] -= 1; want[-1, 0] -= 1; want[-1, -1] -= 1
    out["l64_spmv"] = {"workload": f"L64: {n}^2-grid 9-point Laplacian (BASELINE.json configs[2]), nbgpu_spmv, 50 repetitions",
                       "rows": int(N), "nnz": int(nnz), "ms": round(ms, 4),
                       "GBps_algorithmic": round(alg / (ms * 1e-3) / 1e9, 1),
                       "frac_of_peak_algorithmic": round(alg / (ms * 1e-3) / 1e9 / peak, 4),
                       "GBps_physical": round(phys / (ms * 1e-3) / 1e9, 1),
                       "frac_of_peak_physical": round(phys / (ms * 1e-3) / 1e9 / peak, 4),
                       "layout": {"blocked": A.blocked, "idx16": A.idx16, "uniform_width": A.uniform_width},
                       "A_times_ones_matches_closed_form": bool(np.array_equal(y, want)), "gen_s": round(gen_s, 1)}
    A.destroy(); d_in.free(); d_out.free(); ones.free()
    return out


# ------------------------------------------------------------------------------------------------
def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # torch.distributed is the rendezvous only (exchange of halo lists and CUDA IPC handles, barriers,
        # max-over-ranks of the timings): gloo is enough -- the solver's own traffic goes over NVLink
        # as peer stores issued by its kernels, not through a collective library.
        import torch.distributed as dist
        dist.init_process_group("gloo")
        return rank, world, dist
    if os.environ.get("NBGPU_BENCH_FORCE_DIST"):
        # diagnostic: the row-partitioned code path with a world of one rank (its fixed overhead)
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)
        return 0, 1, dist
    return rank, world, None


def run_ours(args):
    from nbots_b200 import api, capi
    rank, world, dist = dist_setup(args.gpus)
    if dist is not None:
        from nbots_b200 import multigpu
        return multigpu.bench(args, rank, world, dist)

    L = capi.lib()
    capi.check(L.nbgpu_init(int(os.environ.get("LOCAL_RANK", "0"))))
    t_setup = time.perf_counter()
    m = workload_mesh(1)
    mesh = api.Mesh(m)
    t_pat = time.perf_counter()
    K = mesh.create_matrix()                   # pattern built on the device straight into the SELL arrays
    api.sync()
    t_pat = time.perf_counter() - t_pat
    assert K is not None
    rs, cols = K.pattern_csr()                 # host copy for the e2e call's nb_sparse_t (not timed anywhere)
    d_F = api.DeviceBuffer.zeros(K.N)
    api.sync()
    api.timer_start()
    st, _ = mesh.assemble(K, d_F, E_MOD, POISSON, thickness=THICKNESS)
    ms_assembly = api.timer_stop()
    assert st == 0
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import flatten_bcs
    neu_dof, neu_add, dir_dof, dir_val = flatten_bcs(m, workload_bcs())
    api.vector_add_entries(d_F, neu_dof, neu_add)
    K.apply_dirichlet(d_F, dir_dof, dir_val)
    b = d_F.to_host()
    tol = REL_TOL * float(np.linalg.norm(b))
    N, nnz = K.N, K.nnz
    d_x = api.DeviceBuffer.zeros(N)
    t_setup = time.perf_counter() - t_setup

    def solve_resident():
        capi.check(L.nbgpu_memset(d_x.ptr, 0, N * 8))
        api.timer_start()
        st, it, res = K.pcg_jacobi(d_F, d_x, max_iter=N, tol=tol)
        return api.timer_stop(), st, it, res

    for _ in range(args.warmup):
        solve_resident()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    launches0 = api.launch_count()
    times, iters = [], 0
    for _ in range(args.steps):
        ms, st, it, res = solve_resident()
        times.append(ms)
        iters = it
        assert st == 0, "solve did not converge"
    launches = api.launch_count() - launches0
    clocks = sampler.stop()
    ms_step = float(np.mean(times))
    value = N * iters / (ms_step * 1e-3)
    x_resident = d_x.to_host()

    # ---- dominant kernel, live: CUDA events around each kernel of 256 iterations ------------------
    def solve256():
        capi.check(L.nbgpu_memset(d_x.ptr, 0, N * 8))
        K.pcg_jacobi(d_F, d_x, max_iter=256, tol=0.0)
    per_us, n_prof_v = kernel_profile(L, solve256)
    per = per_us * 1e-3

    class n_prof:
        value = n_prof_v
    peak, peak_src = measured_peaks()
    bytes_spmv = 12 * nnz + 20 * N + 4          # SURVEY.md §8d, K1 (CSR-equivalent algorithmic bytes)
    bytes_update, bytes_dir = 40 * N, 40 * N     # K2 (g, w, diag -> g, q), K3 (q, p, x -> p, x)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    traffic_src = None
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj.get("krylov_spmv_kernel_dram_bytes_per_launch")
        traffic_src = {"file": "profiles/traffic.json", "commit": tj.get("commit"), "capture": tj.get("capture")}
    ach = bytes_spmv / (per[0] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "krylov_spmv_stream_kernel<blocked, 16-bit ids> (SpMV + p.w, TMA-staged SELL-32)", "achieved": round(ach, 1),
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": bytes_spmv,
                "physical_bytes_model": matrix_bytes_moved(K) + 16 * N,
                "achieved_physical_GBps": round((traffic or (matrix_bytes_moved(K) + 16 * N)) / (per[0] * 1e-3) / 1e9, 1),
                "ms_per_launch": round(float(per[0]), 5), "launches_timed": int(n_prof.value),
                "other_kernels": {
                    "krylov_update_kernel": {"ms": round(float(per[1]), 5),
                                             "GB/s": round(bytes_update / (per[1] * 1e-3) / 1e9, 1)},
                    "krylov_dir_kernel": {"ms": round(float(per[2]), 5),
                                          "GB/s": round(bytes_dir / (per[2] * 1e-3) / 1e9, 1)}},
                "iteration_bytes_model": 12 * nnz + 108 * N,   # SURVEY §8d model (the kernels move 100 N of vectors)
                "iteration_GBps": round((12 * nnz + 108 * N) * iters / (ms_step * 1e-3) / 1e9, 1)}

    # ---- e2e: the reference-facing call with host buffers --------------------------------------------
    vals = K.values_csr()
    shim = C.CDLL(capi.SHIM_PATH)
    A_host = host_nb_sparse(shim, rs, cols, vals)
    fn = shim.nb_sparse_solve_CG_precond_Jacobi
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(NbSparse), capi.f64p, capi.f64p, C.c_uint32, C.c_double, capi.u32p, capi.f64p,
                   C.c_uint32]
    x_host = np.zeros(N)
    e2e_times, e2e_iters = [], 0
    for k in range(2 + args.steps):
        x_host[:] = 0.0
        it = C.c_uint32(0); res = C.c_double(0)
        api.sync()
        t0 = time.perf_counter()
        st = fn(A_host, b.ctypes.data_as(capi.f64p), x_host.ctypes.data_as(capi.f64p), N, tol,
                C.byref(it), C.byref(res), 1)
        dt = time.perf_counter() - t0
        assert st == 0, capi.lib().nbgpu_last_error()
        if os.environ.get("BENCH_DEBUG"):
            print(f"e2e call {k}: {dt*1e3:.1f} ms, iters {it.value}", file=sys.stderr)
        if k >= 2:
            e2e_times.append(dt)
        e2e_iters = it.value
    e2e_value = N * e2e_iters / float(np.median(e2e_times))
    assert np.array_equal(x_host, x_resident), "e2e and resident solves must be the same computation"
    t3 = (C.c_double * 3)()
    shim.nbshim_last_timings(t3)
    shim.nbshim_sparse_free(A_host)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(12 * nnz + 16 * N),
           "d2h_bytes_per_step": int(8 * N), "ms_per_step": round(float(np.median(e2e_times)) * 1e3, 2),
           "import_ms": round(t3[0], 2), "solve_ms_incl_vector_copies": round(t3[1], 2), "release_ms": round(t3[2], 2),
           "entry_point": "nb_sparse_solve_CG_precond_Jacobi (libnbots_b200.so) on a host nb_sparse_t with the "
                          "reference's layout: two heap blocks per row (sparse_struct.c:23-31)"}

    cpu = cpu_baseline(m, b, tol, x_resident, vals, d_F_host=b) if not args.no_cpu_baseline else None
    K.destroy(); mesh.destroy(); d_F.free(); d_x.free()
    del vals, cols
    target = None
    if not args.no_target:
        target = target_single_gpu(L, peak)
        target["n_gpus"] = 1

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_step, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(1),
                       "N_dof": N, "nnz": int(nnz), "iterations_per_step": int(iters), "rel_tol": REL_TOL,
                       "l2": "working set 265 MB (matrix 216 MB + 6 vectors) exceeds the 126 MB L2; no flush",
                       "assembly_ms_on_device": round(ms_assembly, 3),
                       "pattern_and_matrix_on_device_ms": round(t_pat * 1e3, 2), "setup_s": round(t_setup, 2)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "target": target,
            "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def reference_system(m):
    """The workload's linear system built by the REFERENCE'S OWN CPU pipeline (oracle/_ref), or by the
    pinned port when the compiled reference is not available.  -> (kind, solver object, b)"""
    from oracle import port, ref
    bcs = workload_bcs()
    if ref.available():
        rm = ref.RefMesh.from_arrays(m)
        K = ref.RefSparse.from_mesh(rm)
        st, F = ref.assemble(K, rm, 1, E_MOD, POISSON, thickness=THICKNESS)
        bc = ref.RefBcond()
        for r in bcs:
            bc.push(*r)
        ref.set_bconditions(rm, K, F, bc)
        return "reference", K, F
    rs, cols = port.pattern_from_mesh(m)
    K = port.Csr(rs, cols)
    st, F = port.assemble(K, m, E_MOD, POISSON, thickness=THICKNESS)
    port.set_bconditions(m, K, F, bcs)
    return "port", K, F


def time_cpu_pcg(K, b, threads, budget_s):
    """Bounded sample: a fixed number of PCG iterations from x0 = 0 sized for ~budget_s seconds."""
    t0 = time.perf_counter()
    K.pcg_jacobi(b, max_iter=10, tol=0.0, threads=threads)
    per_iter = (time.perf_counter() - t0) / 10
    n_it = int(min(2000, max(20, budget_s / per_iter)))
    t0 = time.perf_counter()
    st, x, it, res = K.pcg_jacobi(b, max_iter=n_it, tol=0.0, threads=threads)
    dt = time.perf_counter() - t0
    return it, dt


def cpu_baseline(m, b_gpu, tol, x_gpu, vals_gpu, d_F_host):
    cores = os.cpu_count() or 1
    kind, K, F = reference_system(m)
    vals_ref = K.export()[2] if kind == "reference" else K.vals
    it, dt = time_cpu_pcg(K, F, cores, 12.0)
    N = F.size
    return {"value": N * it / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{it} Jacobi-PCG iterations of the same Q1 system from x0=0 "
                      f"({dt:.1f} s, omp_parallel_threads={cores})",
            "fullsize_parity": {"K_bit_exact": bool(np.array_equal(vals_ref, vals_gpu)),
                                "F_bit_exact": bool(np.array_equal(F, b_gpu))}}


def run_reference(args):
    """The reference's own CPU implementation of the path, all host threads; rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    m = workload_mesh(args.gpus)
    kind, K, F = reference_system(m)
    N = F.size
    tot_it, tot_t = 0, 0.0
    per_step_budget = max(2.0, 40.0 / max(1, args.steps + args.warmup))
    for k in range(args.warmup + args.steps):
        it, dt = time_cpu_pcg(K, F, cores, per_step_budget)
        if k >= args.warmup:
            tot_it += it
            tot_t += dt
    value = N * tot_it / tot_t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(tot_t / args.steps * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.gpus), "N_dof": int(N),
                       "sample": "fixed iteration budget per step from x0 = 0 (a full CPU solve takes minutes)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{tot_it // args.steps} PCG iterations per step from x0=0, "
                                       f"nb_sparse_solve_CG_precond_Jacobi with omp_parallel_threads={cores}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-target", action="store_true", help="skip the Q16 / L64 `target` block")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the `parity` block")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
