"""torchrun worker for the BASELINE.json configs that are not the bench headline:

  q16  : configs[3]  16M-DOF cantilever (4000x2000 quads, "plane strain" flag), assembly + Jacobi-PCG to
         1e-8*|b|, row slabs over the ranks
  l64  : configs[2]  9-point Laplacian on an n x n grid (default 8192: 67M rows), distributed SpMV (50 reps),
         200 PCG iterations, and PCG to 1e-8*|b|
Rank 0 prints one JSON line per measurement."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from nbots_b200 import api, capi, meshgen, multigpu

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dist.init_process_group("gloo")
L = capi.lib(); capi.check(L.nbgpu_init(int(os.environ.get("LOCAL_RANK", "0")) % max(1, torch.cuda.device_count())))
PEAK = 6554.9


def gather(obj):
    out = [None] * world; dist.all_gather_object(out, obj); return out


def tmax(ms):
    t = torch.tensor([ms], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); return t.item()


def gsum(v):
    t = torch.tensor([v], dtype=torch.float64); dist.all_reduce(t); return t.item()


def emit(**kw):
    if rank == 0:
        print(json.dumps(kw), flush=True)


def laplacian_rows(n, r0, r1, chunk=1 << 22):
    rs = np.empty(r1 - r0, dtype=np.uint32)
    cols, vals = [], []
    for a in range(r0, r1, chunk):
        b = min(r1, a + chunk)
        s, c, v = meshgen.laplacian9_csr(n, a, b)
        rs[a - r0:b - r0] = s; cols.append(c); vals.append(v)
    return rs, np.concatenate(cols), np.concatenate(vals)


mode = sys.argv[1]
if mode == "q16":
    nx, ny = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4000, 2000)
    t0 = time.perf_counter()
    prob = multigpu.SlabProblem(nx, ny, 2.0, 2.0 * ny / nx, rank, world)
    t_asm = time.perf_counter() - t0
    dc = multigpu.DistContext(rank, world, prob.row_starts, prob.rows_size, prob.cols_global, prob.vals, gather)
    nb = np.sqrt(gsum(float(prob.b @ prob.b))); tol = 1e-8 * nb
    nnz = int(gsum(prob.nnz))
    d_b = api.DeviceBuffer.from_host(prob.b); d_x = api.DeviceBuffer.zeros(prob.N_loc)
    for rep in range(2):
        capi.check(L.nbgpu_memset(d_x.ptr, 0, prob.N_loc * 8)); api.sync(); dist.barrier()
        api.timer_start(); st, it, res = dc.pcg_jacobi(d_b, d_x, prob.N_global, tol); ms = tmax(api.timer_stop())
    x = d_x.to_host()
    umax = max(gather(float(np.abs(x).max())))
    emit(config="configs[3] 16M-DOF cantilever, 'plane strain' flag (reference: plane-stress D), row slabs",
         n_gpus=world, mesh=f"{nx}x{ny} quads", N_dof=prob.N_global, nnz=nnz, status=st, iterations=it,
         residual=res, tol=tol, solve_ms=round(ms, 2), us_per_iter=round(ms * 1e3 / it, 2),
         dof_iter_per_s=prob.N_global * it / (ms * 1e-3),
         alg_GBps_per_gpu=round((12 * nnz + 108 * prob.N_global) * it / (ms * 1e-3) / 1e9 / world, 1),
         frac_of_measured_peak=round((12 * nnz + 108 * prob.N_global) * it / (ms * 1e-3) / 1e9 / world / PEAK, 3),
         slab_setup_s=round(max(gather(t_asm)), 2), max_abs_u=umax)
    dc.close()
elif mode == "l64":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    lines = multigpu.slab_lines(n, world)
    row_starts = np.array([n * j for j in lines], dtype=np.uint32)
    r0, r1 = int(row_starts[rank]), int(row_starts[rank + 1])
    t0 = time.perf_counter()
    rs, cols, vals = laplacian_rows(n, r0, r1)
    t_gen = time.perf_counter() - t0
    dc = multigpu.DistContext(rank, world, row_starts, rs, cols, vals, gather)
    N, nnz = n * n, int(gsum(cols.size))
    b = meshgen.uniform_rhs(r1 - r0, seed=12345, start=r0)
    d_b = api.DeviceBuffer.from_host(b); d_y = api.DeviceBuffer.zeros(r1 - r0); d_x = api.DeviceBuffer.zeros(r1 - r0)
    class Raw:          # the window's own input vector: no copy into the window per SpMV
        ptr = L.nbgpu_dist_input_vector(dc.dist, dc.plan)
    capi.check(L.nbgpu_copy_h2d(Raw.ptr, b.ctypes.data, b.nbytes))
    for _ in range(5):
        dc.spmv(Raw, d_y)
    api.sync(); dist.barrier(); api.timer_start()
    reps = 50
    for _ in range(reps):
        dc.spmv(Raw, d_y)
    ms = tmax(api.timer_stop()) / reps
    bytes_spmv = 12 * nnz + 20 * N + 4
    emit(config="configs[2] 9-pt Laplacian SpMV", n_gpus=world, n=n, rows=N, nnz=nnz, spmv_ms=round(ms, 4),
         alg_GBps_total=round(bytes_spmv / (ms * 1e-3) / 1e9, 1),
         alg_GBps_per_gpu=round(bytes_spmv / (ms * 1e-3) / 1e9 / world, 1),
         frac_of_measured_peak_per_gpu=round(bytes_spmv / (ms * 1e-3) / 1e9 / world / PEAK, 3), gen_s=round(t_gen, 1))
    nb = np.sqrt(gsum(float(b @ b)))
    for label, max_iter, tol in (("200 iterations", 200, 0.0), ("to 1e-8*|b|", N, 1e-8 * nb)):
        capi.check(L.nbgpu_memset(d_x.ptr, 0, (r1 - r0) * 8)); api.sync(); dist.barrier()
        api.timer_start(); st, it, res = dc.pcg_jacobi(d_b, d_x, max_iter, tol); ms = tmax(api.timer_stop())
        emit(config=f"configs[2] 9-pt Laplacian Jacobi-PCG, {label}", n_gpus=world, rows=N, status=st, iterations=it,
             residual=res, solve_ms=round(ms, 2), us_per_iter=round(ms * 1e3 / it, 2),
             rows_iter_per_s=N * it / (ms * 1e-3),
             alg_GBps_per_gpu=round((12 * nnz + 108 * N) * it / (ms * 1e-3) / 1e9 / world, 1))
    dc.close()
dist.destroy_process_group()
