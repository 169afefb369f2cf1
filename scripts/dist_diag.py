"""torchrun worker: per-rank speed of the local (communication-free) SpMV while all ranks are busy,
next to the distributed iteration time -- separates imbalance/jitter from exchange latency."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from nbots_b200 import api, capi, multigpu

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
L = capi.lib(); capi.check(L.nbgpu_init(int(os.environ.get("LOCAL_RANK", "0")) % max(1, torch.cuda.device_count())))
def gather(obj):
    out = [None] * world; dist.all_gather_object(out, obj); return out
nx, ny_per = int(sys.argv[1]), int(sys.argv[2])
prob = multigpu.SlabProblem(nx, ny_per * world, 2.0, 2.0 * ny_per * world / nx, rank, world)
dc = multigpu.DistContext(rank, world, prob.row_starts, prob.rows_size, prob.cols_global, prob.vals, gather)
n_ext = dc.N_loc + dc.n_halo
d_in = api.DeviceBuffer.from_host(np.ones(n_ext)); d_out = api.DeviceBuffer.zeros(dc.N_loc)
for _ in range(10):
    dc.A.spmv(d_in, d_out)
res = []
for rep in range(3):
    api.sync(); dist.barrier(); api.timer_start()
    for _ in range(200):
        dc.A.spmv(d_in, d_out)
    res.append(api.timer_stop() / 200 * 1e3)
local = gather(min(res))
d_b = api.DeviceBuffer.from_host(prob.b); d_x = api.DeviceBuffer.zeros(dc.N_loc)
times = []
for rep in range(3):
    capi.check(L.nbgpu_memset(d_x.ptr, 0, dc.N_loc * 8)); api.sync(); dist.barrier()
    api.timer_start(); st, it, r = dc.pcg_jacobi(d_b, d_x, 600, 0.0); times.append(api.timer_stop() / it * 1e3)
pcg = gather(min(times))
if rank == 0:
    print(f"world={world} N_loc={dc.N_loc} local SpMV us per rank: {[round(v,1) for v in local]}  spread {(max(local)/min(local)-1)*100:.1f}%")
    print(f"   dist PCG us/iter per rank: {[round(v,1) for v in pcg]}")
dc.close(); dist.destroy_process_group()
