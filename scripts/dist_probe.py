"""torchrun worker: time the distributed PCG on the weak-scaling slab problem (fixed iteration count)."""
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
ny_per = int(sys.argv[1]) if len(sys.argv) > 1 else 500
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 600
prob = multigpu.SlabProblem(1000, ny_per * world, 2.0, 1.0 * world, rank, world)
dc = multigpu.DistContext(rank, world, prob.row_starts, prob.rows_size, prob.cols_global, prob.vals, gather)
d_b = api.DeviceBuffer.from_host(prob.b); d_x = api.DeviceBuffer.zeros(prob.N_loc)
for rep in range(3):
    capi.check(L.nbgpu_memset(d_x.ptr, 0, prob.N_loc * 8)); api.sync(); dist.barrier()
    api.timer_start()
    st, it, res = dc.pcg_jacobi(d_b, d_x, iters, 0.0)
    ms = api.timer_stop()
    t = torch.tensor([ms], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    env = {k: v for k, v in os.environ.items() if k.startswith("NBGPU_")}
    print(f"world={world} N={prob.N_global} iters={it} {t.item()/it*1e3:.1f} us/iter  {prob.N_global*it/t.item()/1e6:.2f} GDOFit/s env={env}")
dc.close(); dist.destroy_process_group()
