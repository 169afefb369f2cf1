"""torchrun worker: microseconds per iteration of the row-partitioned Jacobi-PCG on an NX x (NY_PER_RANK * world)
quad cantilever, fixed iteration budget.  For A/B runs of the exchange switches (NBGPU_DIST_PUSH_WARP,
NBGPU_DIST_HALO_LAST, NBGPU_PCG_MODE).   torchrun ... scripts/dist_size.py NX NY_PER_RANK [iters]"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist
from nbots_b200 import api, capi, multigpu
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29545")
dist.init_process_group("gloo", rank=rank, world_size=world)
L = capi.lib(); capi.check(L.nbgpu_init(int(os.environ.get("LOCAL_RANK", "0"))))
nx, nyr = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 500
def gather(o):
    out = [None] * world; dist.all_gather_object(out, o); return out
m, fem, bcs = multigpu._cantilever(nx, nyr * world, 2.0, 2.0 * nyr * world / nx, rank, world, gather)
best = 1e30
for rep in range(3):
    capi.check(L.nbgpu_memset(fem.d_x.ptr, 0, fem.N_loc * 8)); api.sync(); dist.barrier()
    it = C.c_uint32(0); res = C.c_double(0)
    api.timer_start()
    L.nbgpu_dist_pcg_jacobi(fem.dist, fem.plan, fem.A.h, fem.d_b.ptr, fem.d_x.ptr, iters, 0.0, C.byref(it), C.byref(res))
    best = min(best, api.timer_stop())
us = gather(round(best * 1e3 / iters, 2))
if rank == 0:
    print(json.dumps({"nx": nx, "ny_per_rank": nyr, "world": world, "rows_per_rank": fem.N_loc, "us_per_iter_max": max(us),
                      "switches": {k: v for k, v in os.environ.items() if k.startswith("NBGPU_")}}))
fem.close(); dist.barrier(); dist.destroy_process_group()
