"""torchrun worker (diagnostic build: make -C nbots_b200/csrc LIBDIR=../lib_tl EXTRA=-DNB_TIMELINE; NBGPU_LIB_DIR=.../lib_tl):
where an iteration of the row-partitioned CLASSIC solver spends its time, from per-GPU %globaltimer stamps.
Prints per-rank medians over iterations 50..450 of the intervals between the phase stamps (ns)."""
import ctypes as C, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.distributed as dist
from nbots_b200 import api, capi, multigpu
import bench as B
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29544")
dist.init_process_group("gloo", rank=rank, world_size=world)
L = capi.lib(); capi.check(L.nbgpu_init(int(os.environ.get("LOCAL_RANK", "0"))))
def gather(o):
    out = [None] * world; dist.all_gather_object(out, o); return out
m, fem, bcs = multigpu._cantilever(B.NX, B.NY_PER_GPU * world, 2.0, 1.0 * world, rank, world, gather)
for rep in range(2):
    capi.check(L.nbgpu_memset(fem.d_x.ptr, 0, fem.N_loc * 8)); api.sync(); dist.barrier()
    it = C.c_uint32(0); res = C.c_double(0)
    api.timer_start()
    L.nbgpu_dist_pcg_jacobi(fem.dist, fem.plan, fem.A.h, fem.d_b.ptr, fem.d_x.ptr, 500, 0.0, C.byref(it), C.byref(res))
    ms = api.timer_stop()
tl = np.zeros(512 * 10, dtype=np.uint64)
L.nbgpu_dist_timeline.argtypes = [C.c_void_p]
capi.check(L.nbgpu_dist_timeline(tl.ctypes.data))
t = tl.reshape(512, 10).astype(np.int64)[50:450]
nxt = tl.reshape(512, 10).astype(np.int64)[51:451]
med = lambda a: float(np.median(a))
out = {"rank": rank, "us_per_iter": round(ms * 1e3 / 500, 2),
       "K1 start -> CTA0 rows done": med(t[:, 9] - t[:, 0]),
       "K1 CTA0 rows done -> K2 CTA0 past wait": med(t[:, 3] - t[:, 9]),
       "K2 CTA0 past wait -> K3 CTA0 past wait": med(t[:, 6] - t[:, 3]),
       "K1 start -> last CTA reduced (posts)": med(t[:, 1] - t[:, 0]),
       "K1 start -> halo arrived (CTA0 late wait)": med(t[:, 2] - t[:, 0]) if t[:, 2].any() else None,
       "K1 posted -> K2 CTA0 past wait": med(t[:, 3] - t[:, 1]),
       "K2 collect (waiting for the ranks' p.w)": med(t[:, 4] - t[:, 3]),
       "K2 collected -> last CTA reduced (posts)": med(t[:, 5] - t[:, 4]),
       "K2 posted -> K3 CTA0 past wait": med(t[:, 6] - t[:, 5]),
       "K3 collect (waiting for g.g, g.q)": med(t[:, 7] - t[:, 6]),
       "K3 collected -> CTA0 done": med(t[:, 8] - t[:, 7]),
       "K3 CTA0 done -> next K1 start": med(nxt[:, 0] - t[:, 8]),
       "iteration (K1 start to K1 start)": med(nxt[:, 0] - t[:, 0])}
for o in gather(out):
    if rank == 0:
        print(json.dumps(o))
fem.close(); dist.barrier(); dist.destroy_process_group()
