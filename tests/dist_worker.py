"""Worker of the multi-rank tests (launched once per rank).

  mode cpu : gloo, no GPU -- the product's partition plan + list exchange, with the local arithmetic done by the
             oracle (numpy / oracle.port) and the halo / reductions moved over gloo.
  mode gpu : one rank per GPU (or all ranks on GPU 0 when fewer are visible) -- nbgpu_dist_* end to end.
Rank 0 prints DIST_OK on success."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from nbots_b200 import multigpu  # noqa: E402
from oracle import port  # noqa: E402
from util import golden, rel_l2  # noqa: E402


def gather_obj_fn(world):
    def f(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out
    return f


def split_rows(g, world):
    """Row blocks of the golden quad cantilever cut between grid lines (64 x 16 elements -> 65 x 17 nodes)."""
    NX, NY = 65, 17
    lines = multigpu.slab_lines(NY, world)
    return np.array([2 * NX * j for j in lines], dtype=np.uint32)


def run_cpu(rank, world):
    g = golden("quad_cantilever_64x16")
    rs, cols, vals, b = g["rows_size"], g["cols"], g["K_post"], g["F_post"]
    rp = port.row_ptr_of(rs).astype(np.int64)
    row_starts = split_rows(g, world)
    r0, r1 = int(row_starts[rank]), int(row_starts[rank + 1])
    gather = gather_obj_fn(world)
    dc = multigpu.DistContext(rank, world, row_starts, rs[r0:r1], cols[rp[r0]:rp[r1]], None, gather)
    # plan invariants
    assert dc.N_loc == r1 - r0 and dc.n_halo == dc.recv_counts.sum()
    assert np.all(np.diff(dc.halo_global.astype(np.int64)) > 0)
    assert not np.any((dc.halo_global >= r0) & (dc.halo_global < r1))
    assert np.array_equal(np.sort(np.unique(cols[rp[r0]:rp[r1]][(cols[rp[r0]:rp[r1]] < r0) | (cols[rp[r0]:rp[r1]] >= r1)])),
                          dc.halo_global)
    A_loc = port.Csr(rs[r0:r1], dc.cols_local, vals[rp[r0]:rp[r1]])       # local ids, entry order untouched
    # column space "lower halo | owned | upper halo": local ids ascend like the global ones, parts on 128-byte lines
    assert dc.off_own % 16 == 0 and dc.off_up % 16 == 0 and dc.ext_len % 16 == 0
    assert dc.n_lo == int((dc.halo_global < r0).sum()) and dc.off_own >= dc.n_lo
    assert dc.off_up >= dc.off_own + dc.N_loc and dc.ext_len >= dc.off_up + (dc.n_halo - dc.n_lo)
    glob = cols[rp[r0]:rp[r1]].astype(np.int64)
    want = np.where(glob < r0, np.searchsorted(dc.halo_global, glob),
                    np.where(glob < r1, dc.off_own + glob - r0,
                             dc.off_up + np.searchsorted(dc.halo_global, glob) - dc.n_lo))
    assert np.array_equal(dc.cols_local, want.astype(np.uint32))
    # visit order of the 32-row slices (nbgpu_dist_plan_visit_order): every slice that reads a halo column is a
    # "late" visit, with the plan's own order (late ones last) and with the solver kernel's (late ones at the end
    # of the last full round, on warps that have no slice in the final partial round)
    import ctypes as C
    from nbots_b200 import capi
    n_sl = (dc.N_loc + 31) // 32
    row_of = np.repeat(np.arange(dc.N_loc), rs[r0:r1])
    halo_entry = (dc.cols_local < dc.off_own) | (dc.cols_local >= dc.off_up)
    halo_slices = set(np.unique(row_of[halo_entry] // 32).tolist())
    assert (world == 1) == (not halo_slices)
    for W in (0, 7, 20, 33, n_sl, 2960):
        sh, lf, lt = C.c_uint32(), C.c_uint32(), C.c_uint32()
        capi.check(capi.lib().nbgpu_dist_plan_visit_order(dc.plan, W, C.byref(sh), C.byref(lf), C.byref(lt)))
        sh, lf, lt = sh.value, lf.value, lt.value
        late = {(v + sh) % n_sl for v in range(min(lf, n_sl), lt)}
        assert halo_slices <= late, (W, sorted(halo_slices - late))
        if not halo_slices:
            assert not late
            continue
        assert sh < n_sl and lf < lt <= n_sl
        if W == 0 or n_sl % W == 0 or n_sl < W or lt - lf > W - n_sl % W:
            assert lt == n_sl                              # no slack to hide them in: they stay last
        else:
            assert lt == (n_sl // W) * W and lf % W >= n_sl % W and lt - lf <= W - n_sl % W

    def exchange(v_loc):
        """halo of a distributed vector: what nbots_b200/csrc/dist.cu does with peer stores, here over gloo"""
        sends = []
        off = 0
        for d in range(world):
            cnt = int(dc.send_counts[d])
            sends.append(v_loc[dc.send_global[off:off + cnt] - r0].copy())
            off += cnt
        everyone = gather((sends, dc.dst_offsets))
        ext = np.full(dc.ext_len, np.nan)                  # padding must never be read
        ext[dc.off_own:dc.off_own + dc.N_loc] = v_loc
        for src in range(world):
            chunk = everyone[src][0][rank]
            assert chunk.size == dc.recv_counts[src]
            pos = int(everyone[src][1][rank])              # where rank src's kernels store into my ext vector
            ext[pos:pos + chunk.size] = chunk
        assert not np.isnan(ext[:dc.n_lo]).any() and not np.isnan(ext[dc.off_up:dc.off_up + dc.n_halo - dc.n_lo]).any()
        return ext

    def allsum(*vals_):
        t = torch.tensor(list(vals_), dtype=torch.float64)
        parts = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return [float(sum(p[i].item() for p in parts)) for i in range(len(vals_))]     # rank order, like dist.cu

    # distributed SpMV == the serial oracle's rows, bit for bit
    x = g["x"]
    y_loc = A_loc.spmv(exchange(x[r0:r1]))
    assert np.array_equal(y_loc, g["spmv_x"][r0:r1])

    # distributed Jacobi-PCG (cg_precond_jacobi.c:13-90) with the plan's halo exchange
    tol = float(g["tol"])
    xl = np.zeros(dc.N_loc)
    gl = A_loc.spmv(exchange(xl)) - b[r0:r1]
    diag = np.array([vals[rp[r0 + i]:rp[r0 + i + 1]][cols[rp[r0 + i]:rp[r0 + i + 1]] == r0 + i][0]
                     for i in range(dc.N_loc)])
    ql = gl / diag
    pl = -ql
    gg, gq = allsum(gl @ gl, gl @ ql)
    gg_test, k = gg, 0
    while gg_test > tol * tol and k < rs.size:
        wl = A_loc.spmv(exchange(pl))
        pw, = allsum(pl @ wl)
        gg_test = gg
        alpha = gq / pw
        xl += alpha * pl
        gl += alpha * wl
        ql = gl / diag
        gg, gq_new = allsum(gl @ gl, gl @ ql)
        pl = -ql + (gq_new / gq) * pl
        gq = gq_new
        k += 1
    xs = gather(xl)
    if rank == 0:
        xg = np.concatenate(xs)
        assert abs(k - int(g["pcg_iters"])) <= max(1, int(0.02 * int(g["pcg_iters"]))), (k, int(g["pcg_iters"]))
        assert rel_l2(xg, g["x"]) <= 1e-10
        print("DIST_OK cpu", k)


def run_gpu(rank, world):
    from nbots_b200 import api, capi
    n_dev = torch.cuda.device_count()
    L = capi.lib()
    capi.check(L.nbgpu_init(rank % n_dev))
    gather = gather_obj_fn(world)
    g = golden("quad_cantilever_64x16")
    prob = multigpu.SlabProblem(64, 16, 4.0, 1.0, rank, world)
    rp = port.row_ptr_of(g["rows_size"]).astype(np.int64)
    r0, r1 = int(prob.row_starts[rank]), int(prob.row_starts[rank + 1])
    # this rank's rows of the global system, bit for bit the reference's
    assert np.array_equal(prob.rows_size, g["rows_size"][r0:r1])
    assert np.array_equal(prob.cols_global, g["cols"][rp[r0]:rp[r1]])
    assert np.array_equal(prob.vals, g["K_post"][rp[r0]:rp[r1]])
    assert np.array_equal(prob.b, g["F_post"][r0:r1])
    dc = multigpu.DistContext(rank, world, prob.row_starts, prob.rows_size, prob.cols_global, prob.vals, gather)
    d_b = api.DeviceBuffer.from_host(prob.b)
    # distributed SpMV: bit-exact rows of the serial product
    d_in = api.DeviceBuffer.from_host(g["x"][r0:r1])
    d_out = api.DeviceBuffer.zeros(prob.N_loc)
    for _ in range(3):
        dc.spmv(d_in, d_out)
    assert np.array_equal(d_out.to_host(), g["spmv_x"][r0:r1])
    tol = float(g["tol"])
    results = []
    for rep in range(2):
        d_x = api.DeviceBuffer.zeros(prob.N_loc)
        st, it, res = dc.pcg_jacobi(d_b, d_x, prob.N_global, tol)
        results.append((st, it, res, d_x.to_host()))
    assert results[0][1] == results[1][1] and np.array_equal(results[0][3], results[1][3])   # reproducible
    st, it, res, xl = results[0]
    st_c, it_c, res_c = dc.cg(d_b, api.DeviceBuffer.zeros(prob.N_loc), prob.N_global, tol)
    # max_iter exit on every rank at the same count
    d_x = api.DeviceBuffer.zeros(prob.N_loc)
    st_cap, it_cap, _ = dc.pcg_jacobi(d_b, d_x, 17, 0.0)
    everyone = gather((st, it, res, xl, st_cap, it_cap, st_c, it_c))
    assert len({(e[0], e[1], e[2]) for e in everyone}) == 1, "ranks disagree on status / iterations / residual"
    if rank == 0:
        xg = np.concatenate([e[3] for e in everyone])
        assert st == 0 and abs(it - int(g["pcg_iters"])) <= max(1, int(0.02 * int(g["pcg_iters"])))
        assert rel_l2(xg, g["x"]) <= 1e-10
        assert (st_cap, it_cap) == (1, 17)
        assert st_c == int(g["cg_status"]) and abs(it_c - int(g["cg_iters"])) <= max(1, int(0.02 * int(g["cg_iters"])))
        print("DIST_OK gpu", it, api.launch_count())
    dc.close()


def run_cpu_fem(rank, world):
    """Host logic of the distributed FEM path, no GPU: the plan nbgpu_dist_plan_from_mesh derives from the mesh alone
    (ghost nodes, column space, local columns, SEND lists) against the plan built from the pattern of this rank's
    rows plus the exchange of halo lists between the ranks (nbgpu_dist_plan_create + set_sends over gloo), on the
    reference's triangle meshes cut into node ranges and on the quad fixtures cut into slabs."""
    import ctypes as C
    from nbots_b200 import capi
    from util import FEM_CASES, mesh_of
    L = capi.lib()
    gather = gather_obj_fn(world)
    u32p = capi.u32p
    for name in FEM_CASES:
        g = golden(name)
        m = mesh_of(g)
        node_starts = np.zeros(world + 1, dtype=np.uint32)
        capi.check(L.nbgpu_partition_nodes(m.n_nod, world, 65 if name == "quad_cantilever_64x16" else 1,
                                           node_starts.ctypes.data_as(u32p)))
        n0, n1 = int(node_starts[rank]), int(node_starts[rank + 1])
        r0, r1 = 2 * n0, 2 * n1
        desc = capi.MeshDesc.of(m)
        rs_f = np.zeros(r1 - r0, dtype=np.uint32)
        ph = C.c_void_p()
        capi.check(L.nbgpu_dist_plan_from_mesh(C.byref(desc), rank, world, node_starts.ctypes.data_as(u32p),
                                               rs_f.ctypes.data_as(u32p), C.byref(ph)))
        F = multigpu.PlanView(ph.value, world)
        cl_f = np.zeros(max(1, F.nnz), dtype=np.uint32)
        capi.check(L.nbgpu_dist_plan_local_cols(ph.value, cl_f.ctypes.data_as(u32p)))
        sc = np.zeros(world, dtype=np.uint32); so = np.zeros(world, dtype=np.uint32)
        capi.check(L.nbgpu_dist_plan_sends(ph.value, sc.ctypes.data_as(u32p), None, so.ctypes.data_as(u32p)))
        sg = np.zeros(max(1, int(sc.sum())), dtype=np.uint32)
        capi.check(L.nbgpu_dist_plan_sends(ph.value, sc.ctypes.data_as(u32p), sg.ctypes.data_as(u32p), so.ctypes.data_as(u32p)))
        # the same through the pattern of this rank's rows and the list exchange
        rs, cols = g["rows_size"], g["cols"]
        rp = port.row_ptr_of(rs).astype(np.int64)
        dc = multigpu.DistContext(rank, world, 2 * node_starts, rs[r0:r1], cols[rp[r0]:rp[r1]], None, gather)
        assert np.array_equal(rs_f, rs[r0:r1]), name
        assert (F.N_loc, F.n_halo, F.nnz) == (dc.N_loc, dc.n_halo, int(rp[r1] - rp[r0])), name
        assert (F.n_lo, F.off_own, F.off_up, F.ext_len) == (dc.n_lo, dc.off_own, dc.off_up, dc.ext_len), name
        assert np.array_equal(F.halo_global, dc.halo_global) and np.array_equal(F.recv_counts, dc.recv_counts), name
        assert np.array_equal(cl_f[:F.nnz], dc.cols_local), name
        assert np.array_equal(sc, dc.send_counts) and np.array_equal(sg[:int(sc.sum())], dc.send_global), name
        live = sc > 0
        assert np.array_equal(so[live], dc.dst_offsets[live]), name
        L.nbgpu_dist_plan_destroy(ph.value)
        dc.close()
    if rank == 0:
        print("DIST_OK cpu-fem")


def run_gpu_fem(rank, world):
    """nbgpu_dist_fem_*: device-side assembly of the rank-local rows (structured slabs AND the reference's
    unstructured triangle meshes cut into contiguous node ranges), bit for bit the reference's rows; then the
    distributed solve against the reference's run."""
    from nbots_b200 import api, capi
    from util import FEM_CASES, bc_records, flatten_bcs, mesh_of
    n_dev = torch.cuda.device_count()
    L = capi.lib()
    capi.check(L.nbgpu_init(rank % n_dev))
    gather = gather_obj_fn(world)
    for name in FEM_CASES:
        g = golden(name)
        m = mesh_of(g)
        neu_dof, neu_add, dir_dof, dir_val = flatten_bcs(m, bc_records(g))
        node_starts = np.zeros(world + 1, dtype=np.uint32)
        align = 65 if name == "quad_cantilever_64x16" else 1
        capi.check(L.nbgpu_partition_nodes(m.n_nod, world, align, node_starts.ctypes.data_as(capi.u32p)))
        assert node_starts[0] == 0 and node_starts[world] == m.n_nod and np.all(np.diff(node_starts.astype(np.int64)) > 0)
        D = api.constitutive_matrix(float(g["E"]), float(g["nu"]), int(g["analysis"]))
        fem = multigpu.DistFem(m, rank, world, node_starts, D, neu_dof, neu_add, dir_dof, dir_val, gather,
                               density=float(g["density"]), self_weight=bool(g["self_weight"]),
                               gravity=tuple(g["gravity"]), thickness=float(g["thickness"]))
        en = g["enabled"] if "enabled" in g.files else None
        st, bad = fem.assemble(enabled=en)
        assert st == 0
        rp = port.row_ptr_of(g["rows_size"]).astype(np.int64)
        r0, r1 = 2 * fem.n0, 2 * fem.n1
        rs, cols_g, vals = fem.rows_global()
        assert np.array_equal(rs, g["rows_size"][r0:r1]), name
        assert np.array_equal(cols_g, g["cols"][rp[r0]:rp[r1]]), name
        assert np.array_equal(vals, g["K_post"][rp[r0]:rp[r1]]), name          # bit-exact after assembly + BCs
        assert np.array_equal(fem.rhs(), g["F_post"][r0:r1]), name
        assert fem.A.blocked, name
        tol = 1e-8 * float(np.linalg.norm(g["F_post"]))
        st, it, res = fem.solve(tol=tol)
        ost, ox, oit, ores = port.Csr(g["rows_size"], g["cols"], g["K_post"]).pcg_jacobi(g["F_post"], tol=tol)
        xs = gather(fem.results())
        its = gather((st, it, res))
        assert len(set(its)) == 1, "ranks disagree"
        if rank == 0:
            xg = np.concatenate(xs)
            assert st == ost == 0 and abs(it - oit) <= max(1, int(np.ceil(0.02 * oit))), (name, it, oit)
            assert rel_l2(xg, ox) <= (1e-9 if name == "beam_cantilever_trg1000" else 1e-10), name
        # a second assembly + warm-started solve (the session call pattern): converged already -> few iterations
        st, bad = fem.assemble(enabled=en)
        st, it2, res2 = fem.solve(warm_start=True, tol=tol)
        assert st == 0 and it2 <= max(3, it // 10), (name, it2, it)
        fem.close()
    if rank == 0:
        print("DIST_OK gpu-fem")


def main():
    mode = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    try:
        {"cpu": run_cpu, "cpu-fem": run_cpu_fem, "gpu": run_gpu, "gpu-fem": run_gpu_fem}[mode](rank, world)
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
