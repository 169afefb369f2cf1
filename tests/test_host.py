"""CPU: the product's host-side logic and the C-ABI surface (no compute calls: there is no GPU here
and the library has no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nbots_b200 import api, capi, meshgen
from oracle import port
from util import FEM_CASES, bc_records, flatten_bcs, golden, mesh_of


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    text = open(capi.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = sorted(set(re.findall(r"\b(nbgpu_[a-z0-9_]+)\s*\(", text)))
    assert len(declared) > 40
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, f"declared in include/nbgpu.h but not exported: {missing}"
    bound = set(capi.SIGNATURES) | {"nbgpu_bcond_flatten", "nbgpu_fem_static_elasticity2d",
                                    "nbgpu_fem_static_elasticity2d_lists"}
    assert not [n for n in declared if n not in bound and not n.startswith("nbgpu_comm")], \
        "every declared entry point needs a ctypes signature"


def test_shim_exports_reference_names():
    S = C.CDLL(capi.SHIM_PATH)
    for name in ("nb_sparse_solve_CG_precond_Jacobi", "nb_sparse_solve_conjugate_gradient",
                 "nb_sparse_multiply_vector", "pipeline_assemble_system", "nb_fem_compute_2D_Solid_Mechanics"):
        assert hasattr(S, name)


def test_compute_fails_loudly_without_a_device():
    L = capi.lib()
    if L.nbgpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    rs, cols, vals = meshgen.laplacian9_csr(4)
    with pytest.raises(capi.NbgpuError) as e:
        api.Matrix.from_csr(rs, cols, vals)
    assert e.value.code == capi.ERR_CUDA and "no CPU fallback" in str(e.value)


@pytest.mark.parametrize("name", FEM_CASES)
@pytest.mark.parametrize("use_edges", [True, False])
def test_pattern_builder_is_bit_exact(name, use_edges):
    g = golden(name)
    rs, cols = api.pattern_from_mesh(mesh_of(g), 2, use_edges=use_edges)
    assert np.array_equal(rs, g["rows_size"]) and np.array_equal(cols, g["cols"])


def test_pattern_builder_scalar_and_ragged():
    g = golden("lap9_48")
    # vars_per_node = 1 on a quad grid gives the 9-point pattern of config 3
    m = meshgen.structured_mesh(47, 47, 1.0, 1.0, kind=1)
    rs, cols = api.pattern_from_mesh(m, 1)
    assert np.array_equal(rs, g["rows_size"]) and np.array_equal(cols, g["cols"])
    # isolated node (no element touches it): a row holding only its diagonal, like nb_sparse_create
    m = meshgen.structured_mesh(2, 1, 2.0, 1.0, kind=0)
    m.nod = np.concatenate([m.nod, [9.0, 9.0]])
    rs, cols = api.pattern_from_mesh(m, 2)
    prs, pcols = port.pattern_from_mesh(m, 2)
    assert np.array_equal(rs, prs) and np.array_equal(cols, pcols) and rs[-1] == 2


def test_pattern_builder_list_cache_is_validated():
    """The sizing call keeps its sorted node lists for the fill call; a mesh that changes in place between
    the two calls (same buffers, same sizes) must not be served from that cache, and high-degree nodes take
    the general sort."""
    import ctypes as C
    L = capi.lib()
    a = meshgen.structured_mesh(40, 30, 2.0, 1.0, kind=0, diagonal_seed=1)
    b = meshgen.structured_mesh(40, 30, 2.0, 1.0, kind=0, diagonal_seed=2)
    adj = a.adj.copy()
    rs = np.empty(a.n_nod * 2, dtype=np.uint32)
    nnz = C.c_uint64(0)

    def call(cols):
        capi.check(L.nbgpu_pattern_from_mesh(a.n_nod, a.n_elems, 3, adj.ctypes.data_as(capi.u32p), 0, None, 2,
                                             rs.ctypes.data_as(capi.u32p),
                                             None if cols is None else cols.ctypes.data_as(capi.u32p), C.byref(nnz)))

    call(None)                                             # sizing call on mesh a: lists cached
    adj[:] = b.adj                                         # same buffer, other triangles
    call(None)
    cols = np.empty(nnz.value, dtype=np.uint32)
    call(cols)
    prs, pcols = port.pattern_from_mesh(b)                 # elements only would differ from edges+elements
    brs, bcols = api.pattern_from_mesh(b, use_edges=False)
    assert np.array_equal(rs, brs) and np.array_equal(cols, bcols)
    assert np.array_equal(brs, prs) and np.array_equal(bcols, pcols)
    # fill call with a changed mesh and NO new sizing call: rebuilt, not reused
    adj[:] = a.adj
    rs_a, cols_a = api.pattern_from_mesh(a, use_edges=False)
    cols2 = np.empty(cols_a.size, dtype=np.uint32)
    call(cols2)
    assert np.array_equal(rs, rs_a) and np.array_equal(cols2, cols_a)
    # a fan: node 0 belongs to 40 triangles (more neighbours than the small-list sort handles)
    k = 40
    ang = np.linspace(0, 2 * np.pi, k, endpoint=False)
    nod = np.concatenate([[0.0, 0.0], np.stack([np.cos(ang), np.sin(ang)], axis=1).ravel()])
    tri = np.array([[0, 1 + i, 1 + (i + 1) % k] for i in range(k)], dtype=np.uint32)
    edges = np.array([[0, 1 + i] for i in range(k)] + [[1 + i, 1 + (i + 1) % k] for i in range(k)], dtype=np.uint32)
    fan = meshgen.Mesh2D(kind=0, nod=nod, edg=edges.ravel().copy(), adj=tri.ravel().copy(),
                         vtx=np.zeros(0, np.uint32), sgm_sizes=np.zeros(0, np.uint32),
                         sgm_nodes=np.zeros(0, np.uint32), nx=0, ny=0)
    frs, fcols = api.pattern_from_mesh(fan)
    prs, pcols = port.pattern_from_mesh(fan)
    assert frs[0] == 2 * (k + 1) and np.array_equal(frs, prs) and np.array_equal(fcols, pcols)


def test_tables_and_constitutive_match_reference():
    for npe, et in ((3, 0), (4, 1)):
        t = api.elem_tables(npe)
        n, g, w, Ni, dp, de = port.elem_tables(et)
        assert (t.N_nodes, t.N_gp) == (n, g)
        assert np.array_equal(np.array(t.gp_weight[:g]), w) and np.array_equal(np.array(t.Ni[:n * g]), Ni)
        assert np.array_equal(np.array(t.dNi_dpsi[:n * g]), dp) and np.array_equal(np.array(t.dNi_deta[:n * g]), de)
    for name in FEM_CASES:
        g = golden(name)
        # includes analysis = 1 ("plane strain"): the reference still yields the plane-stress D
        assert np.array_equal(api.constitutive_matrix(float(g["E"]), float(g["nu"]), int(g["analysis"])), g["D"])


@pytest.mark.parametrize("name", FEM_CASES)
def test_bcond_flatten_reproduces_reference_bcs(name):
    """The flattened (ordered) dof lists, applied sequentially with the ORACLE's Dirichlet elimination,
    must give the reference's post-BC system bit for bit."""
    g = golden(name)
    m = mesh_of(g)
    neu_dof, neu_add, dir_dof, dir_val = flatten_bcs(m, bc_records(g))
    K = port.Csr(g["rows_size"], g["cols"], g["K_pre"])
    F = g["F_pre"].copy()
    for d, a in zip(neu_dof, neu_add):
        F[d] += a
    for d, v in zip(dir_dof, dir_val):
        K.dirichlet(F, d, v)
    assert np.array_equal(F, g["F_post"]) and np.array_equal(K.vals, g["K_post"])


def test_meshgen_counts():
    for nx, ny in ((64, 16), (7, 3)):
        m = meshgen.structured_mesh(nx, ny, 1.0, 1.0)
        rs, cols = api.pattern_from_mesh(m, 2)
        assert (rs.size, cols.size) == meshgen.quad_counts(nx, ny)
    assert meshgen.quad_counts(1000, 500) == (1003002, 18018004)
    assert meshgen.quad_counts(4000, 2000) == (16012002, 288072004)    # Q16
    a = meshgen.uniform_rhs(100, seed=5)
    assert np.array_equal(a[10:30], meshgen.uniform_rhs(20, seed=5, start=10))


def test_streamed_kernels_fit_two_ctas_per_sm():
    """Static guard on the built library (no GPU needed): every TMA-streamed kernel must stay within the
    register budget of two 320-thread CTAs per SM (65536 / 640 -> 96 after allocation rounding) without
    spilling, and the bulk-copy / mbarrier instructions must be in its SASS."""
    import re
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-res-usage", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    found = push_warp = 0
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        name, reg, stack = m.group(1), int(m.group(2)), int(m.group(3))
        if any(k in name for k in ("spmv_stream_kernel", "init_stream_kernel", "fspmv_kernel", "dist_plain_spmv_kernel")):
            found += 1
            # the row-partitioned variants pass a 2-3 double array to the out-of-line exchange (collect_slots):
            # that argument lives in a 16-24 byte frame; anything else in the frame would be a spill
            assert stack == 0 or ("PeerComm" in name and stack <= 24), (name, "spills")
            assert reg <= 96, (name, reg)
            if "PeerCommELb1" in name:
                # K1 with the extra halo-push warp: two 352-thread CTAs per SM, i.e. 6 warps on the fullest of the four
                # register-file partitions -> at most 80 registers; exists for the blocked layouts only (odd LAYOUT)
                push_warp += 1
                assert reg <= 80 and stack == 0, (name, reg, stack)
                assert re.search(r"Li[13]ENS_8PeerCommELb1", name), name
    # 4 layouts x (spmv, distributed spmv) + 4 layouts x 2 exchange policies x (classic K1, 2 x fused K1, 2 x init)
    assert found >= 48
    assert push_warp == 6    # 2 blocked layouts x (classic K1, 2 x fused K1)
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "nbgpu::spmv_stream_kernel<(int)3>", capi.LIB_PATH],
                          capture_output=True, text=True).stdout
    if "UBLKCP" not in sass:    # older cuobjdump builds want the mangled name
        sass = subprocess.run([cuobjdump, "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass


def test_committed_bench_lines_keep_the_contract():
    """The bench lines under profiles/ (what DESIGN.md and the summaries quote) carry every key of the bench contract,
    one peak number, green parity blocks and iteration counts that do not depend on the GPU count's arithmetic."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    baseline = json.load(open(os.path.join(root, "BASELINE.json")))
    iters = {1: 4543, 2: 5965, 4: 7806, 8: 10319}
    for n in (1, 2, 4, 8):
        for suffix in ("", "_fused") if n > 1 else ("",):
            line = json.loads(open(os.path.join(root, "profiles", "r02_bench_n%d%s.json" % (n, suffix))).read())
            for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                        "scaling", "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
                assert key in line, (n, suffix, key)
            assert line["n_gpus"] == n and line["dtype"] == "f64" and line["scaling"] == "weak"
            assert line["metric"] == "pcg_dof_iter_per_s" and line["unit"] == "DOF*iter/s" and "DOF" in baseline["metric"]
            assert line["vs_baseline"] is None and not baseline["published"]      # no published number to divide by
            assert line["warmup"] >= 3 and line["gpu_launches"] > 0
            assert line["roofline"]["peak"] == 6554.9 and line["roofline"]["bound"] == "hbm"
            assert line["config"]["iterations_per_step"] == iters[n]
            assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
            e2e = line["e2e"]
            assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and 0 < e2e["value"] < line["value"]
            if n == 1:
                assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["value"] > 0
                assert line["roofline"]["traffic"] and line["roofline"]["traffic_source"]["commit"]
                assert line["target"]["l64_spmv"]["A_times_ones_matches_closed_form"]
            else:
                p = line["parity"]
                assert p["ok"] and p["rows_bit_identical_to_single_gpu"] and p["K_bit_identical_to_cpu_reference"]
                assert p["iterations"] == p["iterations_single_gpu"] == iters[n]
            if suffix == "":
                q = line["target"]["q16"]
                layout = q["layout"] if n == 1 else q["local_block_layout"]
                assert q["N_dof"] == 16012002 and q["iterations"] == 1000 and layout["idx16"] and layout["blocked"]
                if n > 1:
                    assert 0.5 < q["strong_scaling_efficiency"] <= 1.05
    ref_line = json.loads(open(os.path.join(root, "profiles", "r02_bench_reference_n1.json")).read())
    assert ref_line["impl"] == "reference" and ref_line["cpu_baseline"]["kind"] == "reference"
