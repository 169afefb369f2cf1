"""GPU: the drop-in claim itself.  In a fresh process libnbots_b200.so is loaded in front of the UNMODIFIED
reference library (oracle/_ref/libnbots_ref.so); the reference's own call sites -- its FEM driver, its
assembly pipeline entry, its solver entry -- then resolve to the shim and run on the device, on genuine
reference objects (nb_sparse_t, nb_mesh2D_t, nb_bcond_t, nb_material_t, nb_fem_elem_t).  Results are compared
with the golden vectors the same reference produced on the CPU."""
import os
import subprocess
import sys
import textwrap

import pytest

from oracle import ref

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import ctypes as C, os, sys
    import numpy as np
    ROOT = %r
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    from nbots_b200 import capi
    C.CDLL(capi.LIB_PATH, mode=C.RTLD_GLOBAL)
    shim = C.CDLL(capi.SHIM_PATH, mode=C.RTLD_GLOBAL)        # in front of the reference
    from oracle import ref
    from util import bc_records, golden, mesh_of, rel_l2
    C.CDLL(ref.LIB_PATH, mode=C.RTLD_GLOBAL)                 # as if the program were linked against libnbots
    L = ref.lib()
    # the reference's PLT entries now point at the shim
    for name in ("nb_sparse_solve_CG_precond_Jacobi", "nb_fem_compute_2D_Solid_Mechanics",
                 "pipeline_assemble_system", "nb_fem_interpolate_from_gpoints_to_nodes"):
        assert C.cast(getattr(shim, name), C.c_void_p).value != C.cast(getattr(L, name), C.c_void_p).value
    before = capi.lib().nbgpu_launch_count()
    for name in ("beam_cantilever_trg1000", "quad_void_selfweight_24x8", "plate_with_hole_trg1000"):
        g = golden(name)
        m = mesh_of(g)
        rm = ref.RefMesh.from_arrays(m)
        kind = m.kind
        en = g["enabled"] if "enabled" in g.files else None
        kw = dict(density=float(g["density"]), self_weight=bool(g["self_weight"]), gravity=tuple(g["gravity"]),
                  analysis=int(g["analysis"]), thickness=float(g["thickness"]), enabled=en)
        # (1) reference harness -> pipeline_assemble_system (shim) on a genuine nb_sparse_t
        K = ref.RefSparse.from_mesh(rm)
        st, F = ref.assemble(K, rm, kind, float(g["E"]), float(g["nu"]), **kw)
        assert st == 0 and np.array_equal(K.export()[2], g["K_pre"]) and np.array_equal(F, g["F_pre"]), name
        # (1b) the same entry with the lumped mass vector wanted (M != NULL, pipeline.c:56-57, :216-222, :256-259)
        gm = golden("lumped_mass")
        K2 = ref.RefSparse.from_mesh(rm)
        st, F2, M2 = ref.assemble_with_mass(K2, rm, kind, float(g["E"]), float(g["nu"]), density=float(gm["density"]),
                                            self_weight=True, gravity=(0.3, -9.81), analysis=int(g["analysis"]),
                                            thickness=float(gm["thickness"]), enabled=gm[name + "/mask"])
        assert st == 0 and np.array_equal(M2, gm[name + "/masked/M"]), name
        # (2) the reference's own BC code (host C, not shimmed), then its solver entry -> shim
        bc = ref.RefBcond()
        for r in bc_records(g):
            bc.push_kirsch(r[2], r[5] - 1) if r[5] else bc.push(*r[:5])
        ref.set_bconditions(rm, K, F, bc)
        assert np.array_equal(K.export()[2], g["K_post"]) and np.array_equal(F, g["F_post"])
        tol = 1e-8 * float(np.linalg.norm(F))
        st, x, it, res = K.pcg_jacobi(F, tol=tol)
        assert st == 0 and res <= tol
        assert np.linalg.norm(K.spmv(x) - F) <= 2 * tol            # spmv entry -> shim as well
        # (3) the reference's FEM driver entry -> shim -> whole pipeline on the device
        ngp = 4 if kind else 1
        st, disp, strain = ref.fem_static(rm, kind, float(g["E"]), float(g["nu"]), bc, n_nod=m.n_nod,
                                          n_elems=m.n_elems, n_gp=ngp, **kw)
        assert st == 0
        if float(g["tol"]) == 1e-8:     # the driver's fixed tolerance (static_elasticity2D.c:88)
            assert rel_l2(disp, g["x"]) <= 1e-10 and rel_l2(strain, g["strain"]) <= 1e-9, name
        # (4) post-processing entry: Gauss points -> nodes through the shim, on the reference's mesh object
        st, nodal = ref.gp_to_nodes(rm, kind, m.n_nod, 3, g["stress"])
        assert st == 0 and np.array_equal(nodal, g["stress_nod"]), name
        # (5) nb_fem_compute_stress_from_strain through the shim (static_elasticity2D.c:99-127)
        assert C.cast(shim.nb_fem_compute_stress_from_strain, C.c_void_p).value != \
            C.cast(L.nb_fem_compute_stress_from_strain, C.c_void_p).value
        stress = ref.stress_from_strain(m.n_elems, kind, float(g["E"]), float(g["nu"]), int(g["analysis"]),
                                        g["strain"], enabled=en)
        assert np.array_equal(stress, g["stress"]), name
        if name.startswith("beam"):
            assert abs(np.sqrt((disp.reshape(-1, 2) ** 2).sum(axis=1)).max() - 1.00701e-1) < 1e-6
    launched = capi.lib().nbgpu_launch_count() - before
    assert launched > 1000, "the calls did not reach the device library"
    print("DROPIN_OK", launched)
''')


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnbots_ref.so not present")
def test_reference_call_sites_run_on_the_device(nbgpu_lib):
    out = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DROPIN_OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnbots_ref.so not present")
def test_reference_call_sites_run_on_two_devices(nbgpu_lib):
    """The same interposition run with NBGPU_DEVICES=2: the reference's single-threaded callers (its FEM driver,
    its solver entry on a genuine nb_sparse_t) are served by two GPUs of the box -- one worker thread per GPU
    inside the shim, peer access between the windows -- and must produce the same results."""
    if nbgpu_lib.nbgpu_device_count() < 2:
        pytest.skip("needs two GPUs in the box")
    env = dict(os.environ, NBGPU_DEVICES="2", NBGPU_MULTI_MIN_ROWS="0", NBGPU_DIST_TIMEOUT_MS="60000")
    out = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and "DROPIN_OK" in out.stdout, out.stdout + out.stderr


OTHER_CALLERS = textwrap.dedent('''
    import ctypes as C, json, os, sys
    import numpy as np
    ROOT = %r
    WITH_SHIM = %d
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    if WITH_SHIM:
        from nbots_b200 import capi
        C.CDLL(capi.LIB_PATH, mode=C.RTLD_GLOBAL)
        shim = C.CDLL(capi.SHIM_PATH, mode=C.RTLD_GLOBAL)        # in front of the reference
        capi.check(capi.lib().nbgpu_set_reduction_order(WITH_SHIM - 1))   # 1: parallel tree, 2: reference order
    from oracle import ref
    from util import golden
    from nbots_b200 import meshgen
    C.CDLL(ref.LIB_PATH, mode=C.RTLD_GLOBAL)
    out = {}
    # (a) inverse power iteration (eigen/inv_power.c:55-105): Jacobi-PCG (:75) and plain CG (:80) inside
    g = golden("lap9_48")
    n_adj, adj = meshgen.laplacian9_graph(48)
    A = ref.RefSparse.from_graph(n_adj, adj, 1)
    A.set_values(g["vals"])
    for name, jac in (("ipower_cgj", True), ("ipower_cg", False)):
        st, vecs, vals, its = ref.inv_power(A, 3, mu=0.0, tolerance=1e-6, use_jacobi=jac)
        out[name] = {"status": st, "vals": vals.tolist(), "its": its, "vecs": vecs.ravel().tolist()}
    # (b) the model regulariser's solve (modules2D/regularizer.c:24-57): a jagged closed polygon, four vertices
    # pinned.  nb_model_regularize itself crashes in the reference (nb_model_load_vtx_graph, model2D.c:276 hands
    # the container array to nb_container_init), so its system is restated here -- matrix from the reference's
    # nb_sparse_create on the vertex graph, entries as regularizer.c:74-108, pinned vertices through the
    # reference's nb_sparse_set_Dirichlet_condition -- and solved with ITS call: x0 = the current vertices,
    # max_iter = N, absolute tolerance 1e-12 (regularizer.c:49).
    rng = np.random.default_rng(7)
    nv, lam = 400, 0.5
    t = np.linspace(0, 2 * np.pi, nv, endpoint=False)
    r = 1.0 + 0.05 * rng.standard_normal(nv)
    vertex = np.stack([r * np.cos(t), r * np.sin(t)], axis=1).ravel()
    nxt, prv = (np.arange(nv) + 1) %% nv, (np.arange(nv) - 1) %% nv
    n_adj = np.full(nv, 2, dtype=np.uint32)
    adj = np.sort(np.stack([prv, nxt], axis=1), axis=1).astype(np.uint32).ravel()
    K = ref.RefSparse.from_graph(n_adj, adj, 2)
    rs, cols, _ = K.export()
    rp = np.zeros(rs.size + 1, dtype=np.int64); np.cumsum(rs, out=rp[1:])
    vals = np.zeros(cols.size); b = np.zeros(2 * nv)
    def add(i, j, v):
        k = rp[i] + int(np.searchsorted(cols[rp[i]:rp[i + 1]], j)); assert cols[k] == j; vals[k] += v
    for e in range(nv):                                     # edge e = (e, e + 1)
        vi, vj = e, int(nxt[e])
        for a in (0, 1):
            add(2 * vi + a, 2 * vi + a, 1.0); add(2 * vi + a, 2 * vj + a, -lam)
            add(2 * vj + a, 2 * vi + a, -lam); add(2 * vj + a, 2 * vj + a, 1.0)
            b[2 * vi + a] += (1.0 - lam) * vertex[2 * vi + a]
            b[2 * vj + a] += (1.0 - lam) * vertex[2 * vj + a]
    K.set_values(vals)
    for fv in (0, 100, 200, 300):
        for a in (0, 1):
            K.dirichlet(b, 2 * fv + a, vertex[2 * fv + a])
    st, v, it, res = K.pcg_jacobi(b, x0=vertex, max_iter=2 * nv, tol=1e-12)
    out["regularize"] = {"status": st, "vertex": v.tolist(), "iters": it}
    # (c) the disk-packing mesher's Newton step (elements2D/mshpack.c:645-679): THREE unknowns per node
    # (nb_sparse_create(&graph, NULL, 3): no 2x2 block structure), solver called with the previous step as initial
    # guess, max_iter = 2 N, absolute tolerance 1e-8
    n_adj3, adj3 = meshgen.laplacian9_graph(20)
    H = ref.RefSparse.from_graph(n_adj3, adj3, 3)
    rs3, cols3, _ = H.export()
    row_of = np.repeat(np.arange(rs3.size), rs3)
    lo, hi = np.minimum(row_of, cols3), np.maximum(row_of, cols3)
    vals3 = -0.05 * (1.0 + ((lo * 7919 + hi * 104729) %% 97) / 97.0)          # symmetric by construction
    vals3[row_of == cols3] = 3.0                                               # diagonally dominant: SPD
    H.set_values(vals3)
    b3 = meshgen.uniform_rhs(rs3.size, seed=11)
    st, h1, it, res = H.pcg_jacobi(b3, x0=0.01 * meshgen.uniform_rhs(rs3.size, seed=12), max_iter=2 * rs3.size, tol=1e-8)
    out["mshpack_step"] = {"status": st, "h": h1.tolist(), "iters": it}
    if WITH_SHIM:
        out["launches"] = int(capi.lib().nbgpu_launch_count())
    print("RESULT " + json.dumps(out))
''')


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnbots_ref.so not present")
def test_other_krylov_callers_of_the_reference_through_the_shim(nbgpu_lib):
    """the reference's other users of the two solver entry points -- inverse power iteration
    (inv_power.c:75 Jacobi-PCG, :80 plain CG), run unmodified, and the model regulariser's solve (regularizer.c:49;
    its system restated, see the script) and the disk-packing mesher's Newton step (mshpack.c:676: three unknowns
    per node, warm start, max_iter = 2 N) -- with libnbots_b200.so interposed; results against the same calls on the CPU: bit-identical with reference-order
    dot products, within solver tolerance with the default parallel-tree reductions."""
    import json

    import numpy as np

    def run(with_shim):
        out = subprocess.run([sys.executable, "-c", OTHER_CALLERS % (ROOT, with_shim)], capture_output=True, text=True,
                             timeout=900)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        return json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("RESULT ")][-1][7:])

    cpu, tree, exact = run(0), run(1), run(2)
    assert tree["launches"] > 1000 and exact["launches"] > 1000, "the calls did not reach the device library"
    for key in ("ipower_cgj", "ipower_cg"):
        assert cpu[key]["status"] == exact[key]["status"] == tree[key]["status"]
        assert exact[key]["vals"] == cpu[key]["vals"] and exact[key]["its"] == cpu[key]["its"], key
        assert exact[key]["vecs"] == cpu[key]["vecs"], key
        np.testing.assert_allclose(tree[key]["vals"], cpu[key]["vals"], rtol=1e-4)
    assert exact["regularize"]["status"] == cpu["regularize"]["status"] == tree["regularize"]["status"] == 0
    assert exact["regularize"]["vertex"] == cpu["regularize"]["vertex"]
    assert exact["regularize"]["iters"] == cpu["regularize"]["iters"]
    assert abs(tree["regularize"]["iters"] - cpu["regularize"]["iters"]) <= max(1, 0.02 * cpu["regularize"]["iters"])
    assert exact["mshpack_step"] == cpu["mshpack_step"] and cpu["mshpack_step"]["status"] == 0
    assert abs(tree["mshpack_step"]["iters"] - cpu["mshpack_step"]["iters"]) <= 1
    np.testing.assert_allclose(tree["mshpack_step"]["h"], cpu["mshpack_step"]["h"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(tree["regularize"]["vertex"], cpu["regularize"]["vertex"], rtol=0, atol=1e-9)


def test_shim_entry_points_tolerate_concurrent_callers(nbgpu_lib):
    """The reference's solver entry points are re-entrant; the shim serialises callers that arrive from
    several host threads (one device context per process) instead of corrupting its work vectors."""
    import ctypes as C
    import threading

    import numpy as np

    import bench
    from nbots_b200 import capi
    from util import golden

    shim = C.CDLL(capi.SHIM_PATH)
    pcg = shim.nb_sparse_solve_CG_precond_Jacobi
    pcg.restype = C.c_int
    pcg.argtypes = [C.POINTER(bench.NbSparse), capi.f64p, capi.f64p, C.c_uint32, C.c_double, capi.u32p, capi.f64p,
                    C.c_uint32]
    mv = shim.nb_sparse_multiply_vector
    mv.restype = None
    mv.argtypes = [C.POINTER(bench.NbSparse), capi.f64p, capi.f64p, C.c_uint32]
    cases = []
    for name in ("quad_cantilever_64x16", "plate_with_hole_trg1000", "quad_void_selfweight_24x8"):
        g = golden(name)
        rs, cols, vals = g["rows_size"].copy(), g["cols"].copy(), g["K_post"].copy()
        A = bench.host_nb_sparse(shim, rs, cols, vals)      # the reference's layout: two heap blocks per row
        cases.append((g, A, None, rs, cols, vals))
    errors = []

    def worker(idx):
        g, A, keep, rs, cols, vals = cases[idx % len(cases)]
        b = np.ascontiguousarray(g["F_post"])
        tol = 1e-8 * float(np.linalg.norm(b))
        try:
            for _ in range(6):
                x = np.zeros(rs.size)
                it = C.c_uint32(0)
                res = C.c_double(0)
                st = pcg(A, b.ctypes.data_as(capi.f64p), x.ctypes.data_as(capi.f64p), rs.size, tol,
                         C.byref(it), C.byref(res), 1)
                y = np.zeros(rs.size)
                mv(A, x.ctypes.data_as(capi.f64p), y.ctypes.data_as(capi.f64p), 1)
                if st != 0 or np.linalg.norm(y - b) > 2 * tol:
                    errors.append((idx, st, float(np.linalg.norm(y - b)), tol))
        except Exception as e:                       # pragma: no cover
            errors.append((idx, repr(e)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("env", [{"NBGPU_NO_POOL": "1"}, {"NBGPU_SPMV_PATH": "reg"}, {"NBGPU_NO_PDL": "1"}])
def test_fallback_configurations_still_pass_smoke(nbgpu_lib, env):
    """Process-wide switches (plain cudaMalloc instead of the pool, register-path SpMV, no programmatic
    dependent launch) are read once per process: run the smoke check under each in a fresh interpreter."""
    e = dict(os.environ)
    e.update(env)
    out = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=ROOT, env=e,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "smoke ok" in out.stdout, out.stdout + out.stderr
