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
