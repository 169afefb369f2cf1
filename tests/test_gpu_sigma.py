"""SELL-32-sigma row ordering (matrix.cuh): storing the rows of a window by descending length must be
invisible above the C ABI.  Every check of the irregular (triangle-mesh) cases is repeated with the
sorting off, with a small window and with the default window, and has to give the same bits."""
import numpy as np
import pytest

from nbots_b200 import api, capi, meshgen
from oracle import port
from util import bc_records, flatten_bcs, golden, rel_l2

from test_gpu_parity import assemble_case, sequential_dots  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu

TRIANGLE_CASES = ["beam_cantilever_trg1000", "plate_with_hole_trg1000"]


@pytest.fixture(params=["1", "64", "256", "1024"])
def sigma(request, monkeypatch):
    monkeypatch.setenv("NBGPU_SIGMA", request.param)     # read when a matrix is created
    return int(request.param)


def test_default_window_is_chosen_by_padding(nbgpu_lib, monkeypatch):
    monkeypatch.delenv("NBGPU_SIGMA", raising=False)
    g = golden("beam_cantilever_trg1000")                 # 27 % padding in natural order
    A = api.Matrix.from_csr(g["rows_size"], g["cols"], g["K_post"])
    assert A.sigma == 256 and A.blocked
    assert A.stored < 1.08 * A.nnz
    g = golden("quad_cantilever_64x16")                   # structured: nothing to gain
    B = api.Matrix.from_csr(g["rows_size"], g["cols"], g["K_post"])
    assert B.sigma == 1


@pytest.mark.parametrize("name", TRIANGLE_CASES)
def test_sorted_rows_keep_every_result_bit_exact(nbgpu_lib, sequential_dots, sigma, name):
    g = golden(name)
    m, mesh, K, d_F = assemble_case(g, capi.ASSEMBLY_GATHER)
    assert K.sigma == sigma and K.blocked                 # row pairs stay together
    assert np.array_equal(K.pattern_csr()[1], g["cols"])
    assert np.array_equal(K.values_csr(), g["K_pre"])
    assert np.array_equal(d_F.to_host(), g["F_pre"])
    neu_dof, neu_add, dir_dof, dir_val = flatten_bcs(m, bc_records(g))
    api.vector_add_entries(d_F, neu_dof, neu_add)
    K.apply_dirichlet(d_F, dir_dof, dir_val)
    assert np.array_equal(K.values_csr(), g["K_post"])
    assert np.array_equal(d_F.to_host(), g["F_post"])
    assert np.array_equal(K.spmv_host(g["x"]), g["spmv_x"])
    st, x, it, res = K.pcg_jacobi_host(g["F_post"], tol=float(g["tol"]))
    assert (st, it, res) == (int(g["pcg_status"]), int(g["pcg_iters"]), float(g["pcg_res"]))
    assert np.array_equal(x, g["x"])


@pytest.mark.parametrize("name", TRIANGLE_CASES)
@pytest.mark.parametrize("mode", [capi.ASSEMBLY_ATOMIC, capi.ASSEMBLY_COLOR])
def test_element_parallel_assembly_finds_sorted_rows(nbgpu_lib, sigma, name, mode):
    g = golden(name)
    m, mesh, K, d_F = assemble_case(g, mode)
    assert K.sigma == sigma
    assert rel_l2(K.values_csr(), g["K_pre"]) <= 1e-14


@pytest.mark.parametrize("path", ["stream", "reg"])
def test_ragged_matrix_any_window(nbgpu_lib, monkeypatch, path):
    """Not a FEM pattern: odd row count, empty rows, lengths 0..40 -> single-row sorting, unblocked."""
    monkeypatch.setenv("NBGPU_SPMV_PATH", path)
    rng = np.random.default_rng(11)
    N = 3001
    rs = rng.integers(0, 41, N).astype(np.uint32)
    rs[rng.integers(0, N, 50)] = 0
    cols = np.concatenate([np.sort(rng.choice(N, size=k, replace=False)) for k in rs]).astype(np.uint32)
    vals = rng.standard_normal(cols.size)
    x = rng.standard_normal(N)
    want = port.Csr(rs, cols, vals).spmv(x)
    for s in ("1", "32", "96", "4096"):
        monkeypatch.setenv("NBGPU_SIGMA", s)
        A = api.Matrix.from_csr(rs, cols, vals)
        # a window of one slice cannot shorten any slice: the layout keeps the natural order
        assert A.sigma == (int(s) if s not in ("1", "32") else 1) and not A.blocked
        assert np.array_equal(A.values_csr(), vals) and np.array_equal(A.pattern_csr()[1], cols)
        assert np.array_equal(A.spmv_host(x), want)


def test_large_irregular_triangle_mesh(nbgpu_lib, monkeypatch):
    """300k-dof mesh with random diagonals (4..8 neighbours per node): pattern, assembly and SpMV against
    the port, default window."""
    monkeypatch.delenv("NBGPU_SIGMA", raising=False)
    m = meshgen.structured_mesh(500, 300, 5.0, 3.0, kind=0, diagonal_seed=3)
    rs, cols = api.pattern_from_mesh(m)
    prs, pcols = port.pattern_from_mesh(m)
    assert np.array_equal(rs, prs) and np.array_equal(cols, pcols)
    K = api.Matrix.from_csr(rs, cols)
    assert K.sigma == 256 and K.blocked and K.stored < 1.06 * K.nnz
    d_F = api.DeviceBuffer.zeros(K.N)
    st, bad = api.Mesh(m).assemble(K, d_F, 2.0e5, 0.3)
    assert st == 0
    P = port.Csr(rs, cols)
    pst, F = port.assemble(P, m, 2.0e5, 0.3)
    assert pst == 0 and np.array_equal(d_F.to_host(), F)
    assert np.array_equal(K.values_csr(), P.vals)
    x = meshgen.uniform_rhs(K.N)
    assert np.array_equal(K.spmv_host(x), P.spmv(x))


# ---------------------------------------------------------------- 16-bit column ids --

@pytest.mark.parametrize("name", TRIANGLE_CASES + ["quad_cantilever_64x16", "lap9_48"])
@pytest.mark.parametrize("ids", ["16", "32"])
def test_column_id_width_is_invisible(nbgpu_lib, monkeypatch, sequential_dots, name, ids):
    if ids == "32":
        monkeypatch.setenv("NBGPU_NO_IDX16", "1")
    g = golden(name)
    fem = "K_post" in g.files
    A = api.Matrix.from_csr(g["rows_size"], g["cols"], g["K_post"] if fem else g["vals"])
    assert A.idx16 == (ids == "16")
    x, want = (g["x"], g["spmv_x"]) if fem else (g["b"], g["spmv_b"])
    assert np.array_equal(A.spmv_host(x), want)
    b = g["F_post"] if fem else g["b"]
    st, sol, it, res = A.pcg_jacobi_host(b, tol=float(g["tol"]))
    assert (st, it, res) == (int(g["pcg_status"]), int(g["pcg_iters"]), float(g["pcg_res"]))
    assert np.array_equal(sol, g["x"])


def test_far_columns_keep_32_bit_ids(nbgpu_lib, monkeypatch):
    """Entries further than 32767 from the diagonal do not fit 16-bit differences: the matrix keeps its
    32-bit ids (decided on the device at creation) and results stay exact."""
    monkeypatch.delenv("NBGPU_NO_IDX16", raising=False)
    rng = np.random.default_rng(9)
    N = 70001
    near = [np.unique(np.clip(i + rng.integers(-40, 41, 6), 0, N - 1)) for i in range(N)]
    far_rows = set(rng.integers(0, N, 25).tolist())
    rows = [np.unique(np.append(c, (i + 40000) % N)) if i in far_rows else c for i, c in enumerate(near)]
    rs = np.array([r.size for r in rows], dtype=np.uint32)
    cols = np.concatenate(rows).astype(np.uint32)
    vals = rng.standard_normal(cols.size)
    x = rng.standard_normal(N)
    A = api.Matrix.from_csr(rs, cols, vals)
    assert not A.idx16 and not A.blocked
    assert np.array_equal(A.spmv_host(x), port.Csr(rs, cols, vals).spmv(x))
    near_only = api.Matrix.from_csr(np.array([r.size for r in near], dtype=np.uint32),
                                    np.concatenate(near).astype(np.uint32))
    assert near_only.idx16
    # the extreme differences that still fit: +-32767
    N2 = 40000
    rows = [np.unique([max(i - 32767, 0), i, min(i + 32767, N2 - 1)]) for i in range(N2)]
    rs2 = np.array([r.size for r in rows], dtype=np.uint32)
    c2 = np.concatenate(rows).astype(np.uint32)
    v2 = rng.standard_normal(c2.size)
    B = api.Matrix.from_csr(rs2, c2, v2)
    x2 = rng.standard_normal(N2)
    assert B.idx16
    assert np.array_equal(B.spmv_host(x2), port.Csr(rs2, c2, v2).spmv(x2))
    rows[5] = np.unique([5, 5 + 32768])
    rs3 = np.array([r.size for r in rows], dtype=np.uint32)
    c3 = np.concatenate(rows).astype(np.uint32)
    assert not api.Matrix.from_csr(rs3, c3).idx16


@pytest.mark.parametrize("ids", ["16", "32"])
@pytest.mark.parametrize("window", ["1", "128"])
def test_streamed_path_with_several_gather_batches(nbgpu_lib, monkeypatch, ids, window):
    """Per-entry layout with slices 10..24 entries wide: the streamed kernel then takes its general loop
    (full gather batches + tail) instead of the single-batch path of the 9-point stencil, with both id
    widths and with sorted rows."""
    monkeypatch.setenv("NBGPU_SIGMA", window)
    monkeypatch.setenv("NBGPU_SPMV_PATH", "stream")
    if ids == "32":
        monkeypatch.setenv("NBGPU_NO_IDX16", "1")
    rng = np.random.default_rng(17)
    N = 5003
    rs = rng.integers(0, 25, N).astype(np.uint32)
    rs[:40] = 24                                            # a few slices of full width right away
    rows = [np.unique(np.clip(i + rng.integers(-300, 301, k), 0, N - 1)) for i, k in enumerate(rs)]
    rs = np.array([r.size for r in rows], dtype=np.uint32)
    cols = np.concatenate(rows).astype(np.uint32)
    vals = rng.standard_normal(cols.size)
    A = api.Matrix.from_csr(rs, cols, vals)
    assert not A.blocked and A.idx16 == (ids == "16") and 9 < A.max_width <= 24
    P = port.Csr(rs, cols, vals)
    for seed in (1, 2):
        x = rng.standard_normal(N)
        assert np.array_equal(A.spmv_host(x), P.spmv(x))
    # a diagonally dominant symmetric matrix of the same shape class, solved by both
    n = 3001
    off = [[j for j in (i + rng.integers(-60, 61, 5)).tolist() if 0 <= j < n] for i in range(n)]
    pairs = {(min(i, j), max(i, j)) for i, c in enumerate(off) for j in c if i != j}
    adj = [[] for _ in range(n)]
    for i, j in pairs:
        adj[i].append(j); adj[j].append(i)
    rows = [np.array(sorted(a + [i])) for i, a in enumerate(adj)]
    rs2 = np.array([r.size for r in rows], dtype=np.uint32)
    c2 = np.concatenate(rows).astype(np.uint32)
    v2 = np.concatenate([np.where(r == i, float(r.size) + 1.0, -1.0) for i, r in enumerate(rows)])
    B = api.Matrix.from_csr(rs2, c2, v2)
    assert 9 < B.max_width <= 28 and not B.blocked
    b = rng.standard_normal(n)
    st, x, it, res = B.pcg_jacobi_host(b, tol=1e-10 * float(np.linalg.norm(b)))
    ost, ox, oit, ores = port.Csr(rs2, c2, v2).pcg_jacobi(b, tol=1e-10 * float(np.linalg.norm(b)))
    assert st == ost == 0 and abs(it - oit) <= max(1, int(0.02 * oit))
    assert np.linalg.norm(x - ox) <= 1e-10 * np.linalg.norm(ox)
