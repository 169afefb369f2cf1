"""CPU: pin the oracle.  The plain-C restatement (oracle/port) must reproduce, bit for bit, the
golden vectors that oracle/make_golden.py dumped from the unmodified reference -- including the
known answers of the reference's own FEM tests -- and, when the compiled reference is present
(oracle/_ref), the live reference on freshly generated inputs."""
import numpy as np
import pytest

from nbots_b200 import meshgen
from oracle import port, ref
from util import FEM_CASES, bc_records, golden, mesh_of


@pytest.mark.parametrize("name", FEM_CASES)
def test_port_reproduces_reference_fem_pipeline(name):
    g = golden(name)
    m = mesh_of(g)
    rs, cols = port.pattern_from_mesh(m)
    assert np.array_equal(rs, g["rows_size"]) and np.array_equal(cols, g["cols"])      # bit-exact pattern
    K = port.Csr(rs, cols)
    en = g["enabled"] if "enabled" in g.files else None
    st, F = port.assemble(K, m, float(g["E"]), float(g["nu"]), density=float(g["density"]),
                          self_weight=bool(g["self_weight"]), gravity=tuple(g["gravity"]),
                          analysis=int(g["analysis"]), thickness=float(g["thickness"]), enabled=en)
    assert st == 0
    assert np.array_equal(K.vals, g["K_pre"]) and np.array_equal(F, g["F_pre"])
    port.set_bconditions(m, K, F, bc_records(g))
    assert np.array_equal(K.vals, g["K_post"]) and np.array_equal(F, g["F_post"])
    st, x, it, res = K.pcg_jacobi(F, tol=float(g["tol"]))
    assert (st, it) == (int(g["pcg_status"]), int(g["pcg_iters"])) and res == float(g["pcg_res"])
    assert np.array_equal(x, g["x"])
    st, x_cg, it, res = K.cg(F, tol=float(g["tol"]))
    assert (st, it) == (int(g["cg_status"]), int(g["cg_iters"])) and np.array_equal(x_cg, g["x_cg"])
    assert np.array_equal(K.spmv(x), g["spmv_x"])
    st, strain = port.compute_strain(m, x)
    assert np.array_equal(strain, g["strain"])
    stress = port.stress_from_strain(m.n_elems, m.kind, float(g["E"]), float(g["nu"]), int(g["analysis"]), strain, en)
    assert np.array_equal(stress, g["stress"])
    st, nodal = port.gp_to_nodes(m, 3, stress)                          # gaussp_to_nodes.c:50
    assert st == 0 and np.array_equal(nodal, g["stress_nod"])
    assert np.array_equal(port.vm_stress(stress), g["vm"])             # formulas.c:65-77
    assert np.array_equal(port.main_stress(stress), g["main_stress"])
    assert np.array_equal(port.constitutive(float(g["E"]), float(g["nu"]), int(g["analysis"])), g["D"])


@pytest.mark.parametrize("name", FEM_CASES)
def test_port_reproduces_reference_lumped_mass(name):
    """M of pipeline_assemble_system(K, M != NULL, ...) (pipeline.c:216-222, :256-259): the restatement against the
    vectors oracle/make_golden_mass.py took from the reference, all elements enabled and with a mask."""
    gm = golden("lumped_mass")
    m = mesh_of(golden(name))
    rho, t = float(gm["density"]), float(gm["thickness"])
    for tag, en in (("all", None), ("masked", gm[f"{name}/mask"])):
        st, M = port.lumped_mass(m, rho, t, en)
        assert st == 0 and np.array_equal(M, gm[f"{name}/{tag}/M"])
    # both components of a node carry the same mass; the total is density * thickness * area when nothing is masked
    st, M = port.lumped_mass(m, rho, t, None)
    assert np.array_equal(M[0::2], M[1::2])
    x, y = m.nod[0::2], m.nod[1::2]
    a = m.adj.reshape(-1, m.npe)
    area = 0.0
    for k in range(m.npe):                                   # shoelace over every element
        i, j = a[:, k], a[:, (k + 1) % m.npe]
        area += 0.5 * np.sum(x[i] * y[j] - x[j] * y[i])
    assert abs(M[0::2].sum() - rho * t * area) <= 1e-11 * rho * t * area


@pytest.mark.parametrize("name", ["beam_cantilever_trg1000", "quad_void_selfweight_24x8"])
def test_port_reproduces_reference_damage_assembly(name):
    """The damage driver's assembly loop (static_damage2D.c:474-569): the restatement against K and F taken from the
    reference's own function (oracle/make_golden_damage.py)."""
    gd = golden("damage_assembly")
    g = golden(name)
    m = mesh_of(g)
    kw = dict(density=float(gd["density"]), self_weight=True, gravity=tuple(gd["gravity"]),
              analysis=int(g["analysis"]), thickness=float(gd["thickness"]))
    for tag, en in (("all", None), ("masked", gd[f"{name}/mask"])):
        K = port.Csr(g["rows_size"], g["cols"])
        st, F = port.assemble(K, m, float(g["E"]), float(g["nu"]), enabled=en, gp_damage=gd[f"{name}/damage"], **kw)
        assert st == 0
        assert np.array_equal(K.vals, gd[f"{name}/{tag}/K"]) and np.array_equal(F, gd[f"{name}/{tag}/F"])
    # zero damage is pipeline_assemble_system bit for bit (D * 1.0)
    K0, K1 = port.Csr(g["rows_size"], g["cols"]), port.Csr(g["rows_size"], g["cols"])
    st, F0 = port.assemble(K0, m, float(g["E"]), float(g["nu"]), gp_damage=np.zeros_like(gd[f"{name}/damage"]), **kw)
    st, F1 = port.assemble(K1, m, float(g["E"]), float(g["nu"]), **kw)
    assert np.array_equal(K0.vals, K1.vals) and np.array_equal(F0, F1)


def test_reference_known_answers():
    """The two asserts of the reference's own FEM suite (utest/.../static_elasticity2D.c:118,136)."""
    g = golden("beam_cantilever_trg1000")
    u = g["x"].reshape(-1, 2)
    assert abs(np.sqrt((u ** 2).sum(axis=1)).max() - 1.00701e-1) < 1e-6
    assert int(g["pcg_iters"]) == 801          # SURVEY.md §6 probe

    g = golden("plate_with_hole_trg1000")
    m = mesh_of(g)
    s = g["stress"].reshape(-1, 3)
    tri = m.adj.reshape(-1, 3)
    cen = m.nod.reshape(-1, 2)[tri].mean(axis=1)         # the single Gauss point of a linear triangle
    import ctypes as C
    ana = np.zeros((tri.shape[0], 3))
    for k, (x, y) in enumerate(cen):
        buf = (C.c_double * 3)()
        port.lib().nbo_kirsch_stress(x, y, buf)
        ana[k] = buf[:]
    vm = lambda t: np.sqrt(t[:, 0] ** 2 + t[:, 1] ** 2 - t[:, 0] * t[:, 1] + 3 * t[:, 2] ** 2)  # noqa: E731
    assert np.abs(1.0 - vm(s) / vm(ana)).mean() < 9.7e-3


def test_port_reproduces_reference_laplacian():
    g = golden("lap9_48")
    n = int(g["n"])
    rs, cols = port.sparse_pattern(*meshgen.laplacian9_graph(n), 1)
    assert np.array_equal(rs, g["rows_size"]) and np.array_equal(cols, g["cols"])
    A = port.Csr(rs, cols, g["vals"])
    b, tol = g["b"], float(g["tol"])
    assert np.array_equal(meshgen.uniform_rhs(n * n, seed=12345), b)
    st, x, it, res = A.pcg_jacobi(b, tol=tol)
    assert (st, it, res) == (int(g["pcg_status"]), int(g["pcg_iters"]), float(g["pcg_res"]))
    assert np.array_equal(x, g["x"])
    st, x, it, res = A.pcg_jacobi(b, tol=0.0, max_iter=25)
    assert (st, it) == (1, 25) and np.array_equal(x, g["x_cap"]) and res == float(g["cap_res"])
    st, x, it, res = A.pcg_jacobi(b, x0=g["x0"], tol=tol)
    assert it == int(g["warm_iters"]) and np.array_equal(x, g["x_warm"])
    st, x, it, res = A.cg(b, tol=tol)
    assert it == int(g["cg_iters"]) and np.array_equal(x, g["x_cg"])
    assert np.array_equal(A.spmv(b), g["spmv_b"])


def test_stale_residual_stopping_rule():
    """cg_precond_jacobi.c:45,84: the loop test and tolerance_reached see the residual of the iterate
    BEFORE the last update, so a solve runs one iteration past first convergence."""
    g = golden("lap9_48")
    A = port.Csr(g["rows_size"], g["cols"], g["vals"])
    b, tol = g["b"], float(g["tol"])
    st, x, it, res = A.pcg_jacobi(b, tol=tol)
    true_res_prev = None
    st2, x_prev, it2, _ = A.pcg_jacobi(b, tol=0.0, max_iter=it - 1)
    true_res_prev = np.linalg.norm(A.spmv(x_prev) - b)
    assert it2 == it - 1 and abs(true_res_prev - res) <= 1e-6 * res and res <= tol


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnbots_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("kind", [0, 1])
def test_port_matches_live_reference(kind):
    rng = np.random.default_rng(3 + kind)
    m = meshgen.structured_mesh(17, 9, 2.0, 1.0, kind=kind)
    m.nod += (rng.random(m.nod.size) - 0.5) * 0.02
    rm = ref.RefMesh.from_arrays(m)
    ga, gb = rm.graph()
    pa, pb = port.graph_nodes_by_elems(m)
    assert np.array_equal(ga, pa) and np.array_equal(gb, pb)
    K = ref.RefSparse.from_mesh(rm)
    rs, cols, _ = K.export()
    prs, pcols = port.pattern_from_mesh(m)
    assert np.array_equal(rs, prs) and np.array_equal(cols, pcols)
    en = (rng.random(m.n_elems) > 0.3).astype(np.uint8)
    kw = dict(density=1.5, self_weight=True, gravity=(0.2, -3.0), analysis=2, thickness=0.3, enabled=en)
    st, F = ref.assemble(K, rm, kind, 5.0, 0.2, **kw)
    PK = port.Csr(prs, pcols)
    st2, F2 = port.assemble(PK, m, 5.0, 0.2, **kw)
    assert st == st2 == 0 and np.array_equal(K.export()[2], PK.vals) and np.array_equal(F, F2)
    recs = [("dirichlet", "sgm", 3, (1, 1), (0.0, 0.0)), ("dirichlet", "sgm", 0, (0, 1), (0.0, 0.001)),
            ("neumann", "sgm", 1, (1, 1), (1.0, -2.0)), ("neumann", "vtx", 2, (1, 1), (0.5, 0.5)),
            ("neumann", "sgm", 2, (1, 1), (0, 0), 1)]
    bc = ref.RefBcond()
    for r in recs:
        bc.push_kirsch(r[2], r[5] - 1) if len(r) > 5 else bc.push(*r)
    ref.set_bconditions(rm, K, F, bc)
    port.set_bconditions(m, PK, F2, recs)
    assert np.array_equal(K.export()[2], PK.vals) and np.array_equal(F, F2)
    r1, r2 = K.pcg_jacobi(F, tol=1e-9), PK.pcg_jacobi(F2, tol=1e-9)
    assert r1[0] == r2[0] and r1[2] == r2[2] and np.array_equal(r1[1], r2[1])
    # the lumped mass vector (M != NULL) and the damage driver's loop on the same perturbed mesh
    K3 = ref.RefSparse.from_mesh(rm)
    st, F3, M3 = ref.assemble_with_mass(K3, rm, kind, 5.0, 0.2, **kw)
    st2, M4 = port.lumped_mass(m, kw["density"], kw["thickness"], en)
    assert st == st2 == 0 and np.array_equal(M3, M4)
    dmg = rng.random(m.n_elems * (4 if kind else 1)) * 0.99
    F5 = ref.assemble_damage(K3, rm, kind, 5.0, 0.2, dmg, **kw)
    PK3 = port.Csr(prs, pcols)
    st2, F6 = port.assemble(PK3, m, 5.0, 0.2, gp_damage=dmg, **kw)
    assert st2 == 0 and np.array_equal(K3.export()[2], PK3.vals) and np.array_equal(F5, F6)
    # a distorted element (clockwise) is reported by both (pipeline.c:158-159)
    m2 = meshgen.structured_mesh(4, 3, 4.0, 3.0, kind=kind)
    npe = m2.npe
    m2.adj[npe * 5:npe * 6] = m2.adj[npe * 5:npe * 6][::-1].copy()
    rm2 = ref.RefMesh.from_arrays(m2)
    K2 = ref.RefSparse.from_mesh(rm2)
    st, _ = ref.assemble(K2, rm2, kind, 1.0, 0.3)
    PK2 = port.Csr(*port.pattern_from_mesh(m2))
    st2, _ = port.assemble(PK2, m2, 1.0, 0.3)
    assert st == st2 == 1
    assert np.array_equal(K2.export()[2], PK2.vals)      # both hold the elements before the bad one
    # Gauss-point -> node projection: random field with 1, 3 and 5 components; distorted mesh -> status 1
    ngp = 4 if kind else 1
    for n_comp in (1, 3, 5):
        gp = rng.standard_normal(m.n_elems * ngp * n_comp)
        (s1, n1), (s2, n2) = ref.gp_to_nodes(rm, kind, m.n_nod, n_comp, gp), port.gp_to_nodes(m, n_comp, gp)
        assert s1 == s2 == 0 and np.array_equal(n1, n2)
    gp = rng.standard_normal(m2.n_elems * ngp * 3)
    assert ref.gp_to_nodes(rm2, kind, m2.n_nod, 3, gp)[0] == port.gp_to_nodes(m2, 3, gp)[0] == 1
    t = rng.standard_normal(300) * 1e3
    assert np.array_equal(ref.vm_stress(t), port.vm_stress(t))
    assert np.array_equal(ref.main_stress(t), port.main_stress(t))
