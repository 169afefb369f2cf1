#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden FROM THE REFERENCE ITSELF.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference
for the problem files and oracle/_ref/libnbots_ref.so = the unmodified reference
compiled by oracle/Makefile).  Every array written here is an output of the
reference library, obtained through oracle/ref_harness.c:

  *_trg1000.npz  the reference's own FEM test problems
                 (utest/sources/nb/pde_bot/static_elasticity2D_inputs/*.txt),
                 meshed by the reference's mesher with MAX_VTX = 1000 exactly as
                 utest/.../finite_element/solid_mechanics/static_elasticity2D.c
                 does; they carry the known answers that test asserts
                 (max|u| = 1.00701e-1 +- 1e-6; mean von-Mises error < 9.7e-3).
  quad_*.npz     synthetic structured-quad cantilevers
  lap9_*.npz     9-point grid Laplacian through nb_sparse_create(graph,NULL,1)

Usage:  python oracle/make_golden.py        (rewrites tests/golden/*.npz)
"""
from __future__ import annotations

import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nbots_b200 import meshgen  # noqa: E402
from oracle import ref  # noqa: E402

INPUTS = "/root/reference/utest/sources/nb/pde_bot/static_elasticity2D_inputs"
OUT = os.path.join(ROOT, "tests", "golden")


def read_problem(path):
    """Tokenise a reference problem file ('#' starts a comment) -> dict."""
    toks = []
    with open(path) as f:
        for line in f:
            toks += line.split("#", 1)[0].split()
    it = iter(toks)
    nxt = lambda conv=float: conv(next(it))  # noqa: E731
    n = nxt(int)
    vertex = np.array([nxt() for _ in range(2 * n)])
    m = nxt(int)
    edge = np.array([nxt(int) for _ in range(2 * m)], dtype=np.uint32)
    h = nxt(int)
    holes = np.array([nxt() for _ in range(2 * h)])
    bcs = []
    # bcond_read.c:12-26 : dirichlet vtx, neumann vtx, dirichlet sgm, neumann sgm
    for kind, where in (("dirichlet", "vtx"), ("neumann", "vtx"), ("dirichlet", "sgm"), ("neumann", "sgm")):
        for _ in range(nxt(int)):
            ident = nxt(int)
            mask = [nxt(int), nxt(int)]
            val = [nxt() if mask[0] else 0.0, 0.0]
            val[1] = nxt() if mask[1] else 0.0
            bcs.append((kind, where, ident, tuple(mask), tuple(val)))
    nu = nxt(); E = nxt(); nxt(); nxt(); nxt()
    analysis = nxt(int)
    thickness = nxt()
    return dict(vertex=vertex, edge=edge, holes=holes, bcs=bcs, nu=nu, E=E, analysis=analysis, thickness=thickness)


def bcs_to_arrays(bcs):
    """(kind, where, id, mask, val[, fn]) records -> flat arrays for the npz."""
    n = len(bcs)
    out = dict(bc_kind=np.zeros(n, np.int32), bc_where=np.zeros(n, np.int32), bc_id=np.zeros(n, np.uint32),
               bc_mask=np.zeros((n, 2), np.int32), bc_val=np.zeros((n, 2)), bc_fn=np.zeros(n, np.int32))
    for k, r in enumerate(bcs):
        out["bc_kind"][k] = 0 if r[0] == "dirichlet" else 1
        out["bc_where"][k] = 0 if r[1] == "vtx" else 1
        out["bc_id"][k] = r[2]
        out["bc_mask"][k] = r[3]
        out["bc_val"][k] = r[4]
        out["bc_fn"][k] = r[5] if len(r) > 5 else 0
    return out


def fem_case(name, m, rm, bcs, E, nu, analysis=0, thickness=1.0, density=0.0, self_weight=False,
             gravity=(0.0, 0.0), enabled=None, tol=1e-8, extra=None):
    """Run the reference pipeline step by step and dump every intermediate."""
    kind = m.kind
    ngp = 4 if kind else 1
    K = ref.RefSparse.from_mesh(rm)
    rows_size, cols, _ = K.export()
    st, F_pre = ref.assemble(K, rm, kind, E, nu, density=density, self_weight=self_weight, gravity=gravity,
                             analysis=analysis, thickness=thickness, enabled=enabled)
    assert st == 0
    K_pre = K.export()[2]
    bc = ref.RefBcond()
    for r in bcs:
        if len(r) > 5 and r[5]:
            bc.push_kirsch(r[2], r[5] - 1)
        else:
            bc.push(*r[:5])
    F_post = F_pre.copy()
    ref.set_bconditions(rm, K, F_post, bc)
    K_post = K.export()[2]
    # the FEM driver's solver call: x0 = 0, max_iter = N, abs tol (static_elasticity2D.c:83-97)
    st, x, iters, res = K.pcg_jacobi(F_post, tol=tol)
    st_cg, x_cg, iters_cg, res_cg = K.cg(F_post, tol=tol)
    # and the whole driver in one call, as the reference's test uses it
    st_drv, disp, strain = ref.fem_static(rm, kind, E, nu, bc, density=density, self_weight=self_weight,
                                          gravity=gravity, analysis=analysis, thickness=thickness, enabled=enabled,
                                          n_nod=m.n_nod, n_elems=m.n_elems, n_gp=ngp)
    assert st_drv == 0
    if tol == 1e-8:
        assert np.array_equal(disp, x), "driver and step-by-step solve disagree"
    strain_x = ref.compute_strain(rm, kind, x, m.n_elems, ngp)
    stress = ref.stress_from_strain(m.n_elems, kind, E, nu, analysis, strain_x, enabled)
    # post-processing next to the solve: nodal projection (gaussp_to_nodes.c), von Mises, main stresses
    st_nod, stress_nod = ref.gp_to_nodes(rm, kind, m.n_nod, 3, stress)
    assert st_nod == 0
    vm = ref.vm_stress(stress)
    main = ref.main_stress(stress)
    y = K.spmv(x)
    data = dict(stress_nod=stress_nod, vm=vm, main_stress=main, kind=kind, nod=m.nod, edg=m.edg, adj=m.adj, vtx=m.vtx, sgm_sizes=m.sgm_sizes, sgm_nodes=m.sgm_nodes,
                E=E, nu=nu, analysis=analysis, thickness=thickness, density=density, self_weight=int(self_weight),
                gravity=np.array(gravity), tol=tol, rows_size=rows_size, cols=cols, K_pre=K_pre, F_pre=F_pre,
                K_post=K_post, F_post=F_post, x=x, pcg_status=st, pcg_iters=iters, pcg_res=res, x_cg=x_cg,
                cg_status=st_cg, cg_iters=iters_cg, cg_res=res_cg, strain=strain_x, stress=stress, spmv_x=y,
                D=ref.constitutive(E, nu, analysis))
    if enabled is not None:
        data["enabled"] = np.asarray(enabled, dtype=np.uint8)
    data.update(bcs_to_arrays(bcs))
    if extra:
        data.update(extra)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **data)
    md = float(np.sqrt((x.reshape(-1, 2) ** 2).sum(axis=1)).max())
    print(f"{name}: N={K.N} nnz={K.nnz} pcg={iters} it (status {st}, res {res:.3e}) cg={iters_cg} it "
          f"max|u|={md:.9e}")
    K.close(); bc.close()
    return md


def main():
    os.makedirs(OUT, exist_ok=True)
    # ---- the reference's own FEM test problems (config 1) ------------------------
    p = read_problem(os.path.join(INPUTS, "beam_cantilever.txt"))
    rm = ref.RefMesh.from_model(p["vertex"], p["edge"], p["holes"], 1000)
    m = rm.export()
    md = fem_case("beam_cantilever_trg1000", m, rm, p["bcs"], p["E"], p["nu"], p["analysis"], p["thickness"])
    assert abs(md - 1.00701e-1) < 1e-6, "reference known answer (utest static_elasticity2D.c:118)"
    rm.close()

    p = read_problem(os.path.join(INPUTS, "plate_with_hole.txt"))
    rm = ref.RefMesh.from_model(p["vertex"], p["edge"], p["holes"], 1000)
    m = rm.export()
    # utest static_elasticity2D.c:208-220: Kirsch tractions on segments 11 (+x face) and 12 (+y face)
    bcs = p["bcs"] + [("neumann", "sgm", 11, (1, 1), (0, 0), 1), ("neumann", "sgm", 12, (1, 1), (0, 0), 2)]
    fem_case("plate_with_hole_trg1000", m, rm, bcs, p["E"], p["nu"], p["analysis"], p["thickness"])
    rm.close()

    # ---- synthetic structured quads (configs 2/4 at test size) ---------------------
    m = meshgen.structured_mesh(64, 16, 4.0, 1.0, kind=1)
    rm = ref.RefMesh.from_arrays(m)
    bcs = [("dirichlet", "sgm", 3, (1, 1), (0, 0)), ("neumann", "sgm", 1, (1, 1), (0, -1))]
    fem_case("quad_cantilever_64x16", m, rm, bcs, 1.0, 0.3, analysis=0, thickness=1.0)
    rm.close()

    # "plane strain" flag (config 4): the reference still uses the plane-stress D (formulas.c:38-45)
    rng = np.random.default_rng(7)
    m = meshgen.structured_mesh(24, 8, 3.0, 1.0, kind=1)
    m.nod += (rng.random(m.nod.size) - 0.5) * 0.03          # non-rectangular quads
    rm = ref.RefMesh.from_arrays(m)
    enabled = (rng.random(m.n_elems) > 0.25).astype(np.uint8)  # void elements (pipeline.c:93-98)
    bcs = [("dirichlet", "sgm", 3, (1, 1), (0, 0.002)), ("dirichlet", "vtx", 1, (0, 1), (0, -0.01)),
           ("neumann", "sgm", 2, (1, 1), (0.3, -1.0)), ("neumann", "vtx", 2, (1, 0), (0.2, 0))]
    fem_case("quad_void_selfweight_24x8", m, rm, bcs, 3.0, 0.25, analysis=1, thickness=0.5, density=2.0,
             self_weight=True, gravity=(0.05, -9.8), enabled=enabled, tol=1e-9)
    rm.close()

    # ---- 9-point Laplacian (config 3 at test size) ---------------------------------
    n = 48
    n_adj, adj = meshgen.laplacian9_graph(n)
    A = ref.RefSparse.from_graph(n_adj, adj, 1)
    rows_size, cols, _ = A.export()
    rs2, cols2, vals = meshgen.laplacian9_csr(n)
    assert np.array_equal(rows_size, rs2) and np.array_equal(cols, cols2)
    A.set_values(vals)
    b = meshgen.uniform_rhs(n * n, seed=12345)
    tol = 1e-8 * float(np.linalg.norm(b))
    st, x, it, res = A.pcg_jacobi(b, tol=tol)
    st2, x2, it2, res2 = A.cg(b, tol=tol)
    st3, x3, it3, res3 = A.pcg_jacobi(b, tol=0.0, max_iter=25)      # max_iter exit, status 1
    x0 = meshgen.uniform_rhs(n * n, seed=99)
    st4, x4, it4, res4 = A.pcg_jacobi(b, x0=x0, tol=tol)           # initial guess honoured
    y = A.spmv(b)
    np.savez_compressed(os.path.join(OUT, f"lap9_{n}.npz"), n=n, rows_size=rows_size, cols=cols, vals=vals, b=b,
                        tol=tol, x=x, pcg_status=st, pcg_iters=it, pcg_res=res, x_cg=x2, cg_status=st2,
                        cg_iters=it2, cg_res=res2, x_cap=x3, cap_status=st3, cap_iters=it3, cap_res=res3, x0=x0,
                        x_warm=x4, warm_status=st4, warm_iters=it4, warm_res=res4, spmv_b=y)
    print(f"lap9_{n}: N={A.N} nnz={A.nnz} pcg={it} cg={it2} cap=({st3},{it3}) warm={it4}")
    A.close()


if __name__ == "__main__":
    main()
