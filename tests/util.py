"""Shared helpers of the test-suite (golden fixtures, norms, BC records)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from nbots_b200 import capi, meshgen

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FEM_CASES = ["beam_cantilever_trg1000", "plate_with_hole_trg1000", "quad_cantilever_64x16",
             "quad_void_selfweight_24x8"]


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def mesh_of(g) -> meshgen.Mesh2D:
    return meshgen.Mesh2D(kind=int(g["kind"]), nod=g["nod"].copy(), edg=g["edg"].copy(), adj=g["adj"].copy(),
                          vtx=g["vtx"].copy(), sgm_sizes=g["sgm_sizes"].copy(), sgm_nodes=g["sgm_nodes"].copy())


def bc_records(g):
    """-> list of (kind, where, id, mask, val, fn) tuples as oracle.port.make_bcs takes them."""
    recs = []
    for k in range(g["bc_kind"].size):
        recs.append(("dirichlet" if g["bc_kind"][k] == 0 else "neumann", "vtx" if g["bc_where"][k] == 0 else "sgm",
                     int(g["bc_id"][k]), tuple(int(v) for v in g["bc_mask"][k]),
                     tuple(float(v) for v in g["bc_val"][k]), int(g["bc_fn"][k])))
    return recs


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / nb) if nb > 0 else float(np.linalg.norm(a))


# ---- function-valued conditions for the product's nbgpu_bcond_t -----------------------------
# The Kirsch tractions of the reference's plate-with-hole test are evaluated by the ORACLE
# (same expression as oracle/ref_harness.c), handed to the product as a C callback.
_keep = []


def kirsch_callback(which):
    from oracle import port
    L = port.lib()

    def fn(x, t, out):
        s = (C.c_double * 3)()
        L.nbo_kirsch_stress(x[0], x[1], s)
        if which == 1:
            out[0], out[1] = s[0], s[2]
        else:
            out[0], out[1] = s[2], s[1]
    cb = capi.BCFUNC(fn)
    _keep.append(cb)
    return cb


def product_bcs(recs):
    arr = (capi.BCond * max(1, len(recs)))()
    for k, r in enumerate(recs):
        arr[k].kind = 0 if r[0] == "dirichlet" else 1
        arr[k].where = 0 if r[1] == "vtx" else 1
        arr[k].id = r[2]
        arr[k].mask[0], arr[k].mask[1] = r[3]
        arr[k].val[0], arr[k].val[1] = r[4]
        fn = r[5] if len(r) > 5 else 0
        arr[k].fval = C.cast(kirsch_callback(fn), C.c_void_p).value if fn else None
    return arr, len(recs)


def flatten_bcs(m, recs, factor=1.0):
    """Product host logic: nbgpu_bcond_flatten -> (neu_dof, neu_add, dir_dof, dir_val)."""
    L = capi.lib()
    L.nbgpu_bcond_flatten.restype = C.c_int
    arr, n = product_bcs(recs)
    nn = C.c_uint32(0); nd = C.c_uint32(0)
    u32p, f64p = capi.u32p, capi.f64p
    args = (m.nod.ctypes.data_as(f64p), m.vtx.ctypes.data_as(u32p), C.c_uint32(m.sgm_sizes.size),
            m.sgm_sizes.ctypes.data_as(u32p), m.sgm_nodes.ctypes.data_as(u32p), C.c_uint32(n), arr,
            C.c_double(factor))
    capi.check(L.nbgpu_bcond_flatten(*args, C.byref(nn), None, None, C.byref(nd), None, None))
    neu_dof = np.zeros(nn.value, np.uint32); neu_add = np.zeros(nn.value)
    dir_dof = np.zeros(nd.value, np.uint32); dir_val = np.zeros(nd.value)
    capi.check(L.nbgpu_bcond_flatten(*args, C.byref(nn), neu_dof.ctypes.data_as(u32p), neu_add.ctypes.data_as(f64p),
                                     C.byref(nd), dir_dof.ctypes.data_as(u32p), dir_val.ctypes.data_as(f64p)))
    return neu_dof, neu_add, dir_dof, dir_val
