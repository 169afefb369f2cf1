"""Diagnostic (GPU box): how far default-mode PCG iterates are from the reference's, per fixture."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nbots_b200 import api, capi
from oracle import port
from util import FEM_CASES, golden, rel_l2

capi.check(capi.lib().nbgpu_init(0))
for name in FEM_CASES:
    g = golden(name)
    A = api.Matrix.from_csr(g["rows_size"], g["cols"], g["K_post"])
    P = port.Csr(g["rows_size"], g["cols"], g["K_post"])
    b = g["F_post"]
    st, x, it, res = A.pcg_jacobi_host(b, tol=float(g["tol"]))
    print(f"{name}: abs tol {float(g['tol']):.0e}: gpu it={it} ref it={int(g['pcg_iters'])} rel_l2={rel_l2(x, g['x']):.2e}")
    tol = 1e-8 * np.linalg.norm(b)
    ost, ox, oit, ores = P.pcg_jacobi(b, tol=tol)
    st, x, it, res = A.pcg_jacobi_host(b, tol=tol)
    print(f"   rel tol 1e-8: gpu it={it} ref it={oit} rel_l2={rel_l2(x, ox):.2e}")
    st, x2, it2, res2 = A.pcg_jacobi_host(b, tol=0.0, max_iter=oit)
    print(f"   same iteration count {oit}: rel_l2={rel_l2(x2, ox):.2e}  res gpu={res2:.3e} ref={ores:.3e}")
    xs = np.linalg.solve  # noqa
    r_gpu = np.linalg.norm(P.spmv(x) - b) / np.linalg.norm(b)
    r_ref = np.linalg.norm(P.spmv(ox) - b) / np.linalg.norm(b)
    print(f"   true relative residuals: gpu={r_gpu:.2e} ref={r_ref:.2e}")
