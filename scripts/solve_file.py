"""Solve a system stored in a MATLAB v4 file (as written by the reference's nb_sparse_save_mat4 /
nb_mat4_save_vec, or by nbots_b200.io) on the GPU and append the solution to the file.

    python scripts/solve_file.py system.mat [--matrix A] [--rhs b] [--solution x] [--solver pcg|cg]
                                            [--rel-tol 1e-8] [--max-iter N]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbots_b200 import api, capi, io  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--matrix", default="A")
    ap.add_argument("--rhs", default="b")
    ap.add_argument("--solution", default="x")
    ap.add_argument("--solver", choices=("pcg", "cg"), default="pcg")
    ap.add_argument("--rel-tol", type=float, default=1e-8)
    ap.add_argument("--max-iter", type=int, default=0)
    a = ap.parse_args()
    rec = io.load_mat4(a.path)
    rs, cols, vals = rec[a.matrix]
    b = rec[a.rhs]
    capi.check(capi.lib().nbgpu_init(0))
    A = api.Matrix.from_csr(rs, cols, vals)
    tol = a.rel_tol * float(np.linalg.norm(b))
    solve = A.pcg_jacobi_host if a.solver == "pcg" else A.cg_host
    st, x, it, res = solve(b, tol=tol, max_iter=a.max_iter or None)
    true_res = float(np.linalg.norm(A.spmv_host(x) - b))
    io.save_mat4_vector(a.path, a.solution, x)
    print(f"N={A.N} nnz={A.nnz} status={st} iterations={it} residual={res:.3e} (true {true_res:.3e}, asked {tol:.3e})")
    return st


if __name__ == "__main__":
    sys.exit(main())
