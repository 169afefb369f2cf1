#!/usr/bin/env python
"""CPU study: iteration counts / field agreement of rearranged Jacobi-PCG recurrences.

  std   the reference's recurrence (cg_precond_jacobi.c:45-76): w = A p, two dependent reductions
  wrec  w-recurrence: s = A q, p = -q + b p, w = -s + b w, p.w summed directly (two reductions)
  cgr   Chronopoulos-Gear: s = A q with g.q, q.s, g.g reduced together; a = gq / (qs - b gq / a_old)

All use the reference's stale-residual gate.  Run on the quad cantilever at a few sizes; prints
iterations and the relative L2 distance of x to the std run.  Test infrastructure (uses oracle/).
"""
import sys, os
import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nbots_b200 import meshgen
from oracle import port


def system(nx, ny):
    m = meshgen.structured_mesh(nx, ny, 2.0, 1.0, kind=1)
    rs, cols = port.pattern_from_mesh(m)
    K = port.Csr(rs, cols)
    st, F = port.assemble(K, m, 1.0, 0.3, thickness=1.0)
    bcs = [("dirichlet", "sgm", 3, (1, 1), (0.0, 0.0)), ("neumann", "sgm", 1, (1, 1), (0.0, -1.0))]
    port.set_bconditions(m, K, F, bcs)
    A = sp.csr_matrix((K.vals, K.cols.astype(np.int64), K.row_ptr.astype(np.int64)), shape=(K.N, K.N))
    return A, F


def std(A, b, tol, max_iter):
    d = A.diagonal(); x = np.zeros_like(b)
    g = A @ x - b; q = g / d; p = -q
    gg = g @ g; gq = g @ q
    k = 0
    while gg > tol * tol and k < max_iter:
        w = A @ p; pw = p @ w
        gg = g @ g; gq = g @ q
        a = gq / pw
        x += a * p; g += a * w; q = g / d
        gq2 = g @ q
        p = -q + (gq2 / gq) * p
        k += 1
    return x, k


def wrec(A, b, tol, max_iter):
    d = A.diagonal(); x = np.zeros_like(b)
    g = A @ x - b; q = g / d
    p = np.zeros_like(b); w = np.zeros_like(b)
    gg = g @ g; gq = g @ q; beta = 0.0
    gg_gate = gg
    k = 0
    while gg_gate > tol * tol and k < max_iter:
        gg_gate = gg
        s = A @ q
        p = -q + beta * p
        w = -s + beta * w
        pw = p @ w
        a = gq / pw
        x += a * p; g += a * w; q = g / d
        gq2 = g @ q; gg = g @ g
        beta = gq2 / gq; gq = gq2
        k += 1
    return x, k


def cgr(A, b, tol, max_iter):
    d = A.diagonal(); x = np.zeros_like(b)
    g = A @ x - b; q = g / d
    p = np.zeros_like(b); w = np.zeros_like(b)
    gq_old = 1.0; a_old = 1.0
    gg_gate = g @ g
    k = 0
    while gg_gate > tol * tol and k < max_iter:
        s = A @ q
        gq = g @ q; qs = q @ s; gg = g @ g
        gg_gate = gg
        if k == 0:
            beta = 0.0; a = gq / qs
        else:
            beta = gq / gq_old
            a = gq / (qs - beta * gq / a_old)
        p = -q + beta * p
        w = -s + beta * w
        x += a * p; g += a * w; q = g / d
        gq_old = gq; a_old = a
        k += 1
    return x, k


if __name__ == "__main__":
    for nx, ny in [(64, 16), (256, 64), (512, 128), (1000, 500)][: int(sys.argv[1]) if len(sys.argv) > 1 else 3]:
        A, b = system(nx, ny)
        tol = 1e-8 * np.linalg.norm(b)
        x0, k0 = std(A, b, tol, b.size)
        r0 = np.linalg.norm(A @ x0 - b) / np.linalg.norm(b)
        print(f"{nx}x{ny} N={b.size}: std {k0} its, true rel residual {r0:.2e}")
        for name, fn in (("wrec", wrec), ("cgr", cgr)):
            x, k = fn(A, b, tol, b.size)
            r = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
            print(f"   {name}: {k} its ({(k - k0) / k0 * 100:+.2f} %), rel_l2(x, std) = "
                  f"{np.linalg.norm(x - x0) / np.linalg.norm(x0):.2e}, true rel residual {r:.2e}")
