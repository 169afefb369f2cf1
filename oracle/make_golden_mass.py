"""Writes tests/golden/lumped_mass.npz: the M vector the REFERENCE's pipeline_assemble_system returns (M != NULL) on
the meshes of the four FEM fixtures, with and without an enabled mask, plus F of the same call (must equal the
fixture's own F).  Needs oracle/_ref.  TEST INFRASTRUCTURE, not product code."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref  # noqa: E402
from util import FEM_CASES, golden, mesh_of  # noqa: E402

DENSITY, THICKNESS = 7.85, 0.35          # the fixtures' own densities are mostly 0
out = {"density": DENSITY, "thickness": THICKNESS}
for name in FEM_CASES:
    g = golden(name)
    m = mesh_of(g)
    rm = ref.RefMesh.from_arrays(m)
    K = ref.RefSparse.from_mesh(rm)
    mask = (np.random.default_rng(11).random(m.n_elems) > 0.25).astype(np.uint8)
    for tag, en in (("all", None), ("masked", mask)):
        st, F, M = ref.assemble_with_mass(K, rm, m.kind, float(g["E"]), float(g["nu"]), density=DENSITY,
                                          self_weight=True, gravity=(0.3, -9.81), analysis=int(g["analysis"]),
                                          thickness=THICKNESS, enabled=en)
        assert st == 0
        st0, F0 = ref.assemble(K, rm, m.kind, float(g["E"]), float(g["nu"]), density=DENSITY, self_weight=True,
                               gravity=(0.3, -9.81), analysis=int(g["analysis"]), thickness=THICKNESS, enabled=en)
        assert st0 == 0 and np.array_equal(F, F0)        # asking for M changes nothing else
        out[f"{name}/{tag}/M"] = M
    out[f"{name}/mask"] = mask
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lumped_mass.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
