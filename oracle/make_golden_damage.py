"""Writes tests/golden/damage_assembly.npz: K and F as the REFERENCE's damage driver assembles them
(DMG_pipeline_assemble_system, static_damage2D.c:474-569, reached through oracle/ref_harness.c) for a random
per-Gauss-point damage field, on one triangle and one quad fixture mesh, with and without an enabled mask.  Needs
oracle/_ref.  TEST INFRASTRUCTURE, not product code."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref  # noqa: E402
from util import golden, mesh_of  # noqa: E402

CASES = ["beam_cantilever_trg1000", "quad_void_selfweight_24x8"]
DENSITY, THICKNESS, GRAVITY = 2.5, 0.7, (0.1, -9.8)
out = {"density": DENSITY, "thickness": THICKNESS, "gravity": np.array(GRAVITY)}
for name in CASES:
    g = golden(name)
    m = mesh_of(g)
    rm = ref.RefMesh.from_arrays(m)
    K = ref.RefSparse.from_mesh(rm)
    rng = np.random.default_rng(5)
    ngp = 4 if m.kind else 1
    dmg = rng.random(m.n_elems * ngp) * 0.9
    dmg[rng.random(dmg.size) < 0.3] = 0.0                  # undamaged points too
    mask = (rng.random(m.n_elems) > 0.2).astype(np.uint8)
    out[f"{name}/damage"], out[f"{name}/mask"] = dmg, mask
    for tag, en in (("all", None), ("masked", mask)):
        F = ref.assemble_damage(K, rm, m.kind, float(g["E"]), float(g["nu"]), dmg, density=DENSITY, self_weight=True,
                                gravity=GRAVITY, analysis=int(g["analysis"]), thickness=THICKNESS, enabled=en)
        out[f"{name}/{tag}/K"], out[f"{name}/{tag}/F"] = K.export()[2].copy(), F
    # no damage array: the loop must reduce to pipeline_assemble_system
    F0 = ref.assemble_damage(K, rm, m.kind, float(g["E"]), float(g["nu"]), None, density=DENSITY, self_weight=True,
                             gravity=GRAVITY, analysis=int(g["analysis"]), thickness=THICKNESS)
    K0 = K.export()[2].copy()
    st, F1 = ref.assemble(K, rm, m.kind, float(g["E"]), float(g["nu"]), density=DENSITY, self_weight=True,
                          gravity=GRAVITY, analysis=int(g["analysis"]), thickness=THICKNESS)
    assert st == 0 and np.array_equal(K0, K.export()[2]) and np.array_equal(F0, F1)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "damage_assembly.npz"), **out)
print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
