"""Writes tests/golden/io/*: small files produced by the REFERENCE's own writers (nb_sparse_save,
nb_sparse_save_mat4, nb_mat4_save_vec, nb_mesh2D_save_vtk, nb_mesh2D_save_nbt), for tests/test_io.py.  Needs oracle/_ref.
TEST INFRASTRUCTURE, not product code."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nbots_b200 import meshgen  # noqa: E402
from oracle import ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "io")
os.makedirs(OUT, exist_ok=True)
os.chdir(OUT)                                            # the VTK header embeds the path as given
for f in os.listdir("."):
    os.remove(f)

rs, cols, vals = meshgen.laplacian9_csr(5)
vals = vals * np.linspace(0.5, 1.5, vals.size) * 1.2345678901234e-3      # same as tests/test_io.py::small_system
A = ref.RefSparse.from_csr(rs, cols, vals)
b = meshgen.uniform_rhs(rs.size, seed=7)
L = ref.lib()
L.refh_sparse_save(A.h, b"lap9_5.coo.txt")
L.refh_sparse_save_mat4(A.h, b"lap9_5.mat", b"A")
L.refh_mat4_save_vec(b"lap9_5.mat", b"b", b.ctypes.data_as(ref.f64p), b.size)
for kind in (0, 1):
    m = meshgen.structured_mesh(4, 3, 2.0, 1.0, kind=kind)
    rm = ref.RefMesh.from_arrays(m)
    name = "grid_%s.vtk" % ("quad" if kind else "trg")
    assert L.refh_mesh_save_vtk(rm.h, name.encode()) == 0
    assert L.refh_mesh_save_nbt(rm.h, name.replace(".vtk", ".nbt").encode()) == 0
print(sorted((f, os.path.getsize(f)) for f in os.listdir(".")))
