"""GPU: external problems enter through the reference's on-disk formats.  A system written by the
REFERENCE (tests/golden/io/lap9_5.mat: nb_sparse_save_mat4 + nb_mat4_save_vec) is read, solved on the device by
scripts/solve_file.py and the solution appended to the file; a mesh written by the reference's VTK writer is read back
and assembled.  Checked against the oracle."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

from nbots_b200 import api, io
from oracle import port
from util import rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "io")


@pytest.mark.parametrize("solver", ["pcg", "cg"])
def test_solve_a_reference_written_mat4_system_end_to_end(nbgpu_lib, tmp_path, solver):
    path = str(tmp_path / "sys.mat")
    shutil.copy(os.path.join(GOLD, "lap9_5.mat"), path)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "solve_file.py"), path, "--solver", solver,
                          "--rel-tol", "1e-10", "--max-iter", "500"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    rec = io.load_mat4(path)
    rs, cols, vals = rec["A"]
    b, x = rec["b"], rec["x"]
    K = port.Csr(rs, cols, vals)
    tol = 1e-10 * float(np.linalg.norm(b))
    ost, ox, oit, ores = (K.pcg_jacobi if solver == "pcg" else K.cg)(b, tol=tol, max_iter=500)
    assert ost == 0 and rel_l2(x, ox) <= 1e-10
    assert np.linalg.norm(K.spmv(x) - b) <= 2 * tol
    it = int(out.stdout.split("iterations=")[1].split()[0])
    assert abs(it - oit) <= max(1, int(np.ceil(0.02 * oit)))


@pytest.mark.parametrize("name", ["grid_trg.vtk", "grid_quad.vtk"])
def test_assemble_on_a_reference_written_vtk_mesh(nbgpu_lib, name):
    m = io.load_vtk(os.path.join(GOLD, name))
    rs, cols = api.pattern_from_mesh(m)
    ors, ocols = port.pattern_from_mesh(m)
    assert np.array_equal(rs, ors) and np.array_equal(cols, ocols)
    K = api.Matrix.from_csr(rs, cols)
    mesh = api.Mesh(m)
    d_F = api.DeviceBuffer.zeros(K.N)
    st, _ = mesh.assemble(K, d_F, 2.1e11, 0.3, density=7850.0, self_weight=True, gravity=(0.0, -9.81), thickness=0.1)
    oK = port.Csr(ors, ocols)
    ost, oF = port.assemble(oK, m, 2.1e11, 0.3, density=7850.0, self_weight=True, gravity=(0.0, -9.81), thickness=0.1)
    assert st == ost == 0
    assert np.array_equal(K.values_csr(), oK.vals) and np.array_equal(d_F.to_host(), oF)
