"""On-disk formats (nbots_b200/io.py): the writers must produce the reference's bytes, the loaders must
bring back what was written.  Fixtures under tests/golden/io/ were written by the reference itself
(oracle/make_golden_io.py); with oracle/_ref present the comparison is repeated live."""
import ctypes as C
import os

import numpy as np
import pytest

from nbots_b200 import io, meshgen
from oracle import ref
from util import golden

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io")


def small_system():
    rs, cols, vals = meshgen.laplacian9_csr(5)
    vals = vals * np.linspace(0.5, 1.5, vals.size) * 1.2345678901234e-3
    return rs, cols, vals


def read(path):
    with open(path, "rb") as fp:
        return fp.read()


def test_coo_text_matches_reference_bytes(tmp_path):
    rs, cols, vals = small_system()
    p = tmp_path / "A.txt"
    io.save_coo_text(p, rs, cols, vals)
    assert read(p) == read(os.path.join(GOLD, "lap9_5.coo.txt"))
    rs2, cols2, vals2 = io.load_coo_text(p)
    assert np.array_equal(rs2, rs) and np.array_equal(cols2, cols)
    assert np.allclose(vals2, vals, rtol=1e-6, atol=0)             # "%e": 7 significant digits


def test_mat4_matches_reference_bytes_and_is_lossless(tmp_path):
    rs, cols, vals = small_system()
    b = meshgen.uniform_rhs(rs.size, seed=7)
    p = tmp_path / "sys.mat"
    io.save_mat4_sparse(p, "A", rs, cols, vals)
    io.save_mat4_vector(p, "b", b)
    assert read(p) == read(os.path.join(GOLD, "lap9_5.mat"))
    got = io.load_mat4(p)
    assert set(got) == {"A", "b"}
    assert all(np.array_equal(a, w) for a, w in zip(got["A"], (rs, cols, vals)))
    assert np.array_equal(got["b"], b)


@pytest.mark.parametrize("kind", [0, 1])
def test_vtk_matches_reference_bytes(tmp_path, kind):
    m = meshgen.structured_mesh(4, 3, 2.0, 1.0, kind=kind)
    name = "grid_%s.vtk" % ("quad" if kind else "trg")
    cwd = os.getcwd()
    os.chdir(tmp_path)                                             # the header embeds the path as given
    try:
        io.save_vtk(name, m)
        assert read(name) == read(os.path.join(GOLD, name))
        back = io.load_vtk(name)
    finally:
        os.chdir(cwd)
    assert back.kind == kind and np.array_equal(back.adj, m.adj)
    assert np.allclose(back.nod, m.nod, rtol=0, atol=1e-6)
    # element sides = the mesh edges of these grids (the quad diagonals are no edges)
    want = np.unique(np.sort(m.edg.reshape(-1, 2), axis=1), axis=0)
    assert np.array_equal(back.edg.reshape(-1, 2), want)


def test_loaders_reject_foreign_files(tmp_path):
    p = tmp_path / "x.txt"
    p.write_text("3 4\n0 0 1.0\n")
    with pytest.raises(ValueError):
        io.load_coo_text(p)
    p.write_text("2 2\n0 0 1.0\n0 0 2.0\n")
    with pytest.raises(ValueError):
        io.load_coo_text(p)                                        # duplicate entry
    p.write_text("# vtk DataFile Version 2.0\nsomething else\nASCII\nDATASET POLYDATA\n\n")
    with pytest.raises(ValueError):
        io.load_vtk(p)


def test_golden_system_round_trips_through_mat4(tmp_path):
    g = golden("quad_cantilever_64x16")
    p = tmp_path / "k.mat"
    io.save_mat4_sparse(p, "K", g["rows_size"], g["cols"], g["K_post"])
    io.save_mat4_vector(p, "F", g["F_post"])
    got = io.load_mat4(p)
    assert np.array_equal(got["K"][0], g["rows_size"]) and np.array_equal(got["K"][1], g["cols"])
    assert np.array_equal(got["K"][2], g["K_post"]) and np.array_equal(got["F"], g["F_post"])


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libnbots_ref.so not built (needs /root/reference)")
def test_writers_match_live_reference(tmp_path):
    rng = np.random.default_rng(2)
    m = meshgen.structured_mesh(9, 6, 3.0, 2.0, kind=0, diagonal_seed=4)
    rm = ref.RefMesh.from_arrays(m)
    K = ref.RefSparse.from_mesh(rm)
    rs, cols, _ = K.export()
    vals = rng.standard_normal(cols.size) * 10.0 ** rng.integers(-12, 12, cols.size)
    K.set_values(vals)
    x = rng.standard_normal(rs.size)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        ref.lib().refh_sparse_save(K.h, b"r.txt"); io.save_coo_text("o.txt", rs, cols, vals)
        ref.lib().refh_sparse_save_mat4(K.h, b"r.mat", b"K"); io.save_mat4_sparse("o.mat", "K", rs, cols, vals)
        ref.lib().refh_mat4_save_vec(b"r.mat", b"x", x.ctypes.data_as(ref.f64p), x.size)
        io.save_mat4_vector("o.mat", "x", x)
        assert read("r.txt") == read("o.txt") and read("r.mat") == read("o.mat")
        os.mkdir("a"); os.mkdir("b")
        assert ref.lib().refh_mesh_save_vtk(rm.h, b"a/m.vtk") == 0
        io.save_vtk("b/m.vtk", m)
        assert read("a/m.vtk").replace(b"a/m_extra", b"b/m_extra") == read("b/m.vtk")
    finally:
        os.chdir(cwd)


def test_nbt_header_matches_the_reference_writer(tmp_path):
    """NBT (file_format_nbt.c): for triangle and quad meshes the reference writes the header only and its reader
    reports failure on the (unimplemented) data section; both behaviours are mirrored."""
    for kind, name in ((0, "grid_trg.nbt"), (1, "grid_quad.nbt")):
        m = meshgen.structured_mesh(4, 3, 2.0, 1.0, kind=kind)
        out = str(tmp_path / name)
        assert io.save_nbt(out, m) == 0
        assert read(out) == read(os.path.join(GOLD, name))          # byte-identical to nb_mesh2D_save_nbt
        assert io.read_nbt_type(os.path.join(GOLD, name)) == (0, kind)
        assert io.load_nbt(os.path.join(GOLD, name), kind) == 1     # as nb_mesh2D_read_nbt
        assert io.load_nbt(os.path.join(GOLD, name), 1 - kind) == 1
    assert io.read_nbt_type(str(tmp_path / "missing.nbt")) == (1, None)
    assert io.read_nbt_type(os.path.join(GOLD, "grid_quad.vtk")) == (1, None)
    assert io.save_nbt(str(tmp_path / "no_such_dir" / "x.nbt"), m) == 1
    if ref.available():
        L = ref.lib()
        t = C.c_int(-1)
        assert L.refh_mesh_read_type_nbt(os.path.join(GOLD, "grid_quad.nbt").encode(), C.byref(t)) == 0 and t.value == 1
        rm = ref.RefMesh.from_arrays(m)
        assert L.refh_mesh_read_nbt(rm.h, os.path.join(GOLD, "grid_quad.nbt").encode()) == 1
        assert L.refh_mesh_save_nbt(rm.h, str(tmp_path / "r.nbt").encode()) == 0
        assert read(str(tmp_path / "r.nbt")) == read(os.path.join(GOLD, "grid_quad.nbt"))
