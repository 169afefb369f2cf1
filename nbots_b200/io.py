"""On-disk formats either side of the solver path, host side only.

Writers produce, byte for byte, what the reference writes, so files can be exchanged with a program
built on it; the loaders are this package's own (the reference has a reader only for MAT v4):

* COO text          -- nb_sparse_save            (sources/nb/solver_bot/sparse/sparse.c:97-112)
                       "N N" then "row col %e" per stored entry (6 significant digits: lossy).
* MATLAB v4 binary  -- nb_sparse_save_mat4       (sources/nb/solver_bot/matlab_v4.c:184-250)
                       nb_mat4_save_vec          (matlab_v4.c:537-564); records are APPENDED to the file.
                       Lossless: the way to move a system between the reference and this library.
* VTK legacy ASCII  -- nb_mesh2D_save_vtk        (sources/nb/geometric_bot/mesh/mesh2D/mesh2D_file_format_vtk.c:23-89)
                       coordinates go through `float` there, so they carry ~7 digits.
"""
from __future__ import annotations

import struct

import numpy as np

from .meshgen import Mesh2D


def _csr(rows_size, cols, vals):
    rows_size = np.ascontiguousarray(rows_size, dtype=np.uint32)
    cols = np.ascontiguousarray(cols, dtype=np.uint32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    if cols.size != int(rows_size.sum(dtype=np.int64)) or vals.size != cols.size:
        raise ValueError("rows_size / cols / vals do not describe one CSR matrix")
    return rows_size, cols, vals


def _from_triplets(N, rows, cols, vals):
    """CSR with ascending columns per row (the invariant of nb_sparse_t); duplicates are an error."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.float64)
    if rows.size and (rows.min() < 0 or rows.max() >= N or cols.min() < 0 or cols.max() >= N):
        raise ValueError("index outside the matrix")
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    if rows.size > 1 and np.any((rows[1:] == rows[:-1]) & (cols[1:] == cols[:-1])):
        raise ValueError("duplicate entry")
    rows_size = np.bincount(rows, minlength=N).astype(np.uint32)
    return rows_size, cols.astype(np.uint32), vals


# --------------------------------------------------------------------------------- COO text --

def save_coo_text(path, rows_size, cols, vals):
    rows_size, cols, vals = _csr(rows_size, cols, vals)
    N = rows_size.size
    rows = np.repeat(np.arange(N, dtype=np.int64), rows_size)
    with open(path, "w") as fp:
        fp.write("%i %i\n" % (N, N))
        fp.write("".join("%i %i %e \n" % (i, j, v) for i, j, v in zip(rows.tolist(), cols.tolist(), vals.tolist())))


def load_coo_text(path):
    """-> (rows_size, cols, vals).  Values are what "%e" kept of them."""
    with open(path) as fp:
        head = fp.readline().split()
        if len(head) != 2 or head[0] != head[1]:
            raise ValueError("not an nb_sparse_save file: bad header")
        N = int(head[0])
        body = np.loadtxt(fp, dtype=np.float64, ndmin=2)
    if body.size == 0:
        return np.zeros(N, np.uint32), np.zeros(0, np.uint32), np.zeros(0)
    if body.shape[1] != 3:
        raise ValueError("not an nb_sparse_save file: expected 3 columns")
    return _from_triplets(N, body[:, 0], body[:, 1], body[:, 2])


# ------------------------------------------------------------------------ MATLAB v4 binary --

def _mat4_header(kind, n_rows, n_cols, label):
    name = label.encode() + b"\0"
    return struct.pack("<5i", kind, n_rows, n_cols, 0, len(name)) + name


def save_mat4_sparse(path, label, rows_size, cols, vals):
    """Appends a sparse record: (nnz + 1) x 3 doubles, columns [row+1 .. N], [col+1 .. N], [value .. 0],
    entries ordered by column then row."""
    rows_size, cols, vals = _csr(rows_size, cols, vals)
    N = rows_size.size
    rows = np.repeat(np.arange(N, dtype=np.int64), rows_size)
    order = np.lexsort((rows, cols.astype(np.int64)))
    irows = rows[order].astype(np.float64) + 1.0
    icols = cols[order].astype(np.float64) + 1.0
    with open(path, "ab") as fp:
        fp.write(_mat4_header(2, cols.size + 1, 3, label))
        fp.write(irows.tobytes()); fp.write(struct.pack("<d", float(N)))
        fp.write(icols.tobytes()); fp.write(struct.pack("<d", float(N)))
        fp.write(vals[order].tobytes()); fp.write(struct.pack("<d", 0.0))


def save_mat4_vector(path, label, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    with open(path, "ab") as fp:
        fp.write(_mat4_header(0, x.size, 1, label))
        fp.write(x.tobytes())


def load_mat4(path):
    """-> {label: vector | dense matrix | (rows_size, cols, vals)} for the double-precision records."""
    out = {}
    with open(path, "rb") as fp:
        while True:
            head = fp.read(20)
            if len(head) < 20:
                break
            kind, n_rows, n_cols, _imag, name_len = struct.unpack("<5i", head)
            label = fp.read(name_len).rstrip(b"\0").decode()
            data = np.frombuffer(fp.read(8 * n_rows * n_cols), dtype=np.float64)
            if data.size != n_rows * n_cols:
                raise ValueError("truncated MAT v4 record %r" % label)
            if kind == 0:
                out[label] = data.copy() if n_cols == 1 else data.reshape(n_cols, n_rows).T.copy()
            elif kind == 2:
                if n_cols != 3:
                    raise ValueError("complex sparse records are not supported")
                t = data.reshape(3, n_rows)          # column-major (nnz + 1) x 3
                N = int(max(t[0, -1], t[1, -1]))
                out[label] = _from_triplets(N, t[0, :-1] - 1, t[1, :-1] - 1, t[2, :-1])
            else:
                raise ValueError("MAT v4 record type %d is not supported" % kind)
    return out


# -------------------------------------------------------------------------- VTK legacy ASCII --

_VTK_TYPE = {0: "NB_TRIAN", 1: "NB_QUAD"}
_VTK_CELL = {3: 5, 4: 9}


def save_vtk(path, mesh: Mesh2D):
    if not str(path).endswith(".vtk") and ".vtk" not in str(path):
        raise ValueError("the reference derives the side-file name from '.vtk' in the path")
    extra = str(path)
    k = extra.find(".vtk")
    extra = extra[:k] + "_extra.txt"
    npe = mesh.npe
    nod = mesh.nod.reshape(-1, 2).astype(np.float32)       # the reference prints `float` copies
    adj = mesh.adj.reshape(-1, npe)
    with open(path, "w") as fp:
        fp.write("# vtk DataFile Version 2.0\n")
        fp.write("# nbots nb_mesh2D_t 1.0 type=%s extra_file=%s\n" % (_VTK_TYPE[mesh.kind], extra))
        fp.write("ASCII\nDATASET UNSTRUCTURED_GRID\n")
        fp.write("POINTS %i float\n" % nod.shape[0])
        fp.write("".join(" %f %f 0\n" % (float(x), float(y)) for x, y in nod))
        fp.write("CELLS %i %i\n" % (adj.shape[0], adj.size + adj.shape[0]))
        fp.write("".join(" %i " % npe + "".join("%i " % v for v in row) + "\n" for row in adj.tolist()))
        fp.write("CELL_TYPES %i\n" % adj.shape[0])
        fp.write((" %i\n" % _VTK_CELL[npe]) * adj.shape[0])


def load_vtk(path) -> Mesh2D:
    """Nodes and elements of a triangle or quad mesh written by save_vtk / nb_mesh2D_save_vtk.  The file
    holds no edges, input vertices or input segments: `edg` is rebuilt from the element sides, the
    boundary lists come back empty (boundary conditions then go in as dof lists)."""
    with open(path) as fp:
        tok = fp.read().split("\n")
    if len(tok) < 5 or not tok[0].startswith("# vtk DataFile") or "nbots nb_mesh2D_t" not in tok[1]:
        raise ValueError("not a mesh2D VTK file")
    kind = 1 if "type=NB_QUAD" in tok[1] else 0 if "type=NB_TRIAN" in tok[1] else None
    if kind is None:
        raise ValueError("only NB_TRIAN and NB_QUAD meshes are supported")
    words = " ".join(tok[4:]).split()
    pos = 0

    def expect(word):
        nonlocal pos
        if words[pos] != word:
            raise ValueError("expected %s, found %s" % (word, words[pos]))
        pos += 1

    expect("POINTS")
    n_nod = int(words[pos]); pos += 2
    pts = np.array(words[pos:pos + 3 * n_nod], dtype=np.float64).reshape(n_nod, 3); pos += 3 * n_nod
    expect("CELLS")
    n_el, total = int(words[pos]), int(words[pos + 1]); pos += 2
    cells = np.array(words[pos:pos + total], dtype=np.int64); pos += total
    npe = 4 if kind else 3
    cells = cells.reshape(n_el, npe + 1)
    if np.any(cells[:, 0] != npe):
        raise ValueError("mixed cell sizes")
    expect("CELL_TYPES")
    if int(words[pos]) != n_el or any(int(w) != _VTK_CELL[npe] for w in words[pos + 1:pos + 1 + n_el]):
        raise ValueError("cell types do not match the mesh type")
    adj = cells[:, 1:].astype(np.uint32)
    sides = np.concatenate([np.stack([adj[:, i], adj[:, (i + 1) % npe]], axis=1) for i in range(npe)])
    sides = np.unique(np.sort(sides, axis=1), axis=0).astype(np.uint32)
    return Mesh2D(kind=kind, nod=pts[:, :2].ravel().copy(), edg=sides.ravel().copy(), adj=adj.ravel().copy(),
                  vtx=np.zeros(0, np.uint32), sgm_sizes=np.zeros(0, np.uint32), sgm_nodes=np.zeros(0, np.uint32),
                  nx=0, ny=0)


# ---- NBT (mesh2D/file_format_nbt.c) ---------------------------------------------------------------------------
# What the reference implements of this format for the FEM path's two mesh kinds is the HEADER: its triangle and
# quad data writers are empty and their readers return 1 (elements2D/msh3trg_file_format_nbt.c:9-17,
# mshquad_file_format_nbt.c:13-22); only polygon meshes carry data.  Mirrored as it is: save_nbt writes the
# byte-identical header, read_nbt_type parses it (nb_mesh2D_read_type_nbt, file_format_nbt.c:99-160), and
# load_nbt fails the way nb_mesh2D_read_nbt does -- a mesh has to travel as VTK (above) or as arrays.
NBT_HEADER = "[Numerical Bots File Format v1.0]"          # headers/nb/io_bot/nbt_file_format.h:4
_NBT_TYPES = {0: "NB_TRIAN", 1: "NB_QUAD"}


def save_nbt(path, m: Mesh2D) -> int:
    """nb_mesh2D_save_nbt (file_format_nbt.c:28-45): 0 on success, 1 if the file cannot be opened."""
    try:
        with open(path, "w") as fp:
            fp.write("%s\nClass = mesh2D\nType = %s\n\n" % (NBT_HEADER, _NBT_TYPES[m.kind]))
    except OSError:
        return 1
    return 0


def read_nbt_type(path):
    """nb_mesh2D_read_type_nbt -> (status, kind): 0 and 0 / 1 (triangles / quads) for a mesh2D NBT file of one of the
    two FEM mesh kinds, status 1 otherwise (missing file, wrong header or class, another type)."""
    try:
        with open(path) as fp:
            lines = [ln.split("#")[0].strip() for ln in fp]
    except OSError:
        return 1, None
    lines = [ln for ln in lines if ln]
    if len(lines) < 3 or lines[0] != NBT_HEADER:
        return 1, None
    var, _, val = lines[1].partition("=")
    if var.strip() != "Class" or val.strip() != "mesh2D":
        return 1, None
    var, _, val = lines[2].partition("=")
    if var.strip() != "Type":
        return 1, None
    for kind, name in _NBT_TYPES.items():
        if val.strip() == name:
            return 0, kind
    return 1, None


def load_nbt(path, kind) -> int:
    """nb_mesh2D_read_nbt for a triangle / quad mesh: status 1 -- after the header and type checks the reference
    calls a data reader that is not implemented for these kinds and reports failure (file_format_nbt.c:163-191)."""
    st, _found = read_nbt_type(path)
    if st != 0 or _found != kind:
        return 1
    return 1
