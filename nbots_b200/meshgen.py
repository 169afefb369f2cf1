"""Deterministic synthetic inputs for the hot path.

Plain numpy, no reference code involved: structured quad / triangle cantilever
meshes in the flat layout of the reference's mesh structs
(``sources/nb/geometric_bot/mesh/mesh2D/elements2D/mshquad_struct.h:6-26``,
``msh3trg_struct.h:6-23``) and the 9-point grid Laplacian of BASELINE.json
config 3 as a node graph (``headers/nb/graph_bot/graph.h:25-31``).

Node numbering is row-major ``j*(nx+1)+i`` (grid line after grid line), so a
contiguous block of matrix rows is a slab of whole grid lines -- the row
partition the multi-GPU path uses.
"""
from __future__ import annotations

import dataclasses
import numpy as np


@dataclasses.dataclass
class Mesh2D:
    """Flat mesh arrays; ``kind`` 0 = 3-node triangles, 1 = 4-node quads."""
    kind: int
    nod: np.ndarray          # float64 [2*N_nod]  x,y interleaved
    edg: np.ndarray          # uint32  [2*N_edg]
    adj: np.ndarray          # uint32  [npe*N_elems]  CCW connectivity
    vtx: np.ndarray          # uint32  [N_vtx]  mesh node of each input vertex
    sgm_sizes: np.ndarray    # uint32  [N_sgm]  nodes per input segment
    sgm_nodes: np.ndarray    # uint32  [sum(sgm_sizes)]  node ids along segments
    nx: int = 0              # structured meshes: elements per direction
    ny: int = 0

    @property
    def npe(self) -> int:
        return 4 if self.kind else 3

    @property
    def n_nod(self) -> int:
        return self.nod.size // 2

    @property
    def n_elems(self) -> int:
        return self.adj.size // self.npe

    @property
    def n_edg(self) -> int:
        return self.edg.size // 2

    def segment(self, s: int) -> np.ndarray:
        off = int(self.sgm_sizes[:s].sum())
        return self.sgm_nodes[off:off + int(self.sgm_sizes[s])]


def structured_mesh(nx: int, ny: int, lx: float, ly: float, kind: int = 1,
                    diagonal_seed: int | None = None) -> Mesh2D:
    """(nx x ny)-cell grid on [0,lx]x[0,ly]; quads, or each cell split in 2 triangles.

    With diagonal_seed the splitting diagonal of every cell is drawn at random, which
    gives the nodes 4 to 8 neighbours -- the ragged row lengths of an unstructured
    (Delaunay) triangle mesh, at any size, without a mesher.

    Input segments (CCW loop): 0 bottom y=0 (left to right), 1 right x=lx (bottom
    to top), 2 top (right to left), 3 left x=0 (top to bottom).  Input vertices:
    the four corners in the same loop order.
    """
    NX, NY = nx + 1, ny + 1
    ii, jj = np.meshgrid(np.arange(NX), np.arange(NY), indexing="xy")
    nod = np.empty((NY * NX, 2), dtype=np.float64)
    nod[:, 0] = (ii * (lx / nx)).ravel()
    nod[:, 1] = (jj * (ly / ny)).ravel()
    nid = (jj * NX + ii).astype(np.uint32)

    h_edges = np.stack([nid[:, :-1].ravel(), nid[:, 1:].ravel()], axis=1)
    v_edges = np.stack([nid[:-1, :].ravel(), nid[1:, :].ravel()], axis=1)
    n00 = nid[:-1, :-1].ravel()
    n10 = nid[:-1, 1:].ravel()
    n11 = nid[1:, 1:].ravel()
    n01 = nid[1:, :-1].ravel()
    if kind == 1:
        adj = np.stack([n00, n10, n11, n01], axis=1)
        edges = np.concatenate([h_edges, v_edges])
    else:
        t0 = np.stack([n00, n10, n11], axis=1)
        t1 = np.stack([n00, n11, n01], axis=1)
        d_edges = np.stack([n00, n11], axis=1)
        if diagonal_seed is not None:
            flip = np.random.default_rng(diagonal_seed).integers(0, 2, n00.size).astype(bool)
            t0[flip] = np.stack([n00, n10, n01], axis=1)[flip]
            t1[flip] = np.stack([n10, n11, n01], axis=1)[flip]
            d_edges[flip] = np.stack([n10, n01], axis=1)[flip]
        adj = np.empty((2 * n00.size, 3), dtype=np.uint32)
        adj[0::2] = t0
        adj[1::2] = t1
        edges = np.concatenate([h_edges, v_edges, d_edges])

    bottom = nid[0, :]
    right = nid[:, -1]
    top = nid[-1, ::-1]
    left = nid[::-1, 0]
    sgm_sizes = np.array([NX, NY, NX, NY], dtype=np.uint32)
    sgm_nodes = np.concatenate([bottom, right, top, left]).astype(np.uint32)
    vtx = np.array([nid[0, 0], nid[0, -1], nid[-1, -1], nid[-1, 0]], dtype=np.uint32)
    return Mesh2D(kind=kind, nod=nod.ravel().copy(),
                  edg=edges.astype(np.uint32).ravel().copy(),
                  adj=adj.astype(np.uint32).ravel().copy(), vtx=vtx,
                  sgm_sizes=sgm_sizes, sgm_nodes=sgm_nodes, nx=nx, ny=ny)


def quad_counts(nx: int, ny: int) -> tuple[int, int]:
    """(N dof, nnz) of the 2-dof elasticity matrix on an nx x ny quad grid."""
    NX, NY = nx + 1, ny + 1
    nnz = 4 * (9 * (NX - 2) * (NY - 2) + 6 * (2 * (NX - 2) + 2 * (NY - 2)) + 16)
    return 2 * NX * NY, nnz


def laplacian9_csr(n: int, row_begin: int = 0, row_end: int | None = None):
    """Rows [row_begin,row_end) of the n*n-grid 9-point operator (diag 8, off-diag -1).

    Returned as (rows_size u32, cols u32 ascending per row, vals f64): the flat
    form of what ``nb_sparse_create(graph,NULL,1)`` (sparse.c:20-60) produces for
    the 8-neighbour grid graph once values are filled in.
    """
    if row_end is None:
        row_end = n * n
    r = np.arange(row_begin, row_end, dtype=np.int64)
    j, i = np.divmod(r, n)
    cols = np.empty((r.size, 9), dtype=np.int64)
    valid = np.empty((r.size, 9), dtype=bool)
    k = 0
    for dj in (-1, 0, 1):
        for di in (-1, 0, 1):
            jj, ii = j + dj, i + di
            valid[:, k] = (jj >= 0) & (jj < n) & (ii >= 0) & (ii < n)
            cols[:, k] = jj * n + ii
            k += 1
    vals = np.where(np.arange(9)[None, :] == 4, 8.0, -1.0) * np.ones((r.size, 1))
    rows_size = valid.sum(axis=1).astype(np.uint32)
    return rows_size, cols[valid].astype(np.uint32), vals[valid].astype(np.float64)


def laplacian9_graph(n: int):
    """8-neighbour grid graph as (N_adj u32[N], adj_flat u32) for nb_sparse_create."""
    rows_size, cols, _ = laplacian9_csr(n)
    N = n * n
    row_of = np.repeat(np.arange(N, dtype=np.uint32), rows_size)
    keep = cols != row_of
    return (rows_size - 1).astype(np.uint32), cols[keep].copy()


def uniform_rhs(n: int, seed: int = 12345, start: int = 0) -> np.ndarray:
    """Deterministic uniform(-0.5,0.5) vector (SURVEY §8d input 3).

    Counter-based (splitmix64 of ``seed + index``) so that any rank can generate
    its own slice ``[start, start+n)`` of the global right-hand side.
    """
    with np.errstate(over="ignore"):
        z = np.arange(start, start + n, dtype=np.uint64) + np.uint64(seed)
        z = z * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53)) - 0.5
