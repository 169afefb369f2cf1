// matrix.cuh -- device-resident form of the reference's nb_sparse_t.
//
// Layout in HBM (SELL-32): rows are grouped in slices of 32 consecutive rows
// (one warp).  Slice s has width w_s = max row length in the slice and stores
// its entries column-major: entry j of row 32*s + lane sits at
//     (slice_off[s] + j) * 32 + lane
// so that a warp reading "entry j of its 32 rows" issues one fully coalesced
// 256-byte (values) / 128-byte (column ids) request.  Rows shorter than w_s
// are padded with value 0 and column id kPadCol; the padding is skipped
// arithmetically (never added), so a row sum is exactly the reference's
// ascending-column sum (sources/nb/solver_bot/sparse/sparse.c:405-414).
// slice_off is in units of 32 entries, which keeps it 32-bit up to 2^37 nnz.
#pragma once

#include <type_traits>
#include <vector>

#include "common.cuh"

struct nbgpu_matrix_s {
	uint32_t N = 0;                       // rows
	uint32_t n_cols = 0;                  // column space (== N except for a rank-local block)
	bool local_block = false;             // rank-local block of a partitioned matrix (dist.cu): column space
					      // "lower halo | owned | upper halo" of n_cols entries
	uint32_t col_shift = 0;               // ... in which row r is column r + col_shift
	uint64_t nnz = 0;
	uint32_t n_slices = 0;
	uint64_t stored = 0;                  // padded entry count = 32 * slice_off[n_slices]
	uint32_t max_width = 0;
	uint32_t uniform_width = 0;           // != 0: every slice is stored max_width wide (slice_off[s] = s * width);
					      // chosen when that costs < 1 % extra entries, lets the kernels
					      // compute slice offsets instead of loading them
	uint32_t *d_slice_off = nullptr;      // [n_slices + 1]
	double *d_val = nullptr;              // [stored]
	uint32_t *d_col = nullptr;            // [stored]
	// 2x2-block structure (2 dofs per node): rows 2i, 2i+1 share their columns and
	// columns come in (2c, 2c+1) pairs.  Then d_bcol holds one node id per block:
	// block jb of the node pair nl of slice s sits at (slice_off[s]/2 + jb) * 16 + nl.
	bool blocked = false;
	uint32_t *d_bcol = nullptr;           // [stored / 4]
	// 16-bit column ids: when every stored column lies within +-32767 of its row (node ids for a blocked
	// matrix) -- true for banded numberings such as grid meshes up to 32 k nodes per line -- a copy of
	// the ids as int16 differences is kept and the streamed kernels read that instead: 10 instead of 12
	// bytes per entry (8.5 instead of 9 when blocked).  Same layout as d_col / d_bcol; padding = -32768.
	bool idx16 = false;
	short *d_idx16 = nullptr;             // [stored] or [stored / 4]
	// SELL-C-sigma: inside windows of `sigma` consecutive rows the rows (row PAIRS when the two
	// dofs of a node always have equal length, so that the block structure survives) are stored in
	// order of descending length, which removes most of the padding of irregular (triangle-mesh)
	// patterns.  d_perm[position] = row (0xFFFFFFFF past the last row), d_inv_perm[row] = position,
	// position = slice * 32 + lane.  Both null when the order is the identity (sigma == 1).
	uint32_t sigma = 1;
	uint32_t *d_perm = nullptr;           // [n_slices * 32]
	uint32_t *d_inv_perm = nullptr;       // [N]
	int layout() const { return (blocked ? 1 : 0) | (idx16 ? 2 : 0); }   // template selector of the kernels
	const void *stream_ids() const
	{
		return idx16 ? (const void *)d_idx16 : blocked ? (const void *)d_bcol : (const void *)d_col;
	}
	std::vector<uint32_t> h_rows_size;    // host copy of the pattern's row lengths
	std::vector<uint64_t> h_row_ptr;      // CSR offsets (host), for value import/export
	// pattern built on the device (pattern_dev.cu): entries per node row / 2; the two host vectors above
	// are then filled on first use (ensure_host_pattern)
	uint32_t *d_node_counts = nullptr;
};

namespace nbgpu {

// f(std::integral_constant<int, L>) for the runtime layout L = blocked | idx16 << 1
template <typename F>
auto by_layout(int layout, F f)
{
	switch (layout) {
	case 0: return f(std::integral_constant<int, 0>{});
	case 1: return f(std::integral_constant<int, 1>{});
	case 2: return f(std::integral_constant<int, 2>{});
	default: return f(std::integral_constant<int, 3>{});
	}
}

// host mirror of the pattern (rows_size, row_ptr) of a matrix whose pattern was built on the device
int ensure_host_pattern(nbgpu_matrix_s *A);

// entry index of (row, j) in the SELL arrays
__host__ __device__ __forceinline__ size_t sell_index(const uint32_t *slice_off,
						       uint32_t row, uint32_t j)
{
	return ((size_t)slice_off[row >> 5] + j) * kSliceRows + (row & 31);
}

}  // namespace nbgpu
