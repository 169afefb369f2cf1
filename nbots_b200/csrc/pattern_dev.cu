// pattern_dev.cu -- the stiffness sparsity pattern built ON THE DEVICE, straight into the SELL-32 column
// arrays, bit-exact with the reference's two host steps
//   nb_mesh2D_load_graph(mesh, graph, NB_NODES_LINKED_BY_ELEMS)   (mesh2D/load_graph.c:230-328)
//   nb_sparse_create(graph, NULL, 2)                               (solver_bot/sparse/sparse.c:20-60)
// which cost the reference two heap blocks and one qsort per row (0.9 s per 1 M dof).
//
// A node's row holds the dofs of every node it shares an element with (mesh edges are element sides
// on a conforming mesh -- checked below -- and the intra-element non-edge pairs are the rest), columns
// ascending.  One thread per node walks its elements (the node -> element lists of nbgpu_mesh_create)
// and emits its neighbours in ascending order by repeated "smallest candidate above the last one":
// at most (valence x nodes per element) candidates, no sorting, no scratch memory.  Pass 1 counts,
// pass 2 writes the per-entry column ids, the 2x2-block node ids and their 16-bit differences of both
// rows of the node directly at their SELL positions.  Nothing of the pattern is built on the host or
// uploaded: the CSR mirror the export calls need is downloaded lazily (matrix.cu: ensure_host_pattern).
//
// This fast path takes matrices whose slices can all be stored max-width wide at < 1 % extra entries
// (structured meshes: the bench, Q16, the SIMP loop); ragged patterns go through the host builder and
// its sigma-window sorting (pattern.cu + matrix.cu).
#include <algorithm>
#include <chrono>
#include <cstring>

#include "matrix.cuh"
#include "mesh.cuh"

using namespace nbgpu;

namespace {

constexpr uint32_t kNone = 0xFFFFFFFFu;

// smallest node id > last (last = kNone: smallest of all) among the nodes of the elements around `node`
template <int NPE>
__device__ __forceinline__ uint32_t next_neighbour(uint32_t e0, uint32_t e1, const uint32_t *__restrict__ n2e,
						  const uint32_t *__restrict__ adj, uint32_t last)
{
	uint32_t best = kNone;
	for (uint32_t t = e0; t < e1; t++) {
		const uint32_t e = n2e[t];
#pragma unroll
		for (int i = 0; i < NPE; i++) {
			const uint32_t v = adj[(size_t)e * NPE + i];
			if ((last == kNone || v > last) && v < best)
				best = v;
		}
	}
	return best;
}

template <int NPE>
__global__ void __launch_bounds__(256)
count_neighbours_kernel(uint32_t N_nod, const uint32_t *__restrict__ adj, const uint32_t *__restrict__ n2e_ptr,
			const uint32_t *__restrict__ n2e, uint32_t *__restrict__ counts, unsigned int *max_count,
			unsigned long long *total, unsigned long long *units)
{
	const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t n = 0;
	if (node < N_nod) {
		const uint32_t e0 = n2e_ptr[node], e1 = n2e_ptr[node + 1];
		if (e0 == e1) {
			n = 1;   // an isolated node still owns its diagonal (sparse.c:40-47)
		} else {
			uint32_t last = kNone;
			for (;;) {
				const uint32_t v = next_neighbour<NPE>(e0, e1, n2e, adj, last);
				if (v == kNone)
					break;
				last = v;
				n++;
			}
		}
		counts[node] = n;
	}
	// a slice (32 rows) is 16 consecutive nodes = half a warp: its width is twice the largest count in it
	unsigned int m16 = n;
	for (int o = 8; o > 0; o >>= 1)
		m16 = max(m16, __shfl_xor_sync(0xffffffffu, m16, o));
	unsigned long long u = ((threadIdx.x & 15) == 0) ? 2ull * m16 : 0ull;
	// block-level reduction of max, sum and slice units, one atomic each per CTA
	__shared__ unsigned int s_max[8];
	__shared__ unsigned long long s_sum[8], s_units[8];
	unsigned int m = n;
	unsigned long long s = n;
	for (int o = 16; o > 0; o >>= 1) {
		m = max(m, __shfl_down_sync(0xffffffffu, m, o));
		s += __shfl_down_sync(0xffffffffu, s, o);
		u += __shfl_down_sync(0xffffffffu, u, o);
	}
	if ((threadIdx.x & 31) == 0) {
		s_max[threadIdx.x >> 5] = m;
		s_sum[threadIdx.x >> 5] = s;
		s_units[threadIdx.x >> 5] = u;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		for (int w = 1; w < 8; w++) {
			m = max(m, s_max[w]);
			s += s_sum[w];
			u += s_units[w];
		}
		atomicMax(max_count, m);
		atomicAdd(total, s);
		atomicAdd(units, u);
	}
}

// both rows of a node, uniform slice width `width` (entries), 2 dofs per node
template <int NPE>
__global__ void __launch_bounds__(256)
fill_pattern_kernel(uint32_t N_nod, uint32_t width, const uint32_t *__restrict__ adj,
		    const uint32_t *__restrict__ n2e_ptr, const uint32_t *__restrict__ n2e, uint32_t *__restrict__ col,
		    uint32_t *__restrict__ bcol, short *__restrict__ idx16, int *too_far)
{
	const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t n_pairs = ((2 * N_nod + kSliceRows - 1) / kSliceRows) * 16u;
	if (node >= n_pairs)
		return;
	const uint32_t slice = node >> 4, nl = node & 15, lane0 = 2 * nl;
	const size_t off = (size_t)slice * width;
	const uint32_t nb_max = width >> 1;
	uint32_t jb = 0;
	if (node < N_nod) {
		const uint32_t e0 = n2e_ptr[node], e1 = n2e_ptr[node + 1];
		uint32_t last = kNone;
		for (;;) {
			const uint32_t v = e0 == e1 ? (jb == 0 ? node : kNone) : next_neighbour<NPE>(e0, e1, n2e, adj, last);
			if (v == kNone)
				break;
			last = v;
			uint32_t *c = col + (off + 2 * jb) * kSliceRows + lane0;
			c[0] = 2 * v;
			c[1] = 2 * v;
			c[kSliceRows] = 2 * v + 1;
			c[kSliceRows + 1] = 2 * v + 1;
			bcol[((off >> 1) + jb) * 16u + nl] = v;
			const int64_t d = (int64_t)v - (int64_t)node;
			if (d < -32767 || d > 32767)
				*too_far = 1;
			idx16[((off >> 1) + jb) * 16u + nl] = (short)d;
			jb++;
		}
	}
	for (; jb < nb_max; jb++) {   // padding (also the rows past N in the last slice)
		uint32_t *c = col + (off + 2 * jb) * kSliceRows + lane0;
		c[0] = c[1] = c[kSliceRows] = c[kSliceRows + 1] = kPadCol;
		bcol[((off >> 1) + jb) * 16u + nl] = kPadCol;
		idx16[((off >> 1) + jb) * 16u + nl] = (short)-32768;
	}
}

__global__ void __launch_bounds__(256)
uniform_offsets_kernel(uint32_t n_slices, uint32_t width, uint32_t *__restrict__ slice_off)
{
	const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s <= n_slices)
		slice_off[s] = s * width;
}

// every mesh edge must be an element side (then the edge list adds nothing to the element graph)
__global__ void __launch_bounds__(256)
check_edges_kernel(uint32_t N_edg, const uint32_t *__restrict__ edg, uint32_t width, const uint32_t *__restrict__ bcol,
		   int *missing)
{
	const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= N_edg)
		return;
	const uint32_t a = edg[2 * e], b = edg[2 * e + 1];
	const uint32_t slice = a >> 4, nl = a & 15;
	const size_t off = (size_t)slice * width;
	bool found = false;
	for (uint32_t jb = 0; jb < (width >> 1) && !found; jb++)
		found = bcol[((off >> 1) + jb) * 16u + nl] == b;
	if (!found)
		*missing = 1;
}

}  // namespace

extern "C" {

/* The matrix of a device-resident mesh, pattern built on the device (values zero).  2 dofs per node.  edg may be
 * NULL; if given (host array), every edge must be an element side.  Returns NBGPU_OK with *out == NULL when the
 * pattern does not qualify for the device path (ragged rows, NBGPU_SIGMA / NBGPU_NO_UNIFORM / NBGPU_HOST_PATTERN
 * set): the caller then uses nbgpu_pattern_from_mesh + nbgpu_matrix_create_from_csr. */
int nbgpu_matrix_create_from_mesh(const nbgpu_mesh_t *mesh, uint32_t N_edg, const uint32_t *edg, nbgpu_matrix_t **out)
{
	NB_INIT();
	NB_ARG(mesh != nullptr && out != nullptr);
	*out = nullptr;
	if (getenv("NBGPU_HOST_PATTERN") || getenv("NBGPU_NO_UNIFORM") || getenv("NBGPU_SIGMA") || getenv("NBGPU_NO_BLOCKED") ||
	    mesh->N_nod == 0 || mesh->N_elems == 0)
		return NBGPU_OK;
	Context &c = ctx();
	const uint32_t N_nod = mesh->N_nod;
	const int grid_n = (int)((N_nod + 255) / 256);
	uint32_t *d_counts = nullptr;
	unsigned long long *d_scal = nullptr;   // [0] total, [1] max (low word), [2] too-far flag, [3] edge flag, [4] slice units
	NB_CUDA(nbgpu::dmalloc(&d_counts, (size_t)N_nod * sizeof(uint32_t)));
	cudaError_t e = nbgpu::dmalloc(&d_scal, 5 * sizeof(unsigned long long));
	if (e != cudaSuccess) {
		nbgpu::dfree(d_counts);
		NB_CUDA(e);
	}
	auto fail = [&](int st) {
		nbgpu::dfree(d_counts);
		nbgpu::dfree(d_scal);
		return st;
	};
	NB_CUDA(cudaMemsetAsync(d_scal, 0, 5 * sizeof(unsigned long long), c.stream));
	if (mesh->npe == 3)
		count_neighbours_kernel<3><<<grid_n, 256, 0, c.stream>>>(N_nod, mesh->d_adj, mesh->d_n2e_ptr, mesh->d_n2e, d_counts,
									  (unsigned int *)(d_scal + 1), d_scal, d_scal + 4);
	else
		count_neighbours_kernel<4><<<grid_n, 256, 0, c.stream>>>(N_nod, mesh->d_adj, mesh->d_n2e_ptr, mesh->d_n2e, d_counts,
									  (unsigned int *)(d_scal + 1), d_scal, d_scal + 4);
	c.launches++;
	unsigned long long h_scal[5] = {0, 0, 0, 0, 0};
	e = cudaMemcpyAsync(h_scal, d_scal, sizeof(h_scal), cudaMemcpyDeviceToHost, c.stream);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(c.stream);
	if (e != cudaSuccess) {
		set_error("device pattern: %s", cudaGetErrorString(e));
		return fail(NBGPU_ERR_CUDA);
	}
	// (the last half-warp of the grid may cover nodes past N_nod: they count 0 and add nothing)
	const uint64_t total_nb = h_scal[0];
	const uint32_t max_nb = (uint32_t)(h_scal[1] & 0xFFFFFFFFu);
	const uint64_t nnz = 4 * total_nb;
	const uint32_t N = 2 * N_nod, n_slices = (N + kSliceRows - 1) / kSliceRows, width = 2 * max_nb;
	const uint64_t stored = (uint64_t)n_slices * width * kSliceRows;
	// the uniform layout only when storing every slice max-width wide costs < 1 % extra entries over the
	// per-slice widths (the rule of matrix.cu: build_layout) and the natural order pads <= 5 % (else the
	// sigma-sorted layout of the host path); slice offsets must stay 32-bit
	const uint64_t units = h_scal[4], uniform_units = (uint64_t)n_slices * width;
	if (uniform_units > units + units / 100 || units * kSliceRows > nnz + nnz / 20 || uniform_units > 0xFFFFFFFFull)
		return fail(NBGPU_OK);
	nbgpu_matrix_t *A = new nbgpu_matrix_t();
	A->N = N;
	A->n_cols = N;
	A->nnz = nnz;
	A->n_slices = n_slices;
	A->stored = stored;
	A->max_width = width;
	A->uniform_width = width;
	A->sigma = 1;
	A->blocked = true;
	A->d_node_counts = d_counts;
	e = nbgpu::dmalloc(&A->d_slice_off, ((size_t)n_slices + 1) * sizeof(uint32_t));
	if (e == cudaSuccess)
		e = nbgpu::dmalloc(&A->d_val, std::max<size_t>(1, stored) * sizeof(double));
	if (e == cudaSuccess)
		e = nbgpu::dmalloc(&A->d_col, std::max<size_t>(1, stored) * sizeof(uint32_t));
	if (e == cudaSuccess)
		e = nbgpu::dmalloc(&A->d_bcol, std::max<size_t>(1, stored / 4) * sizeof(uint32_t));
	if (e == cudaSuccess)
		e = nbgpu::dmalloc(&A->d_idx16, std::max<size_t>(1, stored / 4) * sizeof(short));
	if (e != cudaSuccess) {
		set_error("matrix of %llu stored entries: %s", (unsigned long long)stored, cudaGetErrorString(e));
		cudaGetLastError();
		nbgpu::dfree(d_scal);
		nbgpu_matrix_destroy(A);
		return NBGPU_ERR_NOMEM;
	}
	cudaMemsetAsync(A->d_val, 0, stored * sizeof(double), c.stream);
	cudaMemsetAsync(d_scal + 2, 0, 2 * sizeof(unsigned long long), c.stream);
	uniform_offsets_kernel<<<(n_slices + 256) / 256, 256, 0, c.stream>>>(n_slices, width, A->d_slice_off);
	const int grid_p = (int)((n_slices * 16u + 255) / 256);
	if (mesh->npe == 3)
		fill_pattern_kernel<3><<<grid_p, 256, 0, c.stream>>>(N_nod, width, mesh->d_adj, mesh->d_n2e_ptr, mesh->d_n2e, A->d_col,
								      A->d_bcol, A->d_idx16, (int *)(d_scal + 2));
	else
		fill_pattern_kernel<4><<<grid_p, 256, 0, c.stream>>>(N_nod, width, mesh->d_adj, mesh->d_n2e_ptr, mesh->d_n2e, A->d_col,
								      A->d_bcol, A->d_idx16, (int *)(d_scal + 2));
	c.launches += 2;
	uint32_t *d_edg = nullptr;
	if (edg && N_edg) {
		e = nbgpu::dmalloc(&d_edg, 2 * (size_t)N_edg * sizeof(uint32_t));
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(d_edg, edg, 2 * (size_t)N_edg * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream);
		if (e == cudaSuccess) {
			check_edges_kernel<<<(N_edg + 255) / 256, 256, 0, c.stream>>>(N_edg, d_edg, width, A->d_bcol, (int *)(d_scal + 3));
			c.launches++;
		}
	}
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(h_scal, d_scal, sizeof(h_scal), cudaMemcpyDeviceToHost, c.stream);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(c.stream);
	if (e == cudaSuccess)
		e = cudaGetLastError();
	nbgpu::dfree(d_edg);
	nbgpu::dfree(d_scal);
	if (e != cudaSuccess) {
		set_error("device pattern: %s", cudaGetErrorString(e));
		nbgpu_matrix_destroy(A);
		return NBGPU_ERR_CUDA;
	}
	if ((int)h_scal[3]) {
		// an edge that is no element side: the element graph is not the reference's graph here
		nbgpu_matrix_destroy(A);
		return NBGPU_OK;
	}
	A->idx16 = !(int)h_scal[2] && !getenv("NBGPU_NO_IDX16");
	if (!A->idx16) {
		nbgpu::dfree(A->d_idx16);
		A->d_idx16 = nullptr;
	}
	*out = A;
	return NBGPU_OK;
}

}  // extern "C"
