// matrix.cu -- nb_sparse_t -> SELL-32 in HBM: creation, value import/export.
// Reference: struct nb_sparse_s (sources/nb/solver_bot/sparse/sparse_struct.h:6-11),
// nb_sparse_create (sparse.c:20-60), nb_sparse_reset (sparse.c:127-132).
#include <cstring>
#include <algorithm>
#include <chrono>

#include "matrix.cuh"

using namespace nbgpu;

namespace {

constexpr size_t kStageBytes = size_t(32) << 20;   // per pinned staging buffer

// One warp per slice: transpose CSR rows into the slice's column-major block.
// csr_val == nullptr writes zero values; sell_col == nullptr leaves the
// pattern untouched (value-only import).  *bad is raised if a row's columns
// are not strictly ascending or out of range (the reference's invariant after
// nb_qsort, sparse.c:55).
__global__ void __launch_bounds__(kBlock)
csr_to_sell_kernel(uint32_t N, uint32_t n_cols, bool sorted, uint32_t n_slices, const uint64_t *__restrict__ row_ptr,
		   const uint32_t *__restrict__ csr_col, const double *__restrict__ csr_val,
		   const uint32_t *__restrict__ slice_off, const uint32_t *__restrict__ perm,
		   uint32_t *__restrict__ sell_col, double *__restrict__ sell_val, int *bad)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slices; s += warps) {
		const uint32_t row = perm ? perm[s * kSliceRows + lane] : s * kSliceRows + lane;
		const uint32_t off = slice_off[s], width = slice_off[s + 1] - off;
		uint64_t base = 0;
		uint32_t len = 0;
		if (row < N) {
			base = row_ptr[row];
			len = (uint32_t)(row_ptr[row + 1] - base);
		}
		uint32_t prev = 0;
		for (uint32_t j = 0; j < width; j++) {
			const size_t idx = ((size_t)off + j) * kSliceRows + lane;
			if (j < len) {
				if (sell_col) {
					uint32_t c = csr_col[base + j];
					if (c >= n_cols || (sorted && j > 0 && c <= prev))
						*bad = 1;
					prev = c;
					sell_col[idx] = c;
				}
				sell_val[idx] = csr_val ? csr_val[base + j] : 0.0;
			} else {
				if (sell_col)
					sell_col[idx] = kPadCol;
				sell_val[idx] = 0.0;
			}
		}
	}
}

// Does the pattern have the 2-dofs-per-node block structure?  Lanes 2k, 2k+1 of a
// slice hold rows 2i, 2i+1; *not_blocked is raised on the first violation.
__global__ void __launch_bounds__(kBlock)
check_blocked_kernel(uint32_t N, uint32_t n_slices, const uint32_t *__restrict__ slice_off,
		     const uint32_t *__restrict__ sell_col, int *not_blocked)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slices; s += warps) {
		const uint32_t off = slice_off[s], width = slice_off[s + 1] - off;
		bool bad = (width & 1u) != 0;
		for (uint32_t j = 0; j + 1 < width; j += 2) {
			const uint32_t c0 = sell_col[((size_t)off + j) * kSliceRows + lane];
			const uint32_t c1 = sell_col[((size_t)off + j + 1) * kSliceRows + lane];
			const uint32_t p0 = __shfl_xor_sync(0xffffffffu, c0, 1);
			if (c0 == kPadCol)
				bad |= c1 != kPadCol;
			else
				bad |= (c0 & 1u) != 0 || c1 != c0 + 1;
			bad |= p0 != c0;
		}
		if (bad)
			*not_blocked = 1;
	}
	(void)N;
}

__global__ void __launch_bounds__(kBlock)
build_bcol_kernel(uint32_t n_slices, const uint32_t *__restrict__ slice_off,
		  const uint32_t *__restrict__ sell_col, uint32_t *__restrict__ bcol)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slices; s += warps) {
		const uint32_t off = slice_off[s], width = slice_off[s + 1] - off;
		// 16 node pairs x width/2 blocks; lane handles node pair (lane & 15), blocks lane>>4, +2, ...
		for (uint32_t jb = lane >> 4; jb < (width >> 1); jb += 2) {
			const uint32_t c = sell_col[((size_t)off + 2 * jb) * kSliceRows + 2 * (lane & 15)];
			bcol[((size_t)(off >> 1) + jb) * 16u + (lane & 15)] = (c == kPadCol) ? kPadCol : (c >> 1);
		}
	}
}

// int16 differences of the column ids to the row's own index (per-block node ids to the row's node when
// `blocked`), same positions as the source array; *too_far is raised when one does not fit.
__global__ void __launch_bounds__(kBlock)
build_idx16_kernel(bool blocked, uint32_t N, uint32_t col_shift, uint32_t n_slices,
		   const uint32_t *__restrict__ slice_off, const uint32_t *__restrict__ perm,
		   const uint32_t *__restrict__ ids, short *__restrict__ out, int *too_far)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slices; s += warps) {
		const uint32_t off = slice_off[s], width = slice_off[s + 1] - off;
		if (!blocked) {
			const uint32_t row = perm ? perm[s * kSliceRows + lane] : s * kSliceRows + lane;
			for (uint32_t j = 0; j < width; j++) {
				const size_t idx = ((size_t)off + j) * kSliceRows + lane;
				const uint32_t c = ids[idx];
				int d = -32768;
				if (c != kPadCol) {
					const int64_t diff = (int64_t)c - (int64_t)row - (int64_t)col_shift;
					if (diff < -32767 || diff > 32767 || row >= N)
						*too_far = 1;
					d = (int)diff;
				}
				out[idx] = (short)d;
			}
		} else {
			const uint32_t nl = lane & 15;
			const uint32_t row = perm ? perm[s * kSliceRows + 2 * nl] : s * kSliceRows + 2 * nl;
			for (uint32_t jb = lane >> 4; jb < (width >> 1); jb += 2) {
				const size_t idx = ((size_t)(off >> 1) + jb) * 16u + nl;
				const uint32_t c = ids[idx];
				int d = -32768;
				if (c != kPadCol) {
					const int64_t diff = (int64_t)c - (int64_t)((row + col_shift) >> 1);
					if (diff < -32767 || diff > 32767 || row >= N)
						*too_far = 1;
					d = (int)diff;
				}
				out[idx] = (short)d;
			}
		}
	}
}

__global__ void __launch_bounds__(kBlock)
sell_to_csr_kernel(uint32_t N, uint32_t n_slices, const uint64_t *__restrict__ row_ptr,
		   const uint32_t *__restrict__ slice_off, const uint32_t *__restrict__ perm,
		   const uint32_t *__restrict__ sell_col, const double *__restrict__ sell_val,
		   uint32_t *__restrict__ csr_col, double *__restrict__ csr_val)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slices; s += warps) {
		const uint32_t row = perm ? perm[s * kSliceRows + lane] : s * kSliceRows + lane;
		if (row >= N)
			continue;
		const uint32_t off = slice_off[s];
		const uint64_t base = row_ptr[row];
		const uint32_t len = (uint32_t)(row_ptr[row + 1] - base);
		for (uint32_t j = 0; j < len; j++) {
			const size_t idx = ((size_t)off + j) * kSliceRows + lane;
			if (csr_col)
				csr_col[base + j] = sell_col[idx];
			if (csr_val)
				csr_val[base + j] = sell_val[idx];
		}
	}
}

int grid_for_slices(uint32_t n_slices)
{
	int64_t want = ((int64_t)n_slices * 32 + kBlock - 1) / kBlock;
	int64_t cap = (int64_t)ctx().sm_count * 8;
	return (int)std::max<int64_t>(1, std::min(want, cap));
}

// Host -> device copy of `count` elements of T that the caller can only hand
// over piecewise: fill(chunk, first, last, dst) must write elements [first,last) of
// the logical array to dst[0..last-first).  Double-buffered pinned staging on
// the copy stream, so gathering chunk k+1 overlaps the DMA of chunk k.
template <typename T, typename Fill>
int upload_staged(T *d_dst, size_t count, const std::vector<size_t> &cuts, Fill fill)
{
	Context &c = ctx();
	NB_TRY(ensure_stage(kStageBytes));
	int buf = 0;
	for (size_t k = 0; k + 1 < cuts.size(); k++) {
		size_t first = cuts[k], last = cuts[k + 1];
		if (last == first)
			continue;
		NB_CUDA(cudaEventSynchronize(c.ev_stage[buf]));
		fill(k, first, last, (T *)c.stage[buf]);
		NB_CUDA(cudaMemcpyAsync(d_dst + first, c.stage[buf], (last - first) * sizeof(T),
					cudaMemcpyHostToDevice, c.copy_stream));
		NB_CUDA(cudaEventRecord(c.ev_stage[buf], c.copy_stream));
		buf ^= 1;
	}
	(void)count;
	return NBGPU_OK;
}

// element cuts of at most `max_elems` that fall on row boundaries
std::vector<size_t> row_cuts(const std::vector<uint64_t> &row_ptr, size_t max_elems,
			     std::vector<uint32_t> *cut_rows)
{
	std::vector<size_t> cuts{0};
	cut_rows->assign(1, 0);
	const uint32_t N = (uint32_t)row_ptr.size() - 1;
	uint32_t r = 0;
	while (r < N) {
		uint64_t limit = row_ptr[r] + max_elems;
		uint32_t hi = (uint32_t)(std::upper_bound(row_ptr.begin() + r, row_ptr.end(), limit) -
					 row_ptr.begin()) - 1;
		if (hi <= r)
			hi = r + 1;   // a single row longer than the buffer cannot happen (len <= 2^32)
		cuts.push_back(row_ptr[hi]);
		cut_rows->push_back(hi);
		r = hi;
	}
	return cuts;
}

std::vector<size_t> flat_cuts(size_t count, size_t max_elems)
{
	std::vector<size_t> cuts{0};
	for (size_t p = 0; p < count;) {
		p = std::min(count, p + max_elems);
		cuts.push_back(p);
	}
	return cuts;
}

template <typename T>
int upload_flat(T *d_dst, const T *src, size_t count)
{
	auto cuts = flat_cuts(count, kStageBytes / sizeof(T));
	return upload_staged<T>(d_dst, count, cuts, [&](size_t, size_t first, size_t last, T *dst) {
		const size_t n = last - first;
		const size_t piece = size_t(1) << 18;
#pragma omp parallel for schedule(static)
		for (int64_t p = 0; p < (int64_t)((n + piece - 1) / piece); p++) {
			size_t b = (size_t)p * piece, e = std::min(n, b + piece);
			memcpy(dst + b, src + first + b, (e - b) * sizeof(T));
		}
	});
}

template <typename T>
int upload_rows(T *d_dst, T *const *rows, const std::vector<uint64_t> &row_ptr)
{
	std::vector<uint32_t> cut_rows;
	auto cuts = row_cuts(row_ptr, kStageBytes / sizeof(T), &cut_rows);
	return upload_staged<T>(d_dst, row_ptr.back(), cuts, [&](size_t k, size_t first, size_t, T *dst) {
		const uint32_t r0 = cut_rows[k], r1 = cut_rows[k + 1];
#pragma omp parallel for schedule(static)
		for (int64_t r = r0; r < (int64_t)r1; r++)
			memcpy(dst + (row_ptr[r] - first), rows[r],
			       (row_ptr[r + 1] - row_ptr[r]) * sizeof(T));
	});
}

// NBGPU_TRACE=1: wall-clock of the import steps on stderr (tuning aid)
struct Trace {
	bool on = getenv("NBGPU_TRACE") != nullptr;
	std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	void lap(const char *what)
	{
		if (!on)
			return;
		auto t1 = std::chrono::steady_clock::now();
		fprintf(stderr, "[nbgpu] %-18s %8.3f ms\n", what,
			std::chrono::duration<double, std::milli>(t1 - t0).count());
		t0 = t1;
	}
};

struct DeviceTemp {
	void *p = nullptr;
	~DeviceTemp()
	{
		if (p)
			nbgpu::dfree(p);
	}
	int alloc(size_t bytes)
	{
		cudaError_t e = nbgpu::dmalloc(&p, bytes ? bytes : 1);
		if (e != cudaSuccess) {
			set_error("nbgpu::dmalloc(%zu): %s", bytes, cudaGetErrorString(e));
			cudaGetLastError();
			p = nullptr;
			return NBGPU_ERR_NOMEM;
		}
		return NBGPU_OK;
	}
};

int build_layout(nbgpu_matrix_t *A, uint32_t N, const uint32_t *rows_size)
{
	A->N = N;
	if (A->n_cols == 0)
		A->n_cols = N;
	A->h_rows_size.assign(rows_size, rows_size + N);
	A->h_row_ptr.resize((size_t)N + 1);
	A->h_row_ptr[0] = 0;
	for (uint32_t i = 0; i < N; i++)
		A->h_row_ptr[i + 1] = A->h_row_ptr[i] + rows_size[i];
	A->nnz = A->h_row_ptr[N];
	A->n_slices = (N + kSliceRows - 1) / kSliceRows;
	std::vector<uint32_t> off((size_t)A->n_slices + 1), width((size_t)A->n_slices);
	const size_t n_pos = (size_t)A->n_slices * kSliceRows;
	auto slice_widths = [&](const uint32_t *perm, uint64_t *units_out) {
		uint64_t units = 0;
		uint32_t wmax = 0;
#pragma omp parallel for schedule(static) reduction(+ : units) reduction(max : wmax)
		for (int64_t s = 0; s < (int64_t)A->n_slices; s++) {
			uint32_t w = 0;
			for (uint32_t l = 0; l < kSliceRows; l++) {
				const size_t pos = (size_t)s * kSliceRows + l;
				const uint32_t r = perm ? perm[pos] : (uint32_t)pos;
				if (pos < n_pos && r < N)
					w = std::max(w, rows_size[r]);
			}
			width[s] = w;
			wmax = std::max(wmax, w);
			units += w;
		}
		A->max_width = wmax;
		*units_out = units;
	};
	uint64_t units = 0;
	slice_widths(nullptr, &units);
	// SELL-C-sigma: worth it when the identity order pads more than 5 % (irregular meshes).
	// Rank-local blocks keep the identity order (their visit order is planned on plain row ids).
	A->sigma = 1;
	{
		const char *env = getenv("NBGPU_SIGMA");
		uint32_t want = env ? (uint32_t)atoi(env) : 0;
		if (!env && !A->local_block && A->nnz > 0 && units * kSliceRows > A->nnz + A->nnz / 20)
			want = 256;
		want = (want / kSliceRows) * kSliceRows;
		if (want > kSliceRows - 1 && !A->local_block && N > 1)
			A->sigma = want;
	}
	std::vector<uint32_t> perm;
	if (A->sigma > 1) {
		perm.assign(n_pos, 0xFFFFFFFFu);
		bool pairs = (N % 2) == 0;
		for (uint32_t i = 0; pairs && i + 1 < N; i += 2)
			pairs = rows_size[i] == rows_size[i + 1];
		const uint32_t unit = pairs ? 2 : 1;
		std::vector<uint32_t> ids;
		for (uint32_t w0 = 0; w0 < N; w0 += A->sigma) {
			const uint32_t w1 = (uint32_t)std::min<uint64_t>(N, (uint64_t)w0 + A->sigma);
			ids.clear();
			for (uint32_t r = w0; r < w1; r += unit)
				ids.push_back(r);
			std::stable_sort(ids.begin(), ids.end(),
					 [&](uint32_t a, uint32_t b) { return rows_size[a] > rows_size[b]; });
			uint32_t pos = w0;
			for (uint32_t r : ids)
				for (uint32_t u = 0; u < unit && r + u < w1; u++)
					perm[pos++] = r + u;
		}
		uint64_t sorted_units = 0;
		slice_widths(perm.data(), &sorted_units);
		if (sorted_units >= units) {   // nothing gained: stay with the identity
			A->sigma = 1;
			perm.clear();
			slice_widths(nullptr, &units);
		} else {
			units = sorted_units;
		}
	}
	// near-uniform rows (structured meshes): store every slice max_width wide
	const uint64_t uniform_units = (uint64_t)A->n_slices * A->max_width;
	A->uniform_width = 0;
	if (!getenv("NBGPU_NO_UNIFORM") && A->n_slices > 0 && uniform_units <= units + units / 100) {
		A->uniform_width = A->max_width;
		std::fill(width.begin(), width.end(), A->max_width);
	}
	units = 0;
	for (uint32_t s = 0; s < A->n_slices; s++) {
		off[s] = (uint32_t)units;
		units += width[s];
		if (units > 0xFFFFFFFFull) {
			set_error("matrix too large for 32-bit slice offsets");
			return NBGPU_ERR_ARG;
		}
	}
	off[A->n_slices] = (uint32_t)units;
	A->stored = units * kSliceRows;
	NB_CUDA(nbgpu::dmalloc(&A->d_slice_off, off.size() * sizeof(uint32_t)));
	NB_CUDA(cudaMemcpyAsync(A->d_slice_off, off.data(), off.size() * sizeof(uint32_t),
				cudaMemcpyHostToDevice, ctx().stream));
	if (A->sigma > 1) {
		std::vector<uint32_t> inv(N);
		for (size_t pos = 0; pos < n_pos; pos++)
			if (perm[pos] < N)
				inv[perm[pos]] = (uint32_t)pos;
		NB_CUDA(nbgpu::dmalloc(&A->d_perm, n_pos * sizeof(uint32_t)));
		NB_CUDA(nbgpu::dmalloc(&A->d_inv_perm, (size_t)N * sizeof(uint32_t)));
		NB_CUDA(cudaMemcpy(A->d_perm, perm.data(), n_pos * sizeof(uint32_t), cudaMemcpyHostToDevice));
		NB_CUDA(cudaMemcpy(A->d_inv_perm, inv.data(), (size_t)N * sizeof(uint32_t), cudaMemcpyHostToDevice));
	}
	NB_CUDA(cudaStreamSynchronize(ctx().stream));   // `off` goes out of scope
	cudaError_t e = nbgpu::dmalloc(&A->d_val, std::max<size_t>(1, A->stored) * sizeof(double));
	if (e == cudaSuccess)
		e = nbgpu::dmalloc(&A->d_col, std::max<size_t>(1, A->stored) * sizeof(uint32_t));
	if (e != cudaSuccess) {
		set_error("matrix of %llu stored entries: %s", (unsigned long long)A->stored,
			  cudaGetErrorString(e));
		cudaGetLastError();
		return NBGPU_ERR_NOMEM;
	}
	return NBGPU_OK;
}

// Common tail of the two constructors / value setters: CSR arrays already in
// device temporaries (d_cols may be null = keep pattern, d_vals may be null =
// zeros) -> SELL.
int convert_in(nbgpu_matrix_t *A, const uint32_t *d_cols, const double *d_vals)
{
	Context &c = ctx();
	NB_TRY(ensure_host_pattern(A));
	DeviceTemp rp, bad;
	NB_TRY(rp.alloc(A->h_row_ptr.size() * sizeof(uint64_t)));
	NB_TRY(bad.alloc(sizeof(int)));
	NB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), c.stream));
	NB_CUDA(cudaMemcpyAsync(rp.p, A->h_row_ptr.data(), A->h_row_ptr.size() * sizeof(uint64_t),
				cudaMemcpyHostToDevice, c.stream));
	NB_CUDA(cudaStreamSynchronize(c.copy_stream));   // staged uploads have landed
	if (A->n_slices) {
		csr_to_sell_kernel<<<grid_for_slices(A->n_slices), kBlock, 0, c.stream>>>(
			A->N, A->n_cols, true, A->n_slices, (const uint64_t *)rp.p, d_cols, d_vals, A->d_slice_off,
			A->d_perm, d_cols ? A->d_col : nullptr, A->d_val, (int *)bad.p);
		NB_LAUNCHED();
	}
	int h_bad = 0;
	NB_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
	NB_CUDA(cudaStreamSynchronize(c.stream));
	if (h_bad) {
		set_error("row columns must be strictly ascending and < N (nb_sparse_create invariant)");
		return NBGPU_ERR_ARG;
	}
	if (d_cols && A->n_slices && (A->N & 1u) == 0 && !getenv("NBGPU_NO_BLOCKED")) {
		// pattern (re)built: look for the 2-dofs-per-node block structure
		NB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), c.stream));
		check_blocked_kernel<<<grid_for_slices(A->n_slices), kBlock, 0, c.stream>>>(
			A->N, A->n_slices, A->d_slice_off, A->d_col, (int *)bad.p);
		NB_LAUNCHED();
		NB_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
		NB_CUDA(cudaStreamSynchronize(c.stream));
		if (!h_bad) {
			NB_CUDA(nbgpu::dmalloc(&A->d_bcol, std::max<size_t>(1, A->stored / 4) * sizeof(uint32_t)));
			build_bcol_kernel<<<grid_for_slices(A->n_slices), kBlock, 0, c.stream>>>(
				A->n_slices, A->d_slice_off, A->d_col, A->d_bcol);
			NB_LAUNCHED();
			A->blocked = true;
		}
	}
	if (d_cols && A->n_slices && !getenv("NBGPU_NO_IDX16")) {
		const size_t n_ids = A->blocked ? A->stored / 4 : A->stored;
		NB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), c.stream));
		NB_CUDA(nbgpu::dmalloc(&A->d_idx16, std::max<size_t>(1, n_ids) * sizeof(short)));
		build_idx16_kernel<<<grid_for_slices(A->n_slices), kBlock, 0, c.stream>>>(
			A->blocked, A->N, A->col_shift, A->n_slices, A->d_slice_off, A->d_perm, A->blocked ? A->d_bcol : A->d_col,
			A->d_idx16, (int *)bad.p);
		NB_LAUNCHED();
		NB_CUDA(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
		NB_CUDA(cudaStreamSynchronize(c.stream));
		A->idx16 = !h_bad;
		if (h_bad) {
			nbgpu::dfree(A->d_idx16);
			A->d_idx16 = nullptr;
		}
	}
	return NBGPU_OK;
}

int convert_out(const nbgpu_matrix_t *A, uint32_t *d_cols, double *d_vals)
{
	Context &c = ctx();
	NB_TRY(ensure_host_pattern(const_cast<nbgpu_matrix_t *>(A)));
	DeviceTemp rp;
	NB_TRY(rp.alloc(A->h_row_ptr.size() * sizeof(uint64_t)));
	NB_CUDA(cudaMemcpyAsync(rp.p, A->h_row_ptr.data(), A->h_row_ptr.size() * sizeof(uint64_t),
				cudaMemcpyHostToDevice, c.stream));
	if (A->n_slices) {
		sell_to_csr_kernel<<<grid_for_slices(A->n_slices), kBlock, 0, c.stream>>>(
			A->N, A->n_slices, (const uint64_t *)rp.p, A->d_slice_off, A->d_perm, A->d_col, A->d_val,
			d_cols, d_vals);
		NB_LAUNCHED();
	}
	NB_CUDA(cudaStreamSynchronize(c.stream));
	return NBGPU_OK;
}

}  // namespace

namespace nbgpu {

int ensure_host_pattern(nbgpu_matrix_s *A)
{
	if (!A->d_node_counts || A->h_row_ptr.size() == (size_t)A->N + 1)
		return NBGPU_OK;
	std::vector<uint32_t> counts(A->N / 2);
	NB_CUDA(cudaStreamSynchronize(ctx().stream));
	NB_CUDA(cudaMemcpy(counts.data(), A->d_node_counts, counts.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	A->h_rows_size.resize(A->N);
	A->h_row_ptr.resize((size_t)A->N + 1);
	A->h_row_ptr[0] = 0;
	for (uint32_t i = 0; i < A->N; i++) {
		A->h_rows_size[i] = 2 * counts[i >> 1];
		A->h_row_ptr[i + 1] = A->h_row_ptr[i] + A->h_rows_size[i];
	}
	return NBGPU_OK;
}

// Host vector -> device through the pinned staging pair (threads pack, DMA runs from pinned memory):
// a pageable cudaMemcpy of the solver's 8 MB vectors runs at 6 GB/s, this at PCIe speed.
int upload_vector(double *d_dst, const double *src, size_t n)
{
	NB_TRY(upload_flat<double>(d_dst, src, n));
	NB_CUDA(cudaStreamSynchronize(ctx().copy_stream));
	return NBGPU_OK;
}

// Device vector -> host the same way; the producer must have finished on the library's stream.
int download_vector(double *dst, const double *d_src, size_t n)
{
	Context &c = ctx();
	NB_TRY(ensure_stage(kStageBytes));
	NB_CUDA(cudaStreamSynchronize(c.stream));
	const size_t chunk = kStageBytes / sizeof(double);
	int buf = 0;
	size_t pending_off[2] = {0, 0}, pending_n[2] = {0, 0};
	auto drain = [&](int b) {
		if (!pending_n[b])
			return cudaSuccess;
		cudaError_t e = cudaEventSynchronize(c.ev_stage[b]);
		if (e != cudaSuccess)
			return e;
		const double *from = (const double *)c.stage[b];
		double *to = dst + pending_off[b];
		const size_t cnt = pending_n[b], piece = size_t(1) << 17;
#pragma omp parallel for schedule(static)
		for (int64_t p = 0; p < (int64_t)((cnt + piece - 1) / piece); p++) {
			const size_t lo = (size_t)p * piece, hi = std::min(cnt, lo + piece);
			memcpy(to + lo, from + lo, (hi - lo) * sizeof(double));
		}
		pending_n[b] = 0;
		return cudaSuccess;
	};
	for (size_t off = 0; off < n; off += chunk) {
		const size_t cnt = std::min(chunk, n - off);
		NB_CUDA(drain(buf));
		NB_CUDA(cudaMemcpyAsync(c.stage[buf], d_src + off, cnt * sizeof(double), cudaMemcpyDeviceToHost,
					c.copy_stream));
		NB_CUDA(cudaEventRecord(c.ev_stage[buf], c.copy_stream));
		pending_off[buf] = off;
		pending_n[buf] = cnt;
		buf ^= 1;
	}
	NB_CUDA(drain(0));
	NB_CUDA(drain(1));
	return NBGPU_OK;
}

}  // namespace nbgpu

extern "C" {

int nbgpu_matrix_destroy(nbgpu_matrix_t *A)
{
	if (!A)
		return NBGPU_OK;
	if (ctx().ready) {
		cudaSetDevice(ctx().device);
		cudaStreamSynchronize(ctx().stream);
		nbgpu::dfree(A->d_slice_off);
		nbgpu::dfree(A->d_val);
		nbgpu::dfree(A->d_col);
		nbgpu::dfree(A->d_bcol);
		nbgpu::dfree(A->d_idx16);
		nbgpu::dfree(A->d_perm);
		nbgpu::dfree(A->d_inv_perm);
		nbgpu::dfree(A->d_node_counts);
	}
	delete A;
	return NBGPU_OK;
}

int nbgpu_matrix_create_from_csr(uint32_t N, const uint32_t *rows_size, const uint32_t *cols,
				 const double *vals, nbgpu_matrix_t **out)
{
	NB_INIT();
	NB_ARG(out != nullptr && (N == 0 || (rows_size != nullptr && cols != nullptr)));
	nbgpu_matrix_t *A = new nbgpu_matrix_t();
	int st = build_layout(A, N, rows_size);
	DeviceTemp dc, dv;
	if (st == NBGPU_OK)
		st = dc.alloc(A->nnz * sizeof(uint32_t));
	if (st == NBGPU_OK && vals)
		st = dv.alloc(A->nnz * sizeof(double));
	if (st == NBGPU_OK)
		st = upload_flat<uint32_t>((uint32_t *)dc.p, cols, A->nnz);
	if (st == NBGPU_OK && vals)
		st = upload_flat<double>((double *)dv.p, vals, A->nnz);
	if (st == NBGPU_OK)
		st = convert_in(A, (const uint32_t *)dc.p, vals ? (const double *)dv.p : nullptr);
	if (st != NBGPU_OK) {
		nbgpu_matrix_destroy(A);
		return st;
	}
	*out = A;
	return NBGPU_OK;
}

int nbgpu_matrix_create_local(uint32_t N_rows, uint32_t N_cols, uint32_t col_shift, const uint32_t *rows_size,
			      const uint32_t *cols_local, const double *vals, nbgpu_matrix_t **out)
{
	NB_INIT();
	NB_ARG(out != nullptr && (uint64_t)col_shift + N_rows <= N_cols && (col_shift & 1u) == 0 &&
	       (N_rows == 0 || (rows_size != nullptr && cols_local != nullptr)));
	nbgpu_matrix_t *A = new nbgpu_matrix_t();
	A->n_cols = N_cols;
	A->col_shift = col_shift;
	A->local_block = true;
	int st = build_layout(A, N_rows, rows_size);
	DeviceTemp dc, dv;
	if (st == NBGPU_OK)
		st = dc.alloc(A->nnz * sizeof(uint32_t));
	if (st == NBGPU_OK && vals)
		st = dv.alloc(A->nnz * sizeof(double));
	if (st == NBGPU_OK)
		st = upload_flat<uint32_t>((uint32_t *)dc.p, cols_local, A->nnz);
	if (st == NBGPU_OK && vals)
		st = upload_flat<double>((double *)dv.p, vals, A->nnz);
	if (st == NBGPU_OK)
		st = convert_in(A, (const uint32_t *)dc.p, vals ? (const double *)dv.p : nullptr);
	if (st != NBGPU_OK) {
		nbgpu_matrix_destroy(A);
		return st;
	}
	*out = A;
	return NBGPU_OK;
}

int nbgpu_matrix_create_from_rows(uint32_t N, const uint32_t *rows_size, uint32_t *const *rows_index,
				  double *const *rows_values, nbgpu_matrix_t **out)
{
	NB_INIT();
	NB_ARG(out != nullptr && (N == 0 || (rows_size != nullptr && rows_index != nullptr)));
	nbgpu_matrix_t *A = new nbgpu_matrix_t();
	Trace tr;
	int st = build_layout(A, N, rows_size);
	tr.lap("layout");
	DeviceTemp dc, dv;
	if (st == NBGPU_OK)
		st = dc.alloc(A->nnz * sizeof(uint32_t));
	if (st == NBGPU_OK && rows_values)
		st = dv.alloc(A->nnz * sizeof(double));
	tr.lap("temp alloc");
	if (st == NBGPU_OK)
		st = upload_rows<uint32_t>((uint32_t *)dc.p, rows_index, A->h_row_ptr);
	tr.lap("upload cols");
	if (st == NBGPU_OK && rows_values)
		st = upload_rows<double>((double *)dv.p, rows_values, A->h_row_ptr);
	tr.lap("upload vals");
	if (st == NBGPU_OK)
		st = convert_in(A, (const uint32_t *)dc.p, rows_values ? (const double *)dv.p : nullptr);
	tr.lap("convert");
	if (st != NBGPU_OK) {
		nbgpu_matrix_destroy(A);
		return st;
	}
	*out = A;
	return NBGPU_OK;
}

int nbgpu_matrix_info(const nbgpu_matrix_t *A, uint32_t *N, uint64_t *nnz, uint32_t *n_slices,
		      uint64_t *stored_entries)
{
	NB_ARG(A != nullptr);
	if (N)
		*N = A->N;
	if (nnz)
		*nnz = A->nnz;
	if (n_slices)
		*n_slices = A->n_slices;
	if (stored_entries)
		*stored_entries = A->stored;
	return NBGPU_OK;
}

int nbgpu_matrix_layout(const nbgpu_matrix_t *A, uint32_t *sigma, uint32_t *uniform_width,
			uint32_t *max_width, int *blocked, int *idx16)
{
	NB_ARG(A != nullptr);
	if (sigma)
		*sigma = A->sigma;
	if (uniform_width)
		*uniform_width = A->uniform_width;
	if (max_width)
		*max_width = A->max_width;
	if (blocked)
		*blocked = A->blocked ? 1 : 0;
	if (idx16)
		*idx16 = A->idx16 ? 1 : 0;
	return NBGPU_OK;
}

int nbgpu_matrix_set_values_csr(nbgpu_matrix_t *A, const double *vals)
{
	NB_INIT();
	NB_ARG(A != nullptr && vals != nullptr);
	DeviceTemp dv;
	NB_TRY(dv.alloc(A->nnz * sizeof(double)));
	NB_TRY(upload_flat<double>((double *)dv.p, vals, A->nnz));
	return convert_in(A, nullptr, (const double *)dv.p);
}

int nbgpu_matrix_set_values_rows(nbgpu_matrix_t *A, double *const *rows_values)
{
	NB_INIT();
	NB_ARG(A != nullptr && rows_values != nullptr);
	NB_TRY(ensure_host_pattern(A));
	DeviceTemp dv;
	NB_TRY(dv.alloc(A->nnz * sizeof(double)));
	NB_TRY(upload_rows<double>((double *)dv.p, rows_values, A->h_row_ptr));
	return convert_in(A, nullptr, (const double *)dv.p);
}

int nbgpu_matrix_get_values_csr(const nbgpu_matrix_t *A, double *vals)
{
	NB_INIT();
	NB_ARG(A != nullptr && vals != nullptr);
	DeviceTemp dv;
	NB_TRY(dv.alloc(A->nnz * sizeof(double)));
	NB_TRY(convert_out(A, nullptr, (double *)dv.p));
	NB_CUDA(cudaMemcpy(vals, dv.p, A->nnz * sizeof(double), cudaMemcpyDeviceToHost));
	return NBGPU_OK;
}

int nbgpu_matrix_get_values_rows(const nbgpu_matrix_t *A, double *const *rows_values)
{
	NB_INIT();
	NB_ARG(A != nullptr && rows_values != nullptr);
	std::vector<double> flat(A->nnz);
	NB_TRY(nbgpu_matrix_get_values_csr(A, flat.data()));
#pragma omp parallel for schedule(static)
	for (int64_t r = 0; r < (int64_t)A->N; r++)
		memcpy(rows_values[r], flat.data() + A->h_row_ptr[r],
		       (A->h_row_ptr[r + 1] - A->h_row_ptr[r]) * sizeof(double));
	return NBGPU_OK;
}

int nbgpu_matrix_get_pattern_csr(const nbgpu_matrix_t *A, uint32_t *rows_size, uint32_t *cols)
{
	NB_INIT();
	NB_ARG(A != nullptr);
	NB_TRY(ensure_host_pattern(const_cast<nbgpu_matrix_t *>(A)));
	if (rows_size)
		memcpy(rows_size, A->h_rows_size.data(), (size_t)A->N * sizeof(uint32_t));
	if (cols) {
		DeviceTemp dc;
		NB_TRY(dc.alloc(A->nnz * sizeof(uint32_t)));
		NB_TRY(convert_out(A, (uint32_t *)dc.p, nullptr));
		NB_CUDA(cudaMemcpy(cols, dc.p, A->nnz * sizeof(uint32_t), cudaMemcpyDeviceToHost));
	}
	return NBGPU_OK;
}

int nbgpu_matrix_reset(nbgpu_matrix_t *A)
{
	NB_INIT();
	NB_ARG(A != nullptr);
	NB_CUDA(cudaMemsetAsync(A->d_val, 0, A->stored * sizeof(double), ctx().stream));
	return NBGPU_OK;
}

}  // extern "C"
