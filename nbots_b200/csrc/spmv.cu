// spmv.cu -- nb_sparse_multiply_vector (sources/nb/solver_bot/sparse/sparse.c:405-414).
#include <algorithm>

#include "sell_stream.cuh"

using namespace nbgpu;

namespace nbgpu {

constexpr int kSpmvUnroll = 6;

__global__ void __launch_bounds__(kBlock, 4)
spmv_sell_kernel(uint32_t N, uint32_t n_slices, const uint32_t *__restrict__ slice_off,
		 const uint32_t *__restrict__ perm, const double *__restrict__ val, const uint32_t *__restrict__ col,
		 const double *__restrict__ x, double *__restrict__ y)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slices; s += warps) {
		const uint32_t row = perm ? __ldg(perm + (size_t)s * kSliceRows + lane) : s * kSliceRows + lane;
		const uint32_t off = __ldg(slice_off + s);
		const uint32_t width = __ldg(slice_off + s + 1) - off;
		const double acc = sell_row_times<kSpmvUnroll, false>(val, col, off, width, lane, row,
								      0u, x, nullptr);
		if (row < N)
			y[row] = acc;
	}
}

template <int LAYOUT>   // blocked | idx16 << 1
__global__ void __launch_bounds__(kBlock, kStreamCtas)
spmv_stream_kernel(SellView A, StreamConfig cfg, const double *__restrict__ x, double *__restrict__ y)
{
	extern __shared__ __align__(128) unsigned char smem[];
	sell_stream_rows<(LAYOUT & 1) != 0, false, (LAYOUT & 2) != 0, false>(
		A, x, cfg, smem, [] { return true; }, [] { return true; }, NoPre(),
		[&](uint32_t row, double acc, double, double, double) {
			if (row < A.N)
				y[row] = acc;
		});
}

// Ring depth and CTAs per SM for the streamed kernels of matrix A.  Shared memory
// per CTA = 8 warps x stages x stage_bytes; two CTAs per SM (16 consumer warps)
// are preferred, one is accepted for wide slices.  Env overrides for tuning:
// NBGPU_STREAM_STAGES, NBGPU_STREAM_CTAS; NBGPU_SPMV_PATH=reg disables the path.
bool stream_config(const nbgpu_matrix_s *A, const void *kernel, StreamConfig *cfg, int block)
{
	const char *path = getenv("NBGPU_SPMV_PATH");
	if (path && path[0] == 'r')
		return false;
	if (A->max_width == 0)
		return false;
	const uint32_t cap = (A->max_width + 1u) & ~1u;
	const uint32_t stage_bytes = stream_stage_bytes(cap, A->blocked, A->idx16);
	const uint32_t fixed = kStreamWarps * kStreamMaxStages * (sizeof(uint64_t) + sizeof(uint2));
	int smem_optin = 0;
	cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx().device);
	int want_ctas = getenv("NBGPU_STREAM_CTAS") ? atoi(getenv("NBGPU_STREAM_CTAS")) : kStreamCtas;
	int want_stages = getenv("NBGPU_STREAM_STAGES") ? atoi(getenv("NBGPU_STREAM_STAGES")) : 0;
	for (int ctas = want_ctas; ctas >= 1; ctas--) {
		// 228 KB per SM, 1 KB reserved per CTA, some static shared memory for the reductions
		const int64_t budget = std::min<int64_t>(smem_optin, (228 * 1024) / ctas - 2048);
		int64_t stages = (budget - fixed) / ((int64_t)kStreamWarps * stage_bytes);
		stages = std::min<int64_t>(stages, kStreamMaxStages);
		if (want_stages > 0)
			stages = std::min<int64_t>(stages, want_stages);
		if (stages < 2)
			continue;
		cfg->cap = cap;
		cfg->stages = (uint32_t)stages;
		cfg->stage_bytes = stage_bytes;
		cfg->smem_bytes = (uint32_t)(kStreamWarps * stages * stage_bytes + fixed);
		if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem_bytes) !=
		    cudaSuccess) {
			cudaGetLastError();
			continue;
		}
		int per_sm = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, cfg->smem_bytes) !=
			    cudaSuccess ||
		    per_sm < 1) {
			cudaGetLastError();
			continue;
		}
		per_sm = std::min(per_sm, ctas);
		const int64_t want = ((int64_t)A->n_slices + kStreamWarps - 1) / kStreamWarps;
		const int64_t cap_grid = std::min<int64_t>((int64_t)ctx().sm_count * per_sm, kMaxPartialBlocks);
		cfg->grid = (int)std::max<int64_t>(1, std::min(want, cap_grid));
		return true;
	}
	return false;
}

int spmv_grid(uint32_t n_slices)
{
	static int blocks_per_sm = 0;
	if (!blocks_per_sm) {
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, spmv_sell_kernel, kBlock,
								  0) != cudaSuccess ||
		    blocks_per_sm < 1)
			blocks_per_sm = 4;
	}
	int64_t want = ((int64_t)n_slices * 32 + kBlock - 1) / kBlock;
	int64_t cap = std::min<int64_t>((int64_t)ctx().sm_count * blocks_per_sm, kMaxPartialBlocks);
	return (int)std::max<int64_t>(1, std::min(want, cap));
}

}  // namespace nbgpu

extern "C" {

int nbgpu_spmv(const nbgpu_matrix_t *A, const double *d_in, double *d_out)
{
	NB_INIT();
	NB_ARG(A != nullptr && d_in != nullptr && d_out != nullptr);
	NB_ARG(d_in != d_out);   /* the reference accumulates into out[]: no aliasing (sparse.c:410) */
	NB_ARG(!A->local_block);   /* rank-local blocks go through nbgpu_dist_spmv */
	/* the blocked layout gathers x[2c], x[2c+1] as one 16-byte load */
	NB_ARG(!A->blocked || ((uintptr_t)d_in & 15) == 0);
	if (A->N == 0)
		return NBGPU_OK;
	StreamConfig cfg;
	const void *kernel = by_layout(A->layout(), [](auto L) {
		return (const void *)spmv_stream_kernel<decltype(L)::value>;
	});
	if (stream_config(A, kernel, &cfg)) {
		SellView V;
		V.N = A->N; V.n_slices = A->n_slices; V.slice_off = A->d_slice_off; V.val = A->d_val;
		V.col = A->stream_ids();
		V.uniform_width = A->uniform_width;
		V.perm = A->d_perm;
		by_layout(A->layout(), [&](auto L) {
			spmv_stream_kernel<decltype(L)::value><<<cfg.grid, kBlock, cfg.smem_bytes, ctx().stream>>>(V, cfg, d_in, d_out);
			return 0;
		});
		NB_LAUNCHED();
		return NBGPU_OK;
	}
	spmv_sell_kernel<<<spmv_grid(A->n_slices), kBlock, 0, ctx().stream>>>(
		A->N, A->n_slices, A->d_slice_off, A->d_perm, A->d_val, A->d_col, d_in, d_out);
	NB_LAUNCHED();
	return NBGPU_OK;
}

int nbgpu_spmv_host(const nbgpu_matrix_t *A, const double *in, double *out)
{
	NB_INIT();
	NB_ARG(A != nullptr && in != nullptr && out != nullptr);
	const size_t bytes = (size_t)A->N * sizeof(double);
	NB_TRY(ensure_workspace(2 * bytes));
	double *d_in = ctx().ws, *d_out = ctx().ws + A->N;
	NB_CUDA(cudaMemcpyAsync(d_in, in, bytes, cudaMemcpyHostToDevice, ctx().stream));
	NB_TRY(nbgpu_spmv(A, d_in, d_out));
	NB_CUDA(cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, ctx().stream));
	NB_CUDA(cudaStreamSynchronize(ctx().stream));
	return NBGPU_OK;
}

}  // extern "C"
