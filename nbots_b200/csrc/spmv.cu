// spmv.cu -- nb_sparse_multiply_vector (sources/nb/solver_bot/sparse/sparse.c:405-414).
#include <algorithm>

#include "spmv.cuh"

using namespace nbgpu;

namespace nbgpu {

constexpr int kSpmvUnroll = 6;

__global__ void __launch_bounds__(kBlock, 4)
spmv_sell_kernel(uint32_t N, uint32_t n_slices, const uint32_t *__restrict__ slice_off,
		 const double *__restrict__ val, const uint32_t *__restrict__ col,
		 const double *__restrict__ x, double *__restrict__ y)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slices; s += warps) {
		const uint32_t row = s * kSliceRows + lane;
		const uint32_t off = __ldg(slice_off + s);
		const uint32_t width = __ldg(slice_off + s + 1) - off;
		const double acc = sell_row_times<kSpmvUnroll, false>(val, col, off, width, lane, row,
								      min(row, N - 1), x, nullptr);
		if (row < N)
			y[row] = acc;
	}
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
	if (A->N == 0)
		return NBGPU_OK;
	spmv_sell_kernel<<<spmv_grid(A->n_slices), kBlock, 0, ctx().stream>>>(
		A->N, A->n_slices, A->d_slice_off, A->d_val, A->d_col, d_in, d_out);
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
