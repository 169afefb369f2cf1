// common.cuh -- process-wide context, error handling and device helpers shared
// by every translation unit of libnbgpu.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "nbgpu.h"

namespace nbgpu {

constexpr uint32_t kSliceRows = 32;           // SELL slice height = one warp
constexpr uint32_t kPadCol = 0xFFFFFFFFu;     // column id of a padding entry
// Threads per CTA of every streaming kernel.  320 = 10 warps: the streamed SpMV is bound by the per-warp
// chain "ids -> gathers -> ordered accumulation", so warps per SM are what counts; two CTAs of 10 warps
// (94 registers, two 4.9 KB stages per warp = 196 KB of shared memory) are the most that fit an SM.
// Measured against 256 (8 warps, with the early stage release that 122 registers then allow):
// Q1 47.0 -> 46.0 us/iteration, L4096 544 -> 515, ragged 1 M-dof mesh 47.7 -> 46.1, Q4 158.8 -> 161.5.
#ifndef NB_BLOCK
#define NB_BLOCK 320
#endif
constexpr int kBlock = NB_BLOCK;
constexpr int kMaxPartialBlocks = 4096;       // upper bound on persistent grid sizes

struct Context {
	bool ready = false;
	int device = 0;
	int sm_count = 148;
	cudaStream_t stream = nullptr;        // all kernels
	cudaStream_t copy_stream = nullptr;   // staged uploads
	cudaEvent_t ev_a = nullptr, ev_b = nullptr;     // nbgpu_timer_*
	cudaEvent_t ev_stage[2] = {nullptr, nullptr};
	void *stage[2] = {nullptr, nullptr};  // pinned staging buffers
	size_t stage_bytes = 0;
	double *ws = nullptr;                 // Krylov work vectors (grow-only)
	size_t ws_bytes = 0;
	double *partials = nullptr;           // [2][kPartialRegion] reduction partials of the two producer kernels
	void *dev_state = nullptr;            // solver scalars (krylov.cu)
	void *host_state = nullptr;           // pinned mirror
	cudaEvent_t poll_ev[2] = {nullptr, nullptr};   // convergence polling (krylov.cu)
	uint64_t launches = 0;
	cudaMemPool_t pool = nullptr;         // retained allocation pool (null: plain cudaMalloc)
};

Context &ctx();
int ensure_init();
void set_error(const char *fmt, ...);
// pinned staging buffer pair of at least `bytes` each
int ensure_stage(size_t bytes);
int ensure_workspace(size_t bytes);

// host <-> device vectors through the pinned staging buffers (matrix.cu); both block until done
int upload_vector(double *d_dst, const double *src, size_t n);
int download_vector(double *dst, const double *d_src, size_t n);

// L2 residency control for a solve (context.cu)
bool l2_pin(void *ptr, size_t bytes, bool others_streaming);
void l2_unpin();

// Device memory from a RETAINED stream-ordered pool (cudaMallocAsync on the library's stream with the
// release threshold at maximum): the reference's call pattern re-imports the matrix on every solver call,
// and plain cudaMalloc / cudaFree of its 100+ MB blocks cost 1-160 ms each time (measured).  dmalloc
// synchronises the stream, so the block is usable at once from any stream or blocking copy; dfree is
// stream-ordered (everything that touched the block on another stream must have completed).
// nbgpu_finalize returns the pool to the driver.  NBGPU_NO_POOL=1 falls back to cudaMalloc/cudaFree.
cudaError_t dmalloc_bytes(void **p, size_t bytes);
cudaError_t dfree(void *p);
template <typename T>
inline cudaError_t dmalloc(T **p, size_t bytes)
{
	return dmalloc_bytes(reinterpret_cast<void **>(p), bytes);
}

}  // namespace nbgpu

#define NB_CUDA(expr)                                                         \
	do {                                                                  \
		cudaError_t nb_e_ = (expr);                                   \
		if (nb_e_ != cudaSuccess) {                                   \
			nbgpu::set_error("%s:%d: %s: %s", __FILE__, __LINE__, \
					 #expr, cudaGetErrorString(nb_e_));   \
			return NBGPU_ERR_CUDA;                                \
		}                                                             \
	} while (0)

#define NB_INIT()                                      \
	do {                                           \
		int nb_s_ = nbgpu::ensure_init();      \
		if (nb_s_ != NBGPU_OK)                 \
			return nb_s_;                  \
	} while (0)

#define NB_ARG(cond)                                                            \
	do {                                                                    \
		if (!(cond)) {                                                  \
			nbgpu::set_error("%s:%d: invalid argument: %s",         \
					 __FILE__, __LINE__, #cond);            \
			return NBGPU_ERR_ARG;                                   \
		}                                                               \
	} while (0)

// count + check a kernel launch
#define NB_LAUNCHED()                                 \
	do {                                          \
		nbgpu::ctx().launches++;              \
		NB_CUDA(cudaGetLastError());          \
	} while (0)

#define NB_TRY(expr)                        \
	do {                                \
		int nb_s_ = (expr);         \
		if (nb_s_ != NBGPU_OK)      \
			return nb_s_;       \
	} while (0)

// ------------------------------------------------------------------ device --
#ifdef __CUDACC__
namespace nbgpu {

// Programmatic dependent launch (PDL).  A kernel launched with the
// programmatic-stream-serialization attribute may start while its predecessor
// in the stream is still draining; it must not touch anything the predecessor
// produces before pdl_wait().  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait()
{
	asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_launch_dependents()
{
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		v += __shfl_down_sync(0xffffffffu, v, o);
	return v;
}

// Deterministic grid reduction of NV per-thread values.
//   1. warp shuffle tree, 2. one shared-memory pass per CTA, 3. the CTA that
//   takes the last ticket sums the per-CTA partials in a fixed order.
// Returns true (on every thread of that last CTA) with the totals in out[].
// partials: [NV][gridDim.x]; ticket: zero-initialised counter, self-resetting.
template <int NV>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double *partials,
					     unsigned int *ticket, double (&out)[NV])
{
	__shared__ double sm[NV][kBlock / 32];
	__shared__ bool is_last;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int k = 0; k < NV; k++) {
		double s = warp_sum(v[k]);
		if (lane == 0)
			sm[k][warp] = s;
	}
	__syncthreads();
	if (warp == 0) {
#pragma unroll
		for (int k = 0; k < NV; k++) {
			double s = (lane < kBlock / 32) ? sm[k][lane] : 0.0;
			s = warp_sum(s);
			if (lane == 0)
				partials[k * gridDim.x + blockIdx.x] = s;
		}
		if (lane == 0) {
			__threadfence();
			unsigned int t = atomicInc(ticket, gridDim.x - 1);
			is_last = (t == gridDim.x - 1);
		}
	}
	__syncthreads();
	if (!is_last)
		return false;
	__threadfence();
#pragma unroll
	for (int k = 0; k < NV; k++) {
		double s = 0.0;
		for (unsigned int b = threadIdx.x; b < gridDim.x; b += kBlock)
			s += __ldcg(partials + k * gridDim.x + b);
		s = warp_sum(s);
		if (lane == 0)
			sm[k][warp] = s;
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < NV; k++) {
		double s = 0.0;
#pragma unroll
		for (int w = 0; w < kBlock / 32; w++)
			s += sm[k][w];
		out[k] = s;
	}
	return true;
}

// Ticketless form of the same reduction, split over a kernel boundary (single GPU): the producer
// kernel's CTAs only store their partials (cta_store_partials); every CTA of the CONSUMER kernel
// sums them itself in a fixed order (cta_sum_partials), so all CTAs get bit-identical totals.  What
// the producer saves -- fence, ticket atomic, the last CTA re-reading and summing the partials, one
// more store -- sits on the critical path between two dependent kernels; what the consumer adds is a
// few coalesced L2 loads where it used to read one scalar (one round trip either way).
// Every consumer CTA reads ALL partials at the same moment: the few cache lines that hold them sit in a few L2
// slices, and several hundred CTAs queue there (measured on one GPU: 782 CTAs of K3 summing K2's 2 x 444
// partials 2.2 us, against 1.0 us for one L2 round trip).  The producer therefore stores kPartialReplicas
// copies (one lane each, fire and forget) and consumer CTA b reads copy b % kPartialReplicas.
#ifndef NB_PARTIAL_REPLICAS
#define NB_PARTIAL_REPLICAS 4
#endif
constexpr int kPartialReplicas = NB_PARTIAL_REPLICAS;
constexpr size_t kPartialReplicaStride = 3 * (size_t)kMaxPartialBlocks;              // doubles: [3][kMaxPartialBlocks]
constexpr size_t kPartialRegion = kPartialReplicas * kPartialReplicaStride;         // one producer kernel's copies

template <int NV>
__device__ __forceinline__ void cta_store_partials(double (&v)[NV], double *region)
{
	static_assert(NV <= 3 && NV * kPartialReplicas <= 32, "one lane per value and copy");
	__shared__ double sm[NV][kBlock / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int k = 0; k < NV; k++) {
		double s = warp_sum(v[k]);
		if (lane == 0)
			sm[k][warp] = s;
	}
	__syncthreads();
	if (warp == 0) {
		double mine = 0.0;   // lane l stores value l % NV into copy l / NV
#pragma unroll
		for (int k = 0; k < NV; k++) {
			double s = (lane < kBlock / 32) ? sm[k][lane] : 0.0;
			s = __shfl_sync(0xffffffffu, warp_sum(s), 0);
			if (lane % NV == k)
				mine = s;
		}
		if (lane < NV * kPartialReplicas)
			region[(lane / NV) * kPartialReplicaStride + (lane % NV) * gridDim.x + blockIdx.x] = mine;
	}
}

// region: as written by a grid of n_ctas CTAs; same summation tree as grid_reduce's last CTA
template <int NV>
__device__ __forceinline__ void cta_sum_partials(const double *region, uint32_t n_ctas, double (&out)[NV])
{
	__shared__ double sm[NV][kBlock / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const double *partials = region + (blockIdx.x % kPartialReplicas) * kPartialReplicaStride;
#pragma unroll
	for (int k = 0; k < NV; k++) {
		double s = 0.0;
		for (unsigned int b = threadIdx.x; b < n_ctas; b += kBlock)
			s += __ldcg(partials + k * n_ctas + b);
		s = warp_sum(s);
		if (lane == 0)
			sm[k][warp] = s;
	}
	__syncthreads();
#pragma unroll
	for (int k = 0; k < NV; k++) {
		double s = 0.0;
#pragma unroll
		for (int w = 0; w < kBlock / 32; w++)
			s += sm[k][w];
		out[k] = s;
	}
}

}  // namespace nbgpu
#endif
