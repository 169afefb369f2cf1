// krylov.cu -- Jacobi-preconditioned CG and plain CG on the device.
//
// Reference: nb_sparse_solve_CG_precond_Jacobi
//   (sources/nb/solver_bot/sparse/solvers/cg_precond_jacobi.c:13-90) and
//   nb_sparse_solve_conjugate_gradient (solvers/conjugate_gradient.c:13-77).
//
// One reference iteration is three OpenMP loops separated by two scalar
// reductions; here it is three kernels and the scalars never visit the host:
//
//   K1  spmv     w = A p,  pw = p.w          (reduction fused in the SpMV)
//   K2  update   g += a w, q = g/diag, gq' = g.q, gg' = g.g
//   K3  dir      x += a p, p = -q + b p      (x rides with p: p is read once per iteration)
//
// a = gq/pw and b = gq'/gq are recomputed by every thread from the reduced
// dots kept in a small state block in HBM.  The reference's pass 1 also
// recomputes g.g and g.q, but those are the sums K2 of the previous iteration
// (or the init kernel) already produced over the same vectors, so they are
// carried instead of re-read.  Reductions are deterministic (fixed persistent
// grid, fixed tree), so a solve is reproducible run to run.
//
// Stopping rule (cg_precond_jacobi.c:45,84-89): `while (gg > tol^2 && k <
// max_iter)` where gg is the value pass 1 computed, i.e. the residual of the
// iterate BEFORE the latest update.  Iteration k is therefore gated on
// |g_{k-1}|^2 (|g_0|^2 for k = 0) and tolerance_reached reports that same
// stale number.  The gate is evaluated on the device by K1(k); the host only
// enqueues chunks of iterations and polls a `done` flag one chunk behind, so
// the GPU never waits for the host.  Kernels of iterations enqueued past
// convergence return immediately.
//
// B200 specifics:
//   * K1 streams the matrix through shared memory with TMA bulk copies
//     (sell_stream.cuh); a register-path K1 is kept for matrices whose widest
//     slice does not fit a stage.
//   * All kernels are launched with programmatic dependent launch: the next
//     kernel's CTAs become resident while the current one runs, K1 already
//     requests its first matrix stages and K2/K3 preload the operands the
//     running kernel does not write, each before `griddepcontrol.wait`.  Every
//     kernel releases its dependents only AFTER its own wait, so a kernel's
//     pre-wait section overlaps its immediate predecessor only.
//   * When the work vectors fit the persisting-L2 carve-out they are pinned
//     there for the duration of the solve (access-policy window on the stream,
//     everything else -- the matrix -- marked streaming), so only the matrix
//     comes from HBM in steady state.  (With the matrix stream carrying its own
//     evict-first hint the window measures neutral today: 45.9 vs 46.0 us.)
//   * The scalars of the state block are read with plain cached loads, never
//     `volatile` / `ld.cg`: every thread of every kernel reads the same line.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "sell_stream.cuh"

using namespace nbgpu;

namespace {

struct KrylovState {
	double gg[3];        // |g_k|^2 in slot k % 3
	double gq[2];        // g_k . q_k in slot k & 1 (plain CG: same as gg)
	double pw;           // p_k . A p_k
	double tol2;
	double gg_final;     // the value the reference's loop test failed on
	uint32_t max_iter;
	uint32_t k_final;    // iterations performed
	int32_t done;
	unsigned int ticket;
};

constexpr int kIterUnroll = 6;
constexpr uint32_t kChunkIters = 32;

__device__ __forceinline__ uint32_t gate_slot(uint32_t k) { return k == 0 ? 0u : (k - 1) % 3u; }

// The loop test of iteration k (:45).  Returns true when the iteration runs;
// otherwise records the exit (once) and makes every later kernel a no-op.
__device__ __forceinline__ bool iteration_gate(uint32_t k, KrylovState *st)
{
	// All four loads are issued before the first use (one round trip at the head of the kernel instead
	// of two; ncu: 17 % of the SpMV kernel's stall samples sat on this chain), as PLAIN loads: a warp
	// coalesces them and the first warp of a CTA leaves the line in L1 for the others.  Every thread of
	// the grid reads the same 64 bytes, so `volatile` or L2-only (`ld.cg`) loads turn that line into a
	// hot spot (measured: K2/K3 +3 to +6 us).  The values were written by an earlier kernel of the
	// stream, or, for `done`, lead every CTA of this kernel to the same decision.
	const KrylovState *cs = st;
	const int32_t done = cs->done;
	const double gg = cs->gg[gate_slot(k)];
	const double tol2 = cs->tol2;
	const uint32_t max_iter = cs->max_iter;
	if (done)
		return false;
	if (gg > tol2 && k < max_iter)
		return true;
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		st->k_final = k;
		st->gg_final = gg;
		__threadfence();
		st->done = 1;
	}
	return false;
}

template <bool JACOBI>
__device__ __forceinline__ void init_row(uint32_t row, double acc, double d, const double *__restrict__ b,
					 double *__restrict__ g, double *__restrict__ p, double *__restrict__ q,
					 double *__restrict__ diag, double (&dots)[2])
{
	// g = A x - b, q = g / Aii, p = -q   (:33-43)
	const double gi = __dsub_rn(acc, b[row]);
	g[row] = gi;
	dots[0] = __dadd_rn(dots[0], __dmul_rn(gi, gi));
	if (JACOBI) {
		const double qi = __ddiv_rn(gi, d);
		diag[row] = d;
		q[row] = qi;
		p[row] = -qi;
		dots[1] = __dadd_rn(dots[1], __dmul_rn(gi, qi));
	} else {
		p[row] = -gi;
	}
}

template <bool JACOBI>
__device__ __forceinline__ void init_finish(double (&dots)[2], double *partials, KrylovState *st)
{
	if (!partials)
		return;   // reference-order mode: seq_dot_kernel follows
	double tot[2];
	if (grid_reduce<2>(dots, partials, &st->ticket, tot) && threadIdx.x == 0) {
		st->gg[0] = tot[0];
		st->gq[0] = JACOBI ? tot[1] : tot[0];
	}
}

__device__ __forceinline__ void spmv_finish(double (&dots)[1], double *partials, KrylovState *st)
{
	if (!partials)
		return;
	double tot[1];
	if (grid_reduce<1>(dots, partials, &st->ticket, tot) && threadIdx.x == 0)
		st->pw = tot[0];
}

// ---- register path (fallback for very wide slices) -----------------------------
template <bool JACOBI>
__global__ void __launch_bounds__(kBlock, 4)
krylov_init_kernel(uint32_t N, uint32_t n_slices, const uint32_t *__restrict__ slice_off,
		   const uint32_t *__restrict__ perm, const double *__restrict__ val, const uint32_t *__restrict__ col,
		   const double *__restrict__ b, const double *__restrict__ x, double *__restrict__ g,
		   double *__restrict__ p, double *__restrict__ q, double *__restrict__ diag,
		   double *partials, KrylovState *st)
{
	pdl_wait();
	pdl_launch_dependents();
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	double dots[2] = {0.0, 0.0};
	for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slices; s += warps) {
		const uint32_t row = perm ? __ldg(perm + (size_t)s * kSliceRows + lane) : s * kSliceRows + lane;
		const uint32_t off = __ldg(slice_off + s);
		const uint32_t width = __ldg(slice_off + s + 1) - off;
		double d = 0.0;
		const double acc = sell_row_times<kIterUnroll, JACOBI>(val, col, off, width, lane, row,
								       0u, x, &d);
		if (row < N)
			init_row<JACOBI>(row, acc, d, b, g, p, q, diag, dots);
	}
	init_finish<JACOBI>(dots, partials, st);
}

__global__ void __launch_bounds__(kBlock, 4)
krylov_spmv_kernel(uint32_t k, uint32_t N, uint32_t n_slices, const uint32_t *__restrict__ slice_off,
		   const uint32_t *__restrict__ perm, const double *__restrict__ val, const uint32_t *__restrict__ col,
		   const double *__restrict__ p, double *__restrict__ w, double *partials,
		   KrylovState *st)
{
	pdl_wait();
	pdl_launch_dependents();
	if (!iteration_gate(k, st))
		return;
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	double dots[1] = {0.0};
	for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slices; s += warps) {
		const uint32_t row = perm ? __ldg(perm + (size_t)s * kSliceRows + lane) : s * kSliceRows + lane;
		const uint32_t off = __ldg(slice_off + s);
		const uint32_t width = __ldg(slice_off + s + 1) - off;
		const double acc = sell_row_times<kIterUnroll, false>(val, col, off, width, lane, row,
								      0u, p, nullptr);
		if (row < N) {
			w[row] = acc;
			dots[0] = __dadd_rn(dots[0], __dmul_rn(__ldg(p + row), acc));
		}
	}
	spmv_finish(dots, partials, st);
}

// ---- streamed path: the matrix goes through shared memory (sell_stream.cuh) -----
template <bool JACOBI, int LAYOUT>   // LAYOUT = blocked | idx16 << 1
__global__ void __launch_bounds__(kBlock, kStreamCtas)
krylov_init_stream_kernel(SellView A, StreamConfig cfg, const double *__restrict__ b,
			  const double *__restrict__ x, double *__restrict__ g, double *__restrict__ p,
			  double *__restrict__ q, double *__restrict__ diag, double *partials, KrylovState *st)
{
	extern __shared__ __align__(128) unsigned char smem[];
	double dots[2] = {0.0, 0.0};
	sell_stream_rows<(LAYOUT & 1) != 0, JACOBI, (LAYOUT & 2) != 0>(
		A, x, cfg, smem, [] { return true; }, [] { return true; },
		[&](uint32_t row, double acc, double d, double) {
			if (row < A.N)
				init_row<JACOBI>(row, acc, d, b, g, p, q, diag, dots);
		});
	init_finish<JACOBI>(dots, partials, st);
}

template <int LAYOUT>
__global__ void __launch_bounds__(kBlock, kStreamCtas)
krylov_spmv_stream_kernel(uint32_t k, SellView A, StreamConfig cfg, const double *__restrict__ p,
			  double *__restrict__ w, double *partials, KrylovState *st)
{
	extern __shared__ __align__(128) unsigned char smem[];
	double dots[1] = {0.0};
	bool active = true;
	sell_stream_rows<(LAYOUT & 1) != 0, false, (LAYOUT & 2) != 0>(
		A, p, cfg, smem,
		[&] {
			// a function of (k, state) only: the whole grid takes the same branch
			active = iteration_gate(k, st);
			return active;
		},
		[] { return true; },
		[&](uint32_t row, double acc, double, double p_row) {
			if (row < A.N) {
				w[row] = acc;
				dots[0] = __dadd_rn(dots[0], __dmul_rn(p_row, acc));
			}
		});
	if (!active)
		return;
	spmv_finish(dots, partials, st);
}

// K2: g += a w, q = g / diag, gq' = g.q, gg' = g.g   (:64-68; x += a p is done by K3)
// Two elements per thread and trip, every load of a trip issued before the first
// use; the first trip's g, diag are loaded before the dependency wait (the
// SpMV kernel running ahead of us only writes w and the scalars).
template <bool JACOBI>
__global__ void __launch_bounds__(kBlock)
krylov_update_kernel(uint32_t k, uint32_t N, const double *__restrict__ w, const double *__restrict__ diag,
		     double *__restrict__ g, double *__restrict__ q, double *partials, KrylovState *st)
{
	const uint32_t stride = gridDim.x * blockDim.x;
	const uint32_t base = blockIdx.x * blockDim.x + threadIdx.x;
	double g0 = 0, d0 = 1, g1 = 0, d1 = 1;
	if (base < N) {
		const uint32_t j1 = base + stride < N ? base + stride : base;
		g0 = g[base];
		g1 = g[j1];
		if (JACOBI) {
			d0 = diag[base];
			d1 = diag[j1];
		}
	}
	pdl_wait();
	pdl_launch_dependents();
	const int32_t done = st->done;   // plain loads of the shared state line, see iteration_gate
	const double gq_k = st->gq[k & 1], pw_k = st->pw;
	if (done)
		return;
	const double alpha = __ddiv_rn(gq_k, pw_k);
	double dots[2] = {0.0, 0.0};
	for (uint32_t i0 = base; i0 < N; i0 += 2 * stride) {
		const uint32_t i1 = i0 + stride;
		const bool has1 = i1 < N;
		const uint32_t j1 = has1 ? i1 : i0;
		if (i0 != base) {
			g0 = g[i0];
			g1 = g[j1];
			if (JACOBI) {
				d0 = diag[i0];
				d1 = diag[j1];
			}
		}
		const double w0 = w[i0], w1 = w[j1];
		const double gn0 = __dadd_rn(g0, __dmul_rn(alpha, w0));
		const double gn1 = __dadd_rn(g1, __dmul_rn(alpha, w1));
		g[i0] = gn0;
		dots[0] = __dadd_rn(dots[0], __dmul_rn(gn0, gn0));
		if (JACOBI) {
			const double q0 = __ddiv_rn(gn0, d0);
			q[i0] = q0;
			dots[1] = __dadd_rn(dots[1], __dmul_rn(gn0, q0));
		}
		if (has1) {
			g[i1] = gn1;
			dots[0] = __dadd_rn(dots[0], __dmul_rn(gn1, gn1));
			if (JACOBI) {
				const double q1 = __ddiv_rn(gn1, d1);
				q[i1] = q1;
				dots[1] = __dadd_rn(dots[1], __dmul_rn(gn1, q1));
			}
		}
	}
	if (!partials)
		return;   // reference-order reductions are done by seq_dot_kernel
	double tot[2];
	if (grid_reduce<2>(dots, partials, &st->ticket, tot) && threadIdx.x == 0) {
		st->gg[(k + 1) % 3u] = tot[0];
		st->gq[(k + 1) & 1] = JACOBI ? tot[1] : tot[0];
	}
}

// K3: x += a p (:62-63, moved here from K2: p is read once per iteration instead of twice; same
// operations, same rounding), then p = -q + b p (:70-74); plain CG passes q == g.  p and x are
// preloaded before the dependency wait (the update kernel ahead of us writes neither).
__global__ void __launch_bounds__(kBlock)
krylov_dir_kernel(uint32_t k, uint32_t N, const double *__restrict__ q, double *__restrict__ p,
		  double *__restrict__ x, const KrylovState *st)
{
	const uint32_t stride = gridDim.x * blockDim.x;
	const uint32_t base = blockIdx.x * blockDim.x + threadIdx.x;
	double p0 = 0, p1 = 0, x0 = 0, x1 = 0;
	if (base < N) {
		const uint32_t j1 = base + stride < N ? base + stride : base;
		p0 = p[base];
		p1 = p[j1];
		x0 = x[base];
		x1 = x[j1];
	}
	pdl_wait();
	pdl_launch_dependents();
	const int32_t done = st->done;   // plain loads of the shared state line, see iteration_gate
	const double gq_k = st->gq[k & 1], gq_n = st->gq[(k + 1) & 1];
	const double pw_k = st->pw;
	if (done)
		return;
	const double alpha = __ddiv_rn(gq_k, pw_k);
	const double beta = __ddiv_rn(gq_n, gq_k);
	for (uint32_t i0 = base; i0 < N; i0 += 2 * stride) {
		const uint32_t i1 = i0 + stride;
		const bool has1 = i1 < N;
		const uint32_t j1 = has1 ? i1 : i0;
		if (i0 != base) {
			p0 = p[i0];
			p1 = p[j1];
			x0 = x[i0];
			x1 = x[j1];
		}
		const double q0 = q[i0], q1 = q[j1];
		x[i0] = __dadd_rn(x0, __dmul_rn(alpha, p0));
		p[i0] = __dadd_rn(-q0, __dmul_rn(beta, p0));
		if (has1) {
			x[i1] = __dadd_rn(x1, __dmul_rn(alpha, p1));
			p[i1] = __dadd_rn(-q1, __dmul_rn(beta, p1));
		}
	}
}

// Verification mode (nbgpu_set_reduction_order(1)): the dot products are summed
// by ONE thread in index order, exactly like the reference's single-threaded
// loops (the FEM driver passes omp_parallel_threads = 1,
// static_elasticity2D.c:90).  Every other operation of the solver already
// rounds like the reference, so in this mode the whole solve -- iterates,
// iteration count, tolerance_reached -- is bit-identical to the reference's.
// A warp loads 32 products at a time; lane 0 adds them in order.
__global__ void seq_dot_kernel(uint32_t N, const double *__restrict__ a1, const double *__restrict__ b1,
			       double *out1, const double *__restrict__ a2, const double *__restrict__ b2,
			       double *out2, const KrylovState *st)
{
	if (*(volatile const int32_t *)&st->done)
		return;
	const uint32_t lane = threadIdx.x;
	double s1 = 0.0, s2 = 0.0;
	for (uint32_t base = 0; base < N; base += 32) {
		const uint32_t i = base + lane;
		const double t1 = i < N ? __dmul_rn(a1[i], b1[i]) : 0.0;
		const double t2 = (a2 && i < N) ? __dmul_rn(a2[i], b2[i]) : 0.0;
		const uint32_t n = min(32u, N - base);
		for (uint32_t l = 0; l < n; l++) {
			const double u1 = __shfl_sync(0xffffffffu, t1, l);
			const double u2 = __shfl_sync(0xffffffffu, t2, l);
			s1 = __dadd_rn(s1, u1);
			s2 = __dadd_rn(s2, u2);
		}
	}
	if (lane == 0) {
		*out1 = s1;
		if (out2)
			*out2 = a2 ? s2 : s1;
	}
}

// ---- host ------------------------------------------------------------------------

// Persistent grid of one kernel: exactly the number of CTAs that are resident at
// once (SMs x occupancy), so the grid-stride loops run as a single full wave.
template <typename Kernel>
int resident_grid(Kernel kernel, int64_t want_blocks)
{
	int per_sm = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, 0) != cudaSuccess || per_sm < 1) {
		cudaGetLastError();
		per_sm = 1;
	}
	int64_t cap = std::min<int64_t>((int64_t)ctx().sm_count * per_sm, kMaxPartialBlocks);
	return (int)std::max<int64_t>(1, std::min(want_blocks, cap));
}

// kernel launch with (optionally) the programmatic-dependent-launch attribute
template <typename... KArgs, typename... Args>
cudaError_t launch(bool pdl, void (*kernel)(KArgs...), int grid, size_t smem, Args &&...args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)grid);
	cfg.blockDim = dim3(kBlock);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = ctx().stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = pdl ? 1 : 0;
	ctx().launches++;
	return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

bool g_seq_dots = false;
cudaEvent_t g_poll_ev[2] = {nullptr, nullptr};

// Optional per-kernel timing (nbgpu_krylov_profile): CUDA events around each of
// the three kernels for the first kProfIters iterations of a solve, on the
// stream the kernels run on.  Off by default; bench.py uses it in a separate
// solve to get the live duration of the dominant kernel.
constexpr uint32_t kProfIters = 256;
bool g_prof_on = false;
std::vector<cudaEvent_t> g_prof_ev;    // 4 events per iteration
uint32_t g_prof_recorded = 0;
double g_prof_ms[3] = {0, 0, 0};
uint32_t g_prof_n = 0;

int solve_impl(const nbgpu_matrix_t *A, const double *d_b, double *d_x, uint32_t max_iter, double tol,
	       uint32_t *niter, double *tol_reached, bool jacobi)
{
	Context &c = ctx();
	const uint32_t N = A->N;
	// the reference allocates g,p,q,w,Aii in one block (:24-30); same here, grow-only.
	// A private copy of x lives in the same block so that one L2 window covers every
	// vector the iteration touches; it is copied in and out around the loop.
	const size_t Np = ((size_t)N + 1) & ~(size_t)1;   // keep every vector 16-byte aligned
	double *xw = c.ws, *g = xw + Np, *p = g + Np, *w = p + Np;
	double *q = jacobi ? w + Np : g, *diag = jacobi ? q + Np : nullptr;
	if (!c.dev_state) {
		NB_CUDA(nbgpu::dmalloc(&c.dev_state, sizeof(KrylovState)));
		NB_CUDA(cudaMallocHost(&c.host_state, 4 * sizeof(KrylovState)));
		NB_CUDA(cudaEventCreateWithFlags(&g_poll_ev[0], cudaEventDisableTiming));
		NB_CUDA(cudaEventCreateWithFlags(&g_poll_ev[1], cudaEventDisableTiming));
	}
	KrylovState *st = (KrylovState *)c.dev_state;
	KrylovState *hst = (KrylovState *)c.host_state;   // [0..1] poll slots, [2] init image, [3] final
	memset(&hst[2], 0, sizeof(KrylovState));
	hst[2].tol2 = tol * tol;
	hst[2].max_iter = max_iter;
	NB_CUDA(cudaMemcpyAsync(st, &hst[2], sizeof(KrylovState), cudaMemcpyHostToDevice, c.stream));
	NB_CUDA(cudaMemcpyAsync(xw, d_x, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));

	const int64_t slice_blocks = ((int64_t)A->n_slices * 32 + kBlock - 1) / kBlock;
	const int64_t vec_blocks = ((int64_t)N + 2 * kBlock - 1) / (2 * kBlock);
	const int igrid = jacobi ? resident_grid(krylov_init_kernel<true>, slice_blocks)
				 : resident_grid(krylov_init_kernel<false>, slice_blocks);
	const int sgrid = resident_grid(krylov_spmv_kernel, slice_blocks);
	const int ugrid = jacobi ? resident_grid(krylov_update_kernel<true>, vec_blocks)
				 : resident_grid(krylov_update_kernel<false>, vec_blocks);
	const int dgrid = resident_grid(krylov_dir_kernel, vec_blocks);
	// streamed (TMA) path: same kernels, matrix staged through shared memory
	SellView V;
	V.N = N; V.n_slices = A->n_slices; V.slice_off = A->d_slice_off; V.val = A->d_val;
	V.col = A->stream_ids();
	V.uniform_width = A->uniform_width;
	V.perm = A->d_perm;
	StreamConfig scfg, icfg;
	const int layout = A->layout();
	const void *sk = by_layout(layout, [](auto L) {
		return (const void *)krylov_spmv_stream_kernel<decltype(L)::value>;
	});
	const void *ik = by_layout(layout, [&](auto L) {
		return jacobi ? (const void *)krylov_init_stream_kernel<true, decltype(L)::value>
			      : (const void *)krylov_init_stream_kernel<false, decltype(L)::value>;
	});
	const bool stream = stream_config(A, sk, &scfg) && stream_config(A, ik, &icfg);
	const bool seq = g_seq_dots;
	const bool pdl = !seq && !getenv("NBGPU_NO_PDL");
	double *partials = seq ? nullptr : c.partials;

	cudaError_t e;
	if (stream) {
		e = by_layout(layout, [&](auto L) {
			constexpr int kL = decltype(L)::value;
			return jacobi ? launch(false, krylov_init_stream_kernel<true, kL>, icfg.grid, icfg.smem_bytes, V,
					       icfg, d_b, xw, g, p, q, diag, partials, st)
				      : launch(false, krylov_init_stream_kernel<false, kL>, icfg.grid, icfg.smem_bytes, V,
					       icfg, d_b, xw, g, p, q, diag, partials, st);
		});
	} else if (jacobi) {
		e = launch(false, krylov_init_kernel<true>, igrid, 0, N, A->n_slices, A->d_slice_off, A->d_perm, A->d_val, A->d_col,
			   d_b, xw, g, p, q, diag, partials, st);
	} else {
		e = launch(false, krylov_init_kernel<false>, igrid, 0, N, A->n_slices, A->d_slice_off, A->d_perm, A->d_val, A->d_col,
			   d_b, xw, g, p, q, diag, partials, st);
	}
	NB_CUDA(e);
	if (seq) {
		seq_dot_kernel<<<1, 32, 0, c.stream>>>(N, g, g, &st->gg[0], jacobi ? g : nullptr, q, &st->gq[0], st);
		NB_LAUNCHED();
	}

	g_prof_recorded = 0;
	if (g_prof_on && g_prof_ev.empty()) {
		g_prof_ev.resize(4 * kProfIters);
		for (auto &ev : g_prof_ev)
			NB_CUDA(cudaEventCreate(&ev));
	}
	uint32_t k = 0;
	int slot = 0;
	bool pending[2] = {false, false};
	bool finished = false;
	while (!finished) {
		const uint32_t k_end = (uint32_t)std::min<uint64_t>(max_iter, (uint64_t)k + kChunkIters);
		for (; k < k_end; k++) {
			const bool prof = g_prof_on && k < kProfIters;
			if (prof)
				NB_CUDA(cudaEventRecord(g_prof_ev[4 * k], c.stream));
			if (stream)
				e = by_layout(layout, [&](auto L) {
					return launch(pdl, krylov_spmv_stream_kernel<decltype(L)::value>, scfg.grid,
						      scfg.smem_bytes, k, V, scfg, p, w, partials, st);
				});
			else
				e = launch(pdl, krylov_spmv_kernel, sgrid, 0, k, N, A->n_slices, A->d_slice_off, A->d_perm, A->d_val,
					   A->d_col, p, w, partials, st);
			NB_CUDA(e);
			if (seq) {
				seq_dot_kernel<<<1, 32, 0, c.stream>>>(N, p, w, &st->pw, nullptr, nullptr, nullptr, st);
				NB_LAUNCHED();
			}
			if (prof)
				NB_CUDA(cudaEventRecord(g_prof_ev[4 * k + 1], c.stream));
			if (jacobi)
				e = launch(pdl, krylov_update_kernel<true>, ugrid, 0, k, N, w, diag, g, q, partials, st);
			else
				e = launch(pdl, krylov_update_kernel<false>, ugrid, 0, k, N, w, diag, g, q, partials, st);
			NB_CUDA(e);
			if (seq) {
				seq_dot_kernel<<<1, 32, 0, c.stream>>>(N, g, g, &st->gg[(k + 1) % 3u], jacobi ? g : nullptr,
								       q, &st->gq[(k + 1) & 1], st);
				NB_LAUNCHED();
			}
			if (prof)
				NB_CUDA(cudaEventRecord(g_prof_ev[4 * k + 2], c.stream));
			NB_CUDA(launch(pdl, krylov_dir_kernel, dgrid, 0, k, N, q, p, xw, st));
			if (prof) {
				NB_CUDA(cudaEventRecord(g_prof_ev[4 * k + 3], c.stream));
				g_prof_recorded = k + 1;
			}
		}
		if (k == max_iter) {
			// the loop test that ends the reference's while at k == max_iter
			NB_CUDA(launch(false, krylov_spmv_kernel, 1, 0, k, N, A->n_slices, A->d_slice_off, A->d_perm, A->d_val,
				       A->d_col, p, w, partials, st));
		}
		NB_CUDA(cudaMemcpyAsync(&hst[slot], st, sizeof(KrylovState), cudaMemcpyDeviceToHost, c.stream));
		NB_CUDA(cudaEventRecord(g_poll_ev[slot], c.stream));
		pending[slot] = true;
		const int other = slot ^ 1;
		if (pending[other]) {
			NB_CUDA(cudaEventSynchronize(g_poll_ev[other]));
			if (hst[other].done)
				finished = true;
		}
		if (k == max_iter)
			finished = true;
		slot ^= 1;
	}
	NB_CUDA(cudaMemcpyAsync(d_x, xw, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
	NB_CUDA(cudaMemcpyAsync(&hst[3], st, sizeof(KrylovState), cudaMemcpyDeviceToHost, c.stream));
	NB_CUDA(cudaStreamSynchronize(c.stream));
	if (!hst[3].done) {
		set_error("Krylov driver ended without the device gate firing (k=%u)", k);
		return NBGPU_ERR_CUDA;
	}
	if (g_prof_on) {
		// only iterations that actually ran (kernels past convergence return at once)
		const uint32_t n = std::min(g_prof_recorded, hst[3].k_final);
		g_prof_ms[0] = g_prof_ms[1] = g_prof_ms[2] = 0;
		for (uint32_t i = 0; i < n; i++)
			for (int j = 0; j < 3; j++) {
				float ms = 0;
				NB_CUDA(cudaEventElapsedTime(&ms, g_prof_ev[4 * i + j], g_prof_ev[4 * i + j + 1]));
				g_prof_ms[j] += ms;
			}
		g_prof_n = n;
	}
	if (niter)
		*niter = hst[3].k_final;
	if (tol_reached)
		*tol_reached = sqrt(hst[3].gg_final);
	// cg_precond_jacobi.c:86-89 (written so that NaN behaves like the reference's `>`)
	return (hst[3].gg_final > hst[3].tol2) ? NBGPU_NOT_CONVERGED : NBGPU_OK;
}

int solve(const nbgpu_matrix_t *A, const double *d_b, double *d_x, uint32_t max_iter, double tol,
	  uint32_t *niter, double *tol_reached, bool jacobi)
{
	NB_INIT();
	NB_ARG(A != nullptr && d_b != nullptr && d_x != nullptr);
	if (A->N == 0) {
		if (niter)
			*niter = 0;
		if (tol_reached)
			*tol_reached = 0.0;
		return NBGPU_OK;
	}
	const size_t Np = ((size_t)A->N + 1) & ~(size_t)1;
	const size_t ws_bytes = (jacobi ? 6 : 4) * Np * sizeof(double);
	NB_TRY(ensure_workspace(ws_bytes));
	const bool pinned = l2_pin(ctx().ws, ws_bytes, true);
	const int status = solve_impl(A, d_b, d_x, max_iter, tol, niter, tol_reached, jacobi);
	if (pinned)
		l2_unpin();
	return status;
}

int solve_host(const nbgpu_matrix_t *A, const double *b, double *x, uint32_t max_iter, double tol,
	       uint32_t *niter, double *tol_reached, bool jacobi)
{
	NB_INIT();
	NB_ARG(A != nullptr && b != nullptr && x != nullptr);
	const size_t Np = ((size_t)A->N + 1) & ~(size_t)1;
	double *d_b = nullptr, *d_x = nullptr;
	NB_TRY(nbgpu_malloc((void **)&d_b, 2 * Np * sizeof(double)));
	d_x = d_b + Np;
	int st = upload_vector(d_b, b, A->N);
	if (st == NBGPU_OK)
		st = upload_vector(d_x, x, A->N);
	int solver_status = NBGPU_OK;
	if (st == NBGPU_OK) {
		solver_status = solve(A, d_b, d_x, max_iter, tol, niter, tol_reached, jacobi);
		if (solver_status != NBGPU_OK && solver_status != NBGPU_NOT_CONVERGED)
			st = solver_status;
	}
	if (st == NBGPU_OK)
		st = download_vector(x, d_x, A->N);
	nbgpu_free(d_b);
	return st != NBGPU_OK ? st : solver_status;
}

}  // namespace

extern "C" {

int nbgpu_set_reduction_order(int mode)
{
	NB_ARG(mode == 0 || mode == 1);
	g_seq_dots = mode == 1;
	return NBGPU_OK;
}

int nbgpu_krylov_profile(int enable)
{
	g_prof_on = enable != 0;
	return NBGPU_OK;
}

int nbgpu_krylov_profile_get(double ms_total[3], uint32_t *n_iters)
{
	NB_ARG(ms_total != nullptr);
	for (int j = 0; j < 3; j++)
		ms_total[j] = g_prof_ms[j];
	if (n_iters)
		*n_iters = g_prof_n;
	return NBGPU_OK;
}

int nbgpu_pcg_jacobi(const nbgpu_matrix_t *A, const double *d_b, double *d_x, uint32_t max_iter,
		     double tolerance, uint32_t *niter_performed, double *tolerance_reached)
{
	return solve(A, d_b, d_x, max_iter, tolerance, niter_performed, tolerance_reached, true);
}

int nbgpu_cg(const nbgpu_matrix_t *A, const double *d_b, double *d_x, uint32_t max_iter,
	     double tolerance, uint32_t *niter_performed, double *tolerance_reached)
{
	return solve(A, d_b, d_x, max_iter, tolerance, niter_performed, tolerance_reached, false);
}

int nbgpu_pcg_jacobi_host(const nbgpu_matrix_t *A, const double *b, double *x, uint32_t max_iter,
			  double tolerance, uint32_t *niter_performed, double *tolerance_reached)
{
	return solve_host(A, b, x, max_iter, tolerance, niter_performed, tolerance_reached, true);
}

int nbgpu_cg_host(const nbgpu_matrix_t *A, const double *b, double *x, uint32_t max_iter,
		  double tolerance, uint32_t *niter_performed, double *tolerance_reached)
{
	return solve_host(A, b, x, max_iter, tolerance, niter_performed, tolerance_reached, false);
}

}  // extern "C"
