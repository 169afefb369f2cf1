// krylov.cu -- Jacobi-preconditioned CG and plain CG on the device.
//
// Reference: nb_sparse_solve_CG_precond_Jacobi
//   (sources/nb/solver_bot/sparse/solvers/cg_precond_jacobi.c:13-90) and
//   nb_sparse_solve_conjugate_gradient (solvers/conjugate_gradient.c:13-77).
//
// One reference iteration is three OpenMP loops separated by two scalar
// reductions; here it is three kernels and the scalars never visit the host:
//
//   K1  iter_spmv    w = A p,  pw = p.w          (reduction fused in the SpMV)
//   K2  iter_update  x += a p, g += a w, q = g/diag, gq' = g.q, gg' = g.g
//   K3  iter_dir     p = -q + b p
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
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "spmv.cuh"

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

// g = A x - b, diag, q = g / diag, p = -q, gg = g.g, gq = g.q   (init, :33-43)
template <bool JACOBI>
__global__ void __launch_bounds__(kBlock, 4)
krylov_init_kernel(uint32_t N, uint32_t n_slices, const uint32_t *__restrict__ slice_off,
		   const double *__restrict__ val, const uint32_t *__restrict__ col,
		   const double *__restrict__ b, const double *__restrict__ x, double *__restrict__ g,
		   double *__restrict__ p, double *__restrict__ q, double *__restrict__ diag,
		   double *partials, KrylovState *st)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	double dots[2] = {0.0, 0.0};
	for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slices; s += warps) {
		const uint32_t row = s * kSliceRows + lane;
		const uint32_t off = __ldg(slice_off + s);
		const uint32_t width = __ldg(slice_off + s + 1) - off;
		double d = 0.0;
		const double acc = sell_row_times<kIterUnroll, JACOBI>(val, col, off, width, lane, row,
								       min(row, N - 1), x, &d);
		if (row < N) {
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
	}
	double tot[2];
	if (grid_reduce<2>(dots, partials, &st->ticket, tot) && threadIdx.x == 0) {
		st->gg[0] = tot[0];
		st->gq[0] = JACOBI ? tot[1] : tot[0];
	}
}

// K1: gate, w = A p, pw = p.w   (:45, :50-58)
__global__ void __launch_bounds__(kBlock, 4)
krylov_spmv_kernel(uint32_t k, uint32_t N, uint32_t n_slices, const uint32_t *__restrict__ slice_off,
		   const double *__restrict__ val, const uint32_t *__restrict__ col,
		   const double *__restrict__ p, double *__restrict__ w, double *partials,
		   KrylovState *st)
{
	if (*(volatile int32_t *)&st->done)
		return;
	{
		const double gg = st->gg[gate_slot(k)];
		const bool active = gg > st->tol2 && k < st->max_iter;
		if (!active) {
			if (blockIdx.x == 0 && threadIdx.x == 0) {
				st->k_final = k;
				st->gg_final = gg;
				__threadfence();
				st->done = 1;
			}
			return;
		}
	}
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
	double dots[1] = {0.0};
	for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < n_slices; s += warps) {
		const uint32_t row = s * kSliceRows + lane;
		const uint32_t off = __ldg(slice_off + s);
		const uint32_t width = __ldg(slice_off + s + 1) - off;
		const double acc = sell_row_times<kIterUnroll, false>(val, col, off, width, lane, row,
								      min(row, N - 1), p, nullptr);
		if (row < N) {
			w[row] = acc;
			dots[0] = __dadd_rn(dots[0], __dmul_rn(__ldg(p + row), acc));
		}
	}
	double tot[1];
	if (grid_reduce<1>(dots, partials, &st->ticket, tot) && threadIdx.x == 0)
		st->pw = tot[0];
}

// K2: x += a p, g += a w, q = g / diag, gq' = g.q, gg' = g.g   (:59-68)
template <bool JACOBI>
__global__ void __launch_bounds__(kBlock)
krylov_update_kernel(uint32_t k, uint32_t N, const double *__restrict__ p, const double *__restrict__ w,
		     const double *__restrict__ diag, double *__restrict__ x, double *__restrict__ g,
		     double *__restrict__ q, double *partials, KrylovState *st)
{
	if (*(volatile int32_t *)&st->done)
		return;
	const double alpha = __ddiv_rn(st->gq[k & 1], st->pw);
	double dots[2] = {0.0, 0.0};
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride) {
		const double pi = p[i], wi = w[i];
		x[i] = __dadd_rn(x[i], __dmul_rn(alpha, pi));
		const double gi = __dadd_rn(g[i], __dmul_rn(alpha, wi));
		g[i] = gi;
		dots[0] = __dadd_rn(dots[0], __dmul_rn(gi, gi));
		if (JACOBI) {
			const double qi = __ddiv_rn(gi, diag[i]);
			q[i] = qi;
			dots[1] = __dadd_rn(dots[1], __dmul_rn(gi, qi));
		}
	}
	double tot[2];
	if (grid_reduce<2>(dots, partials, &st->ticket, tot) && threadIdx.x == 0) {
		st->gg[(k + 1) % 3u] = tot[0];
		st->gq[(k + 1) & 1] = JACOBI ? tot[1] : tot[0];
	}
}

// K3: p = -q + b p   (:70-74); plain CG passes q == g
__global__ void __launch_bounds__(kBlock)
krylov_dir_kernel(uint32_t k, uint32_t N, const double *__restrict__ q, double *__restrict__ p,
		  const KrylovState *st)
{
	if (*(volatile const int32_t *)&st->done)
		return;
	const double beta = __ddiv_rn(st->gq[(k + 1) & 1], st->gq[k & 1]);
	const uint32_t stride = gridDim.x * blockDim.x;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += stride)
		p[i] = __dadd_rn(-q[i], __dmul_rn(beta, p[i]));
}

int vector_grid(uint32_t N)
{
	int64_t want = ((int64_t)N + kBlock - 1) / kBlock;
	int64_t cap = std::min<int64_t>((int64_t)ctx().sm_count * 8, kMaxPartialBlocks);
	return (int)std::max<int64_t>(1, std::min(want, cap));
}

cudaEvent_t g_poll_ev[2] = {nullptr, nullptr};

// Optional per-kernel timing (nbgpu_krylov_profile): CUDA events around each of
// the three kernels for the first kProfIters iterations of a solve, on the
// stream the kernels run on.  Off by default; bench.py uses it in a separate,
// untimed-for-throughput solve to get the live duration of the dominant kernel.
constexpr uint32_t kProfIters = 256;
bool g_prof_on = false;
std::vector<cudaEvent_t> g_prof_ev;    // 4 events per iteration
uint32_t g_prof_recorded = 0;
double g_prof_ms[3] = {0, 0, 0};
uint32_t g_prof_n = 0;

int solve(const nbgpu_matrix_t *A, const double *d_b, double *d_x, uint32_t max_iter, double tol,
	  uint32_t *niter, double *tol_reached, bool jacobi)
{
	NB_INIT();
	NB_ARG(A != nullptr && d_b != nullptr && d_x != nullptr);
	Context &c = ctx();
	const uint32_t N = A->N;
	if (N == 0) {
		if (niter)
			*niter = 0;
		if (tol_reached)
			*tol_reached = 0.0;
		return NBGPU_OK;
	}
	// the reference allocates g,p,q,w,Aii in one block (:24-30); same here, grow-only
	const size_t n_vec = jacobi ? 5 : 3;
	NB_TRY(ensure_workspace(n_vec * (size_t)N * sizeof(double)));
	double *g = c.ws, *p = g + N, *w = p + N;
	double *q = jacobi ? w + N : g, *diag = jacobi ? q + N : nullptr;
	if (!c.dev_state) {
		NB_CUDA(cudaMalloc(&c.dev_state, sizeof(KrylovState)));
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

	const int sgrid = spmv_grid(A->n_slices), vgrid = vector_grid(N);
	if (jacobi)
		krylov_init_kernel<true><<<sgrid, kBlock, 0, c.stream>>>(
			N, A->n_slices, A->d_slice_off, A->d_val, A->d_col, d_b, d_x, g, p, q, diag,
			c.partials, st);
	else
		krylov_init_kernel<false><<<sgrid, kBlock, 0, c.stream>>>(
			N, A->n_slices, A->d_slice_off, A->d_val, A->d_col, d_b, d_x, g, p, q, diag,
			c.partials, st);
	NB_LAUNCHED();

	g_prof_recorded = 0;
	if (g_prof_on && g_prof_ev.empty()) {
		g_prof_ev.resize(4 * kProfIters);
		for (auto &e : g_prof_ev)
			NB_CUDA(cudaEventCreate(&e));
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
			krylov_spmv_kernel<<<sgrid, kBlock, 0, c.stream>>>(
				k, N, A->n_slices, A->d_slice_off, A->d_val, A->d_col, p, w, c.partials, st);
			NB_LAUNCHED();
			if (prof)
				NB_CUDA(cudaEventRecord(g_prof_ev[4 * k + 1], c.stream));
			if (jacobi)
				krylov_update_kernel<true><<<vgrid, kBlock, 0, c.stream>>>(
					k, N, p, w, diag, d_x, g, q, c.partials, st);
			else
				krylov_update_kernel<false><<<vgrid, kBlock, 0, c.stream>>>(
					k, N, p, w, diag, d_x, g, q, c.partials, st);
			NB_LAUNCHED();
			if (prof)
				NB_CUDA(cudaEventRecord(g_prof_ev[4 * k + 2], c.stream));
			krylov_dir_kernel<<<vgrid, kBlock, 0, c.stream>>>(k, N, q, p, st);
			NB_LAUNCHED();
			if (prof) {
				NB_CUDA(cudaEventRecord(g_prof_ev[4 * k + 3], c.stream));
				g_prof_recorded = k + 1;
			}
		}
		if (k == max_iter) {
			// the loop test that ends the reference's while at k == max_iter
			krylov_spmv_kernel<<<1, kBlock, 0, c.stream>>>(
				k, N, A->n_slices, A->d_slice_off, A->d_val, A->d_col, p, w, c.partials, st);
			NB_LAUNCHED();
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

int solve_host(const nbgpu_matrix_t *A, const double *b, double *x, uint32_t max_iter, double tol,
	       uint32_t *niter, double *tol_reached, bool jacobi)
{
	NB_INIT();
	NB_ARG(A != nullptr && b != nullptr && x != nullptr);
	Context &c = ctx();
	const size_t bytes = (size_t)A->N * sizeof(double);
	double *d_b = nullptr, *d_x = nullptr;
	NB_TRY(nbgpu_malloc((void **)&d_b, 2 * bytes));
	d_x = d_b + A->N;
	int st = NBGPU_OK;
	if (cudaMemcpyAsync(d_b, b, bytes, cudaMemcpyHostToDevice, c.stream) != cudaSuccess ||
	    cudaMemcpyAsync(d_x, x, bytes, cudaMemcpyHostToDevice, c.stream) != cudaSuccess) {
		set_error("upload of b/x failed: %s", cudaGetErrorString(cudaGetLastError()));
		st = NBGPU_ERR_CUDA;
	}
	int solver_status = NBGPU_OK;
	if (st == NBGPU_OK) {
		solver_status = solve(A, d_b, d_x, max_iter, tol, niter, tol_reached, jacobi);
		if (solver_status != NBGPU_OK && solver_status != NBGPU_NOT_CONVERGED)
			st = solver_status;
	}
	if (st == NBGPU_OK && cudaMemcpy(x, d_x, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) {
		set_error("download of x failed: %s", cudaGetErrorString(cudaGetLastError()));
		st = NBGPU_ERR_CUDA;
	}
	nbgpu_free(d_b);
	return st != NBGPU_OK ? st : solver_status;
}

}  // namespace

extern "C" {

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
