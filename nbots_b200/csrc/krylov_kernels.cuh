// krylov_kernels.cuh -- the kernels of Jacobi-PCG / CG, shared by the single-GPU solver (krylov.cu,
// Comm = NoComm) and the row-partitioned one (dist.cu, Comm = PeerComm, dist_comm.cuh).
//
// Reference: nb_sparse_solve_CG_precond_Jacobi
//   (sources/nb/solver_bot/sparse/solvers/cg_precond_jacobi.c:13-90) and
//   nb_sparse_solve_conjugate_gradient (solvers/conjugate_gradient.c:13-77).
//
// Two formulations of the same iteration:
//
// CLASSIC (3 kernels, 2 dependent reductions; the reference's recurrence operation by operation --
// with index-ordered dots, `nbgpu_set_reduction_order(1)`, a solve is bit-identical to the reference):
//   K1  spmv     w = A p,  pw = p.w
//   K2  update   g += a w, q = g/diag, gq' = g.q, gg' = g.g          a = gq/pw
//   K3  dir      x += a p, p = -q + b p                               b = gq'/gq
//
// FUSED (2 kernels, ONE reduction; opt-in, see nbgpu_set_pcg_mode): the matrix is applied to
// q instead of p and w = A p is carried by recurrence (Chronopoulos & Gear 1989):
//   K1  s = A q,  {g.q, q.s, g.g} reduced together;  b = gq/gq_old,  a = gq / (qs - b gq / a_old)
//   K2  p = -q + b p,  w = -s + b w,  x += a p,  g += a w,  q = g/diag     (no reduction)
// Same Krylov iterates in exact arithmetic; measured on the cantilever fixtures and at Q1
// (scripts/cg_variants_study.py): identical iteration counts, fields within 3e-11 -- but on the
// ill-conditioned void-material fixture 359 instead of 385 iterations, outside the +-2 % bar, hence
// not the default.  One kernel and one reduction tail fewer per iteration, ONE exchange per
// iteration between GPUs instead of two, +12 % vector traffic.
// g.g of K1(k) is |g_k|^2, the residual BEFORE this iteration's update -- exactly what the
// reference's pass 1 computes -- so the stale-residual stopping rule carries over unchanged.
//
// Stopping rule (cg_precond_jacobi.c:45,84-89): `while (gg > tol^2 && k < max_iter)` where gg is the
// value pass 1 computed, i.e. iteration k is gated on |g_{k-1}|^2 (|g_0|^2 for k = 0) and
// tolerance_reached reports that same stale number.  The gate is evaluated on the device by K1(k);
// the host only enqueues chunks of iterations and polls a `done` flag one chunk behind.
//
// Exchange (Comm): K1 pushes the boundary entries of the vector it gathers to the neighbours at its
// start and waits for theirs before its last (halo-reading) slices.  The CTA that finishes a grid
// reduction posts the rank's partial sums to every rank; the consuming kernel's CTAs collect the
// ranks' partials from their own memory and add them in rank order (dist_comm.cuh).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "dist_comm.cuh"
#include "sell_stream.cuh"

namespace nbgpu {

struct KrylovState {
	double gg[3];        // |g_k|^2 in slot k % 3
	double gq[2];        // g_k . q_k in slot k & 1 (plain CG: same as gg)
	double pw;           // p_k . A p_k
	double tol2;
	double gg_final;     // the value the reference's loop test failed on
	double alpha[2];     // FUSED: a_k in slot k & 1
	double beta;         // FUSED: b_k
	uint32_t max_iter;
	uint32_t k_final;    // iterations performed
	int32_t done;
	unsigned int ticket;
	unsigned int push_ticket;   // halo pushes (separate from the reductions' ticket: both live in K1)
};

// Optional timeline (build with EXTRA=-DNB_TIMELINE into another LIBDIR; scripts/dist_timeline.py): per-GPU
// %globaltimer stamps at the phase boundaries of the first 512 iterations, to see where an iteration of the
// row-partitioned solver waits.  Compiled out of the product.
#ifdef NB_TIMELINE
constexpr int kTlSlots = 10, kTlIters = 512;
static __device__ unsigned long long g_timeline[kTlIters * kTlSlots];
__device__ __forceinline__ void tl_stamp(uint32_t k, int slot)
{
	if (k < kTlIters) {
		unsigned long long t;
		asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
		g_timeline[k * kTlSlots + slot] = t;
	}
}
#define NB_TL(k, slot) tl_stamp(k, slot)
#else
#define NB_TL(k, slot) ((void)0)
#endif

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

// a failed exchange ends the solve: the host reports NBGPU_ERR_COMM from the window's error flag
__device__ __forceinline__ void comm_abort(KrylovState *st)
{
	if (threadIdx.x == 0) {
		__threadfence();
		st->done = 1;
	}
}

// g = A x - b, q = g / Aii, p = -q   (:33-43); FUSED starts from p = w = 0 (b_0 = 0)
template <bool JACOBI>
__device__ __forceinline__ void init_row(uint32_t row, double acc, double d, const double *__restrict__ b,
					 double *__restrict__ g, double *__restrict__ p, double *__restrict__ q,
					 double *__restrict__ diag, double *__restrict__ w, bool fused, double (&dots)[2])
{
	const double gi = __dsub_rn(acc, b[row]);
	g[row] = gi;
	dots[0] = __dadd_rn(dots[0], __dmul_rn(gi, gi));
	if (JACOBI) {
		const double qi = __ddiv_rn(gi, d);
		diag[row] = d;
		q[row] = qi;
		p[row] = fused ? 0.0 : -qi;
		dots[1] = __dadd_rn(dots[1], __dmul_rn(gi, qi));
	} else {
		p[row] = fused ? 0.0 : -gi;
	}
	if (fused)
		w[row] = 0.0;
}

template <bool JACOBI, typename Comm>
__device__ __forceinline__ void init_finish(double (&dots)[2], double *partials, KrylovState *st, const Comm &comm,
					    unsigned long long seq_in, unsigned long long seq_red)
{
	if (!partials)
		return;   // reference-order mode: seq_dot_kernel follows
	double tot[2];
	if (grid_reduce<2>(dots, partials, &st->ticket, tot)) {
		if (!JACOBI)
			tot[1] = tot[0];
		if (!comm.template all_reduce<2>(tot, seq_red)) {
			comm_abort(st);
			return;
		}
		if (threadIdx.x == 0) {
			st->gg[0] = tot[0];
			st->gq[0] = tot[1];
			st->gq[1] = 1.0;
			st->alpha[0] = st->alpha[1] = 1.0;
			st->beta = 0.0;
		}
		comm.ack_input(seq_in);
	}
}

// ---- register path (fallback for very wide slices; single GPU, CLASSIC only) ------------------
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
			init_row<JACOBI>(row, acc, d, b, g, p, q, diag, nullptr, false, dots);
	}
	init_finish<JACOBI>(dots, partials, st, NoComm(), 0, 0);
}

static __global__ void __launch_bounds__(kBlock, 4)
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
	if (!partials)
		return;
	double tot[1];
	if (grid_reduce<1>(dots, partials, &st->ticket, tot) && threadIdx.x == 0)
		st->pw = tot[0];
}

// ---- streamed path: the matrix goes through shared memory (sell_stream.cuh) -----------------
// x_ext / v_ext: base of the gathered vector in the matrix' column space; the row-indexed vectors
// (b, g, p, ...) are the owned parts.  On one GPU both coincide (col_shift = 0).
template <bool JACOBI, int LAYOUT, typename Comm>   // LAYOUT = blocked | idx16 << 1
__global__ void __launch_bounds__(kBlock, kStreamCtas)
krylov_init_stream_kernel(SellView A, StreamConfig cfg, Comm comm, unsigned long long seq_in,
			  unsigned long long seq_red, int fused, const double *__restrict__ b, const double *x_ext,
			  double *__restrict__ g, double *__restrict__ p, double *__restrict__ q,
			  double *__restrict__ diag, double *__restrict__ w, double *partials, KrylovState *st)
{
	extern __shared__ __align__(128) unsigned char smem[];
	double dots[2] = {0.0, 0.0};
	sell_stream_rows<(LAYOUT & 1) != 0, JACOBI, (LAYOUT & 2) != 0, Comm::kDist>(
		A, x_ext, cfg, smem, [] { return true; }, [&] { return comm.wait_halo(1, seq_in); }, NoPre(),
		[&](uint32_t row, double acc, double d, double, double) {
			if (row < A.N)
				init_row<JACOBI>(row, acc, d, b, g, p, q, diag, w, fused != 0, dots);
		});
	// a failed wait is raised in the window's error flag; the reduction below still has to be
	// taken by every CTA (ticket), its result is simply not used
	init_finish<JACOBI>(dots, partials, st, comm, seq_in, seq_red);
}

// CLASSIC K1: gate, (halo push,) w = A p, pw = p.w
// Row-partitioned, PUSH_WARP: the CTA carries ONE MORE WARP than the streaming ones.  It does nothing but the
// halo push (PeerComm::push_halo) and leaves.  Slices are dealt statically over the streaming warps, so the
// 5-6 us a pushing warp spends (index load, value load, peer stores, system fence until NVLink acknowledges,
// ticket) come out of the kernel's tail one for one when a streaming warp does it (measured with the phase
// timeline: K1 24.5 -> 30 us going from one rank to two; scripts/dist_timeline.py).  The price: 22 warps per
// SM leave 80 registers per thread instead of 96, which costs a bandwidth-bound K1 8 % (8 M dof per rank:
// 212 -> 229 us).  The host therefore uses the extra warp only where the fixed 5 us are the larger loss
// (krylov_run: rows per rank below kPushWarpMaxRows); without it the last streaming warp of the first CTAs
// pushes.
constexpr uint32_t kPushWarpMaxRows = 3000000;
template <bool PUSH_WARP>
constexpr int k1_block()
{
	return PUSH_WARP ? kBlock + 32 : kBlock;
}

template <int LAYOUT, typename Comm, bool PUSH_WARP = false>
__global__ void __launch_bounds__(k1_block<PUSH_WARP>(), kStreamCtas)
krylov_spmv_stream_kernel(uint32_t k, SellView A, StreamConfig cfg, Comm comm, unsigned long long seq_halo,
			  unsigned long long seq_red, const double *p_ext, double *__restrict__ w, double *partials,
			  KrylovState *st, int ticketless)
{
	extern __shared__ __align__(128) unsigned char smem[];
	double dots[1] = {0.0};
	bool active = true;
	if (PUSH_WARP && threadIdx.x >= kBlock) {
		pdl_wait();
		const int failed = comm.failed();
		active = iteration_gate(k, st) && !failed;
		pdl_launch_dependents();
		if (active)
			comm.push_halo(p_ext + A.col_shift, 0, seq_halo, &st->push_ticket);
		return;
	}
	sell_stream_rows<(LAYOUT & 1) != 0, false, (LAYOUT & 2) != 0, Comm::kDist>(
		A, p_ext, cfg, smem,
		[&] {
			// a function of (k, state) only: the whole grid takes the same branch
			const int failed = comm.failed();   // loaded beside the gate's state line, not behind it
			active = iteration_gate(k, st) && !failed;
			if (blockIdx.x == 0 && threadIdx.x == 0)
				NB_TL(k, 0);
			if (!PUSH_WARP && active)
				comm.push_halo(p_ext + A.col_shift, 0, seq_halo, &st->push_ticket);
			return active;
		},
		[&] {
			const bool ok = comm.wait_halo(0, seq_halo);   // only the slices that read halo columns wait
			if (blockIdx.x == 0 && threadIdx.x == 0)
				NB_TL(k, 2);
			return ok;
		},
		NoPre(),
		[&](uint32_t row, double acc, double, double p_row, double) {
			if (row < A.N) {
				w[row] = acc;
				dots[0] = __dadd_rn(dots[0], __dmul_rn(p_row, acc));
			}
		});
	if (!active || !partials)
		return;
	if (blockIdx.x == 0 && threadIdx.x == 0)
		NB_TL(k, 9);
	if (ticketless) {
		cta_store_partials<1>(dots, partials);   // K2's CTAs sum them (common.cuh)
		return;
	}
	double tot[1];
	if (grid_reduce<1>(dots, partials, &st->ticket, tot)) {
		if (threadIdx.x == 0)
			NB_TL(k, 1);
		if (Comm::kDist)
			comm.template post<1>(tot, seq_red);   // K2's CTAs collect the ranks' partials
		else if (threadIdx.x == 0)
			st->pw = tot[0];
	}
}

// CLASSIC K2: g += a w, q = g / diag, gq' = g.q, gg' = g.g   (:64-68; x += a p is done by K3)
// kVecBatch elements per thread and trip (a grid-stride apart, so every access is coalesced), every load
// of a trip issued before the first use.  The first trip's g, diag are loaded BEFORE the dependency wait
// (the SpMV kernel running ahead of us only writes w and the scalars) and its w right after the wait,
// ahead of the reduction of K1's partials: one round trip to L2 between the wait and the first store.
// At 1 M dof the resident grid covers the vector in a single trip.
constexpr int kVecBatch = 4;

template <bool JACOBI, typename Comm>
__global__ void __launch_bounds__(kBlock, 3)
krylov_update_kernel(uint32_t k, uint32_t N, Comm comm, unsigned long long seq_prev, unsigned long long seq_red,
		     const double *w, const double *__restrict__ diag, double *__restrict__ g,
		     double *__restrict__ q, double *partials, KrylovState *st, uint32_t n_prev, int ticketless)
{
	const uint32_t stride = gridDim.x * blockDim.x;
	const uint32_t base = blockIdx.x * blockDim.x + threadIdx.x;
	double gv[kVecBatch], dv[kVecBatch], wv[kVecBatch];
#pragma unroll
	for (int b = 0; b < kVecBatch; b++) {
		const uint64_t i = (uint64_t)base + (uint64_t)b * stride;
		gv[b] = i < N ? g[i] : 0.0;
		dv[b] = (JACOBI && i < N) ? diag[i] : 1.0;
	}
	pdl_wait();
	pdl_launch_dependents();
	const int32_t done = st->done;   // plain loads of the shared state line, see iteration_gate
	const double gq_k = st->gq[k & 1];
	double pw_k = st->pw;
	if (done)
		return;
#pragma unroll
	for (int b = 0; b < kVecBatch; b++) {
		const uint64_t i = (uint64_t)base + (uint64_t)b * stride;
		wv[b] = i < N ? w[i] : 0.0;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0)
		NB_TL(k, 3);
	if (n_prev) {
		// ticketless: K1's CTAs left their partials, every CTA sums them (same order everywhere); row-partitioned:
		// CTA 0 posts this rank's total to the peers and every CTA adds the peers' totals (PeerComm::exchange)
		double t[1];
		cta_sum_partials<1>(partials, n_prev, t);
		if (Comm::kDist && !comm.template exchange<1>(seq_prev, t)) {
			if (blockIdx.x == 0)
				comm_abort(st);
			return;
		}
		if (blockIdx.x == 0 && threadIdx.x == 0)
			NB_TL(k, 4);
		pw_k = t[0];
		if (blockIdx.x == 0 && threadIdx.x == 0)
			st->pw = pw_k;   // K3 reads it
	} else if (Comm::kDist) {
		// ticketed producer: every CTA collects the ranks' partials of K1 itself (dist_comm.cuh)
		double t[1];
		if (!comm.template collect<1>(seq_prev, t)) {
			if (blockIdx.x == 0)
				comm_abort(st);
			return;
		}
		if (blockIdx.x == 0 && threadIdx.x == 0)
			NB_TL(k, 4);
		pw_k = t[0];
		if (blockIdx.x == 0 && threadIdx.x == 0)
			st->pw = pw_k;
	}
	const double alpha = __ddiv_rn(gq_k, pw_k);
	double dots[2] = {0.0, 0.0};
	for (uint64_t i0 = base; i0 < N; i0 += (uint64_t)kVecBatch * stride) {
		if (i0 != base) {
#pragma unroll
			for (int b = 0; b < kVecBatch; b++) {
				const uint64_t i = i0 + (uint64_t)b * stride;
				gv[b] = i < N ? g[i] : 0.0;
				dv[b] = (JACOBI && i < N) ? diag[i] : 1.0;
				wv[b] = i < N ? w[i] : 0.0;
			}
		}
#pragma unroll
		for (int b = 0; b < kVecBatch; b++) {
			const uint64_t i = i0 + (uint64_t)b * stride;
			if (i >= N)
				break;
			const double gn = __dadd_rn(gv[b], __dmul_rn(alpha, wv[b]));
			g[i] = gn;
			dots[0] = __dadd_rn(dots[0], __dmul_rn(gn, gn));
			if (JACOBI) {
				const double qn = __ddiv_rn(gn, dv[b]);
				q[i] = qn;
				dots[1] = __dadd_rn(dots[1], __dmul_rn(gn, qn));
			}
		}
	}
	if (!partials)
		return;   // reference-order reductions are done by seq_dot_kernel
	if (ticketless) {
		if (!JACOBI)
			dots[1] = dots[0];
		cta_store_partials<2>(dots, partials + kPartialRegion);   // K3's CTAs sum them
		return;
	}
	double tot[2];
	if (grid_reduce<2>(dots, partials, &st->ticket, tot)) {
		if (threadIdx.x == 0)
			NB_TL(k, 5);
		if (!JACOBI)
			tot[1] = tot[0];
		if (Comm::kDist) {
			comm.template post<2>(tot, seq_red);   // K3's CTAs collect
		} else if (threadIdx.x == 0) {
			st->gg[(k + 1) % 3u] = tot[0];
			st->gq[(k + 1) & 1] = tot[1];
		}
	}
}

// CLASSIC K3: x += a p (:62-63, moved here from K2: p is read once per iteration instead of twice; same
// operations, same rounding), then p = -q + b p (:70-74); plain CG passes q == g.  p and x are
// preloaded before the dependency wait (the update kernel ahead of us writes neither), q right after it.
template <typename Comm>
__global__ void __launch_bounds__(kBlock)
krylov_dir_kernel(uint32_t k, uint32_t N, Comm comm, unsigned long long seq_prev, const double *q,
		  double *__restrict__ p, double *__restrict__ x, KrylovState *st, const double *partials,
		  uint32_t n_prev)
{
	const uint32_t stride = gridDim.x * blockDim.x;
	const uint32_t base = blockIdx.x * blockDim.x + threadIdx.x;
	double pv[kVecBatch], xv[kVecBatch], qv[kVecBatch];
#pragma unroll
	for (int b = 0; b < kVecBatch; b++) {
		const uint64_t i = (uint64_t)base + (uint64_t)b * stride;
		pv[b] = i < N ? p[i] : 0.0;
		xv[b] = i < N ? x[i] : 0.0;
	}
	pdl_wait();
	pdl_launch_dependents();
	const int32_t done = st->done;   // plain loads of the shared state line, see iteration_gate
	const double gq_k = st->gq[k & 1];
	double gq_n = st->gq[(k + 1) & 1];
	const double pw_k = st->pw;
	if (done)
		return;
#pragma unroll
	for (int b = 0; b < kVecBatch; b++) {
		const uint64_t i = (uint64_t)base + (uint64_t)b * stride;
		qv[b] = i < N ? q[i] : 0.0;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0)
		NB_TL(k, 6);
	if (n_prev) {
		double t[2];
		cta_sum_partials<2>(partials + kPartialRegion, n_prev, t);
		if (Comm::kDist && !comm.template exchange<2>(seq_prev, t)) {
			if (blockIdx.x == 0)
				comm_abort(st);
			return;
		}
		if (blockIdx.x == 0 && threadIdx.x == 0)
			NB_TL(k, 7);
		gq_n = t[1];
		if (blockIdx.x == 0 && threadIdx.x == 0) {
			st->gg[(k + 1) % 3u] = t[0];   // the gate of K1(k+1)
			st->gq[(k + 1) & 1] = t[1];
		}
	} else if (Comm::kDist) {
		double t[2];
		if (!comm.template collect<2>(seq_prev, t)) {
			if (blockIdx.x == 0)
				comm_abort(st);
			return;
		}
		if (blockIdx.x == 0 && threadIdx.x == 0)
			NB_TL(k, 7);
		gq_n = t[1];
		if (blockIdx.x == 0 && threadIdx.x == 0) {
			st->gg[(k + 1) % 3u] = t[0];
			st->gq[(k + 1) & 1] = t[1];
		}
	}
	const double alpha = __ddiv_rn(gq_k, pw_k);
	const double beta = __ddiv_rn(gq_n, gq_k);
	for (uint64_t i0 = base; i0 < N; i0 += (uint64_t)kVecBatch * stride) {
		if (i0 != base) {
#pragma unroll
			for (int b = 0; b < kVecBatch; b++) {
				const uint64_t i = i0 + (uint64_t)b * stride;
				pv[b] = i < N ? p[i] : 0.0;
				xv[b] = i < N ? x[i] : 0.0;
				qv[b] = i < N ? q[i] : 0.0;
			}
		}
#pragma unroll
		for (int b = 0; b < kVecBatch; b++) {
			const uint64_t i = i0 + (uint64_t)b * stride;
			if (i >= N)
				break;
			x[i] = __dadd_rn(xv[b], __dmul_rn(alpha, pv[b]));
			p[i] = __dadd_rn(-qv[b], __dmul_rn(beta, pv[b]));
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0)
		NB_TL(k, 8);
}

// FUSED: this iteration's step lengths from the reduced {g.q, q.Aq, g.g} and the previous iteration's
// g.q and a:  b = gq / gq_old,  p.Ap = q.Aq - b gq / a_old (three-term identity),  a = gq / p.Ap
__device__ __forceinline__ void fused_scalars(uint32_t k, const double (&tot)[3], const KrylovState *st, double *alpha,
					      double *beta)
{
	const double gq = tot[0], qs = tot[1];
	const double gq_old = st->gq[(k + 1) & 1], a_old = st->alpha[(k + 1) & 1];
	*beta = k ? __ddiv_rn(gq, gq_old) : 0.0;
	const double pw = k ? __dsub_rn(qs, __ddiv_rn(__dmul_rn(*beta, gq), a_old)) : qs;
	*alpha = __ddiv_rn(gq, pw);
}
__device__ __forceinline__ void fused_store(uint32_t k, const double (&tot)[3], double alpha, double beta, KrylovState *st)
{
	st->gg[k % 3u] = tot[2];
	st->gq[k & 1] = tot[0];
	st->pw = __ddiv_rn(tot[0], alpha);
	st->beta = beta;
	st->alpha[k & 1] = alpha;
}

// FUSED K1: gate, (halo push,) s = A v with v = q (Jacobi) or g (plain CG); g.v, v.s, g.g reduced
// together; the CTA that finishes the reduction (and the exchange) derives this iteration's a and b.
template <bool JACOBI, int LAYOUT, typename Comm, bool PUSH_WARP = false>
__global__ void __launch_bounds__(k1_block<PUSH_WARP>(), kStreamCtas)
krylov_fspmv_kernel(uint32_t k, SellView A, StreamConfig cfg, Comm comm, unsigned long long seq_halo,
		    unsigned long long seq_red, const double *v_ext, const double *g, double *__restrict__ s,
		    double *partials, KrylovState *st, int ticketless)
{
	extern __shared__ __align__(128) unsigned char smem[];
	double dots[3] = {0.0, 0.0, 0.0};   // g.v, v.s, g.g
	bool active = true;
	const uint32_t N = A.N;
	if (PUSH_WARP && threadIdx.x >= kBlock) {   // see krylov_spmv_stream_kernel
		pdl_wait();
		const int failed = comm.failed();
		active = iteration_gate(k, st) && !failed;
		pdl_launch_dependents();
		if (active)
			comm.push_halo(v_ext + A.col_shift, 0, seq_halo, &st->push_ticket);
		return;
	}
	sell_stream_rows<(LAYOUT & 1) != 0, false, (LAYOUT & 2) != 0, Comm::kDist>(
		A, v_ext, cfg, smem,
		[&] {
			const int failed = comm.failed();   // loaded beside the gate's state line, not behind it
			active = iteration_gate(k, st) && !failed;
			if (!PUSH_WARP && active)
				comm.push_halo(v_ext + A.col_shift, 0, seq_halo, &st->push_ticket);
			return active;
		},
		[&] { return comm.wait_halo(0, seq_halo); },
		[&](uint32_t row) { return (JACOBI && row < N) ? g[row] : 0.0; },
		[&](uint32_t row, double acc, double, double v_row, double g_row) {
			if (row < N) {
				s[row] = acc;
				const double gi = JACOBI ? g_row : v_row;
				dots[0] = __dadd_rn(dots[0], __dmul_rn(gi, v_row));
				dots[1] = __dadd_rn(dots[1], __dmul_rn(v_row, acc));
				if (JACOBI)
					dots[2] = __dadd_rn(dots[2], __dmul_rn(gi, gi));
			}
		});
	if (!active)
		return;
	if (ticketless) {
		if (!JACOBI)
			dots[2] = dots[0];
		cta_store_partials<3>(dots, partials);   // K2's CTAs sum them and derive a, b
		return;
	}
	double tot[3];
	if (grid_reduce<3>(dots, partials, &st->ticket, tot)) {
		if (!JACOBI)
			tot[2] = tot[0];
		if (Comm::kDist) {
			comm.template post<3>(tot, seq_red);   // K2's CTAs collect and derive a, b
		} else if (threadIdx.x == 0) {
			double alpha, beta;
			fused_scalars(k, tot, st, &alpha, &beta);
			fused_store(k, tot, alpha, beta, st);
		}
	}
}

// FUSED K2: p = -q + b p, w = -s + b w, x += a p, g += a w, q = g / diag -- no reduction.
// The old q is recomputed as g / diag (the same division that produced it: bit-identical) instead
// of being read.  Each CTA owns a contiguous chunk (grid = a multiple of the SM count, so the SMs
// carry equal shares); two consecutive elements per thread and trip as one 16-byte access.  The
// first trip's p, w, x, g, diag are loaded before the dependency wait (K1 writes only s).
template <bool JACOBI, typename Comm>
__global__ void __launch_bounds__(kBlock)
krylov_fupdate_kernel(uint32_t k, uint32_t N, Comm comm, unsigned long long seq_prev, const double *s,
		      const double *__restrict__ diag, double *g, double *q, double *__restrict__ p,
		      double *__restrict__ w, double *__restrict__ x, KrylovState *st, const double *partials,
		      uint32_t n_prev)
{
	const uint32_t n_pairs = (N + 1) >> 1;
	const uint32_t per_cta = (n_pairs + gridDim.x - 1) / gridDim.x;
	const uint32_t first = blockIdx.x * per_cta;
	const uint32_t last = min(n_pairs, first + per_cta);
	const uint32_t base = first + threadIdx.x;
	double2 p2 = {0, 0}, w2 = {0, 0}, x2 = {0, 0}, g2 = {0, 0}, d2 = {1, 1};
	if (base < last) {
		p2 = reinterpret_cast<const double2 *>(p)[base];
		w2 = reinterpret_cast<const double2 *>(w)[base];
		x2 = reinterpret_cast<const double2 *>(x)[base];
		g2 = reinterpret_cast<const double2 *>(g)[base];
		if (JACOBI)
			d2 = reinterpret_cast<const double2 *>(diag)[base];
	}
	pdl_wait();
	pdl_launch_dependents();
	const int32_t done = st->done;   // plain loads of the shared state line, see iteration_gate
	double alpha = st->alpha[k & 1], beta = st->beta;
	if (done)
		return;
	if (n_prev) {
		double t[3];
		cta_sum_partials<3>(partials, n_prev, t);
		if (Comm::kDist && !comm.template exchange<3>(seq_prev, t)) {
			if (blockIdx.x == 0)
				comm_abort(st);
			return;
		}
		fused_scalars(k, t, st, &alpha, &beta);
		if (blockIdx.x == 0 && threadIdx.x == 0)
			fused_store(k, t, alpha, beta, st);   // slots other than the ones this kernel's CTAs read
	} else if (Comm::kDist) {
		// ticketed producer: every CTA collects the ranks' partials of K1 and derives a, b
		double t[3];
		if (!comm.template collect<3>(seq_prev, t)) {
			if (blockIdx.x == 0)
				comm_abort(st);
			return;
		}
		fused_scalars(k, t, st, &alpha, &beta);
		if (blockIdx.x == 0 && threadIdx.x == 0)
			fused_store(k, t, alpha, beta, st);
	}
	for (uint32_t i = base; i < last; i += blockDim.x) {
		if (i != base) {
			p2 = reinterpret_cast<const double2 *>(p)[i];
			w2 = reinterpret_cast<const double2 *>(w)[i];
			x2 = reinterpret_cast<const double2 *>(x)[i];
			g2 = reinterpret_cast<const double2 *>(g)[i];
			if (JACOBI)
				d2 = reinterpret_cast<const double2 *>(diag)[i];
		}
		const double2 s2 = reinterpret_cast<const double2 *>(s)[i];
		const bool two = 2 * i + 1 < N;
		double2 qo, pn, wn, xn, gn, qn;
		qo.x = JACOBI ? __ddiv_rn(g2.x, d2.x) : g2.x;
		qo.y = JACOBI ? __ddiv_rn(g2.y, d2.y) : g2.y;
		pn.x = __dadd_rn(-qo.x, __dmul_rn(beta, p2.x));
		pn.y = __dadd_rn(-qo.y, __dmul_rn(beta, p2.y));
		wn.x = __dadd_rn(-s2.x, __dmul_rn(beta, w2.x));
		wn.y = __dadd_rn(-s2.y, __dmul_rn(beta, w2.y));
		xn.x = __dadd_rn(x2.x, __dmul_rn(alpha, pn.x));
		xn.y = __dadd_rn(x2.y, __dmul_rn(alpha, pn.y));
		gn.x = __dadd_rn(g2.x, __dmul_rn(alpha, wn.x));
		gn.y = __dadd_rn(g2.y, __dmul_rn(alpha, wn.y));
		if (two) {
			reinterpret_cast<double2 *>(p)[i] = pn;
			reinterpret_cast<double2 *>(w)[i] = wn;
			reinterpret_cast<double2 *>(x)[i] = xn;
			reinterpret_cast<double2 *>(g)[i] = gn;
		} else {
			p[2 * i] = pn.x;
			w[2 * i] = wn.x;
			x[2 * i] = xn.x;
			g[2 * i] = gn.x;
		}
		if (JACOBI) {
			qn.x = __ddiv_rn(gn.x, d2.x);
			qn.y = __ddiv_rn(gn.y, d2.y);
			if (two)
				reinterpret_cast<double2 *>(q)[i] = qn;
			else
				q[2 * i] = qn.x;
		}
	}
}

// Verification mode (nbgpu_set_reduction_order(1)): the dot products are summed
// by ONE thread in index order, exactly like the reference's single-threaded
// loops (the FEM driver passes omp_parallel_threads = 1,
// static_elasticity2D.c:90).  Every other operation of the CLASSIC solver already
// rounds like the reference, so in this mode the whole solve -- iterates,
// iteration count, tolerance_reached -- is bit-identical to the reference's.
// A warp loads 32 products at a time; lane 0 adds them in order.
static __global__ void seq_dot_kernel(uint32_t N, const double *__restrict__ a1, const double *__restrict__ b1,
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

// ---- host --------------------------------------------------------------------------

// Persistent grid of one kernel: exactly the number of CTAs that are resident at
// once (SMs x occupancy), so the grid-stride loops run as a single full wave.
template <typename Kernel>
int resident_grid(Kernel kernel, int64_t want_blocks, bool whole_sms = false)
{
	int per_sm = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, 0) != cudaSuccess || per_sm < 1) {
		cudaGetLastError();
		per_sm = 1;
	}
	const int sms = ctx().sm_count;
	int64_t cap = std::min<int64_t>((int64_t)sms * per_sm, kMaxPartialBlocks);
	int64_t grid = std::max<int64_t>(1, std::min(want_blocks, cap));
	if (whole_sms && grid >= sms)
		grid -= grid % sms;   // equal number of CTAs on every SM
	return (int)grid;
}

// kernel launch with (optionally) the programmatic-dependent-launch attribute
template <typename... KArgs, typename... Args>
cudaError_t launch_on(bool pdl, void (*kernel)(KArgs...), int grid, int block, size_t smem, Args &&...args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned)grid);
	cfg.blockDim = dim3((unsigned)block);
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

// Optional per-kernel timing (nbgpu_krylov_profile): CUDA events around each kernel of the first
// kProfIters iterations of a solve, on the stream the kernels run on.  Off by default; bench.py
// uses it in a separate solve to get the live duration of the dominant kernel.
constexpr uint32_t kProfIters = 256;
struct KrylovProfile {
	bool on = false;
	std::vector<cudaEvent_t> ev;   // 4 events per iteration
	uint32_t recorded = 0;
	double ms[3] = {0, 0, 0};      // [K1, K2, K3] (FUSED: K3 = 0)
	uint32_t n = 0;
};
KrylovProfile &krylov_profile();

// Everything one solve needs.  Row-indexed vectors point at the OWNED part; `v_ext` / `x_ext` are the
// bases of the gathered vectors in the matrix' column space (one GPU: the vectors themselves).
struct KrylovRun {
	const nbgpu_matrix_s *A = nullptr;
	SellView V;
	bool jacobi = true, fused = false, seq_dots = false, pdl = true;
	uint32_t max_iter = 0;
	double tol = 0;
	const double *b = nullptr;
	double *x = nullptr, *g = nullptr, *p = nullptr, *w = nullptr, *q = nullptr, *diag = nullptr, *s = nullptr;
	const double *v_ext = nullptr;   // CLASSIC: p; FUSED: q (Jacobi) / g (plain)
	const double *x_ext = nullptr;
	double *partials = nullptr;
	KrylovState *st = nullptr, *hst = nullptr;   // device block; pinned host: [0..1] poll slots, [2] init image, [3] final
	cudaEvent_t poll_ev[2] = {nullptr, nullptr};
	// exchange (PeerComm only)
	unsigned long long seq_in = 0;      // input-halo sequence of the init kernel
	unsigned long long msg_seq = 0;     // last reduction message posted before this solve
	unsigned long long halo_seq = 0;    // last Krylov-vector halo push before this solve
	const int *d_err = nullptr;         // the window's error flag
	int *h_err = nullptr;               // pinned [3]
};

// The host loop: init, chunks of iterations, polling one chunk behind.  Returns NBGPU_OK /
// NBGPU_NOT_CONVERGED / an error; *k_final = iterations performed (also when the exchange failed).
// compile-time switch for K1's extra push warp.  Only the row-partitioned kernels of the blocked layouts (the
// 2-dof FEM matrices) have the variant: at 80 registers the per-entry layouts would spill.
template <typename Comm, int LAYOUT, typename F>
auto by_push_warp(bool push_warp, F f)
{
	if constexpr (Comm::kDist && (LAYOUT & 1) != 0) {
		if (push_warp)
			return f(std::true_type{});
	}
	return f(std::false_type{});
}

template <typename Comm>
int krylov_run(KrylovRun &R, const Comm &comm, uint32_t *niter, double *tol_reached)
{
	Context &c = ctx();
	const nbgpu_matrix_s *A = R.A;
	const uint32_t N = A->N;
	KrylovState *st = R.st, *hst = R.hst;
	memset(&hst[2], 0, sizeof(KrylovState));
	hst[2].tol2 = R.tol * R.tol;
	hst[2].max_iter = R.max_iter;
	NB_CUDA(cudaMemcpyAsync(st, &hst[2], sizeof(KrylovState), cudaMemcpyHostToDevice, c.stream));
	if (R.h_err)
		R.h_err[0] = R.h_err[1] = R.h_err[2] = 0;

	const int layout = A->layout();
	const bool jacobi = R.jacobi, fused = R.fused;
	StreamConfig scfg, icfg;
	const void *ik = by_layout(layout, [&](auto L) {
		constexpr int kL = decltype(L)::value;
		return jacobi ? (const void *)krylov_init_stream_kernel<true, kL, Comm>
			      : (const void *)krylov_init_stream_kernel<false, kL, Comm>;
	});
	// the extra halo-push warp of K1 (see krylov_spmv_stream_kernel): NBGPU_DIST_PUSH_WARP=0|1 overrides
	const bool push_warp = Comm::kDist && (getenv("NBGPU_DIST_PUSH_WARP") ? atoi(getenv("NBGPU_DIST_PUSH_WARP")) != 0
										   : A->N <= kPushWarpMaxRows) &&
			       (layout & 1) != 0;
	const int k1_threads = push_warp ? k1_block<true>() : k1_block<false>();
	const void *sk = by_layout(layout, [&](auto L) {
		constexpr int kL = decltype(L)::value;
		return by_push_warp<Comm, kL>(push_warp, [&](auto PW) {
			constexpr bool kPW = decltype(PW)::value;
			if (!fused)
				return (const void *)krylov_spmv_stream_kernel<kL, Comm, kPW>;
			return jacobi ? (const void *)krylov_fspmv_kernel<true, kL, Comm, kPW>
				      : (const void *)krylov_fspmv_kernel<false, kL, Comm, kPW>;
		});
	});
	const bool stream = stream_config(A, sk, &scfg, k1_threads) && stream_config(A, ik, &icfg);
	if (!stream && (fused || Comm::kDist)) {
		set_error("this solver mode needs the streamed SpMV path (slice too wide or NBGPU_SPMV_PATH=reg)");
		return NBGPU_ERR_ARG;
	}
	// the init kernel keeps the plan's visit order (halo slices last); K1 hides them in the slack of the
	// warps that have no slice in the final partial round (sell_stream.cuh: place_halo_slices)
	SellView VK = R.V;
	if (Comm::kDist && stream && !getenv("NBGPU_DIST_HALO_LAST"))
		place_halo_slices(&VK, (uint32_t)scfg.grid * kStreamWarps);
	const bool seq = R.seq_dots;
	const bool pdl = R.pdl && !seq;
	double *partials = seq ? nullptr : R.partials;
	const int64_t slice_blocks = ((int64_t)A->n_slices * 32 + kBlock - 1) / kBlock;
	const int64_t vec_blocks = ((int64_t)N + kVecBatch * kBlock - 1) / (kVecBatch * kBlock);
	int igrid = 0, sgrid = 0, ugrid = 0, dgrid = 0, fgrid = 0;
	if (!stream) {
		igrid = jacobi ? resident_grid(krylov_init_kernel<true>, slice_blocks)
			       : resident_grid(krylov_init_kernel<false>, slice_blocks);
		sgrid = resident_grid(krylov_spmv_kernel, slice_blocks);
	}
	if (!fused) {
		ugrid = jacobi ? resident_grid(krylov_update_kernel<true, Comm>, vec_blocks)
			       : resident_grid(krylov_update_kernel<false, Comm>, vec_blocks);
		dgrid = resident_grid(krylov_dir_kernel<Comm>, vec_blocks);
	} else {
		const int64_t pair_blocks = ((int64_t)N / 2 + kBlock) / kBlock;   // one 16-byte pair per thread and trip
		fgrid = jacobi ? resident_grid(krylov_fupdate_kernel<true, Comm>, pair_blocks, true)
			       : resident_grid(krylov_fupdate_kernel<false, Comm>, pair_blocks, true);
	}

	// ticketless reductions (single GPU, parallel-tree dots, streamed K1): see common.cuh
	const bool tless = (!Comm::kDist || comm.consumer_posts()) && !seq && stream && !getenv("NBGPU_TICKETED");
	const int tl = tless ? 1 : 0;
	const uint32_t n_k1 = tless ? (uint32_t)scfg.grid : 0u, n_k2 = tless ? (uint32_t)ugrid : 0u;
	// message / halo sequence numbers: one reduction message for the init kernel, then 2 (CLASSIC) or
	// 1 (FUSED) per executed iteration; one halo push per executed iteration.  Kernels enqueued past
	// convergence post nothing, so every rank ends the solve with the same counters.
	const unsigned long long m0 = R.msg_seq + 1;
	auto msg_k1 = [&](uint32_t k) { return fused ? m0 + 1 + k : m0 + 1 + 2ull * k; };
	auto msg_k2 = [&](uint32_t k) { return m0 + 2 + 2ull * k; };
	auto halo_k = [&](uint32_t k) { return R.halo_seq + 1 + k; };

	cudaError_t e;
	if (stream) {
		e = by_layout(layout, [&](auto L) {
			constexpr int kL = decltype(L)::value;
			return jacobi ? launch_on(false, krylov_init_stream_kernel<true, kL, Comm>, icfg.grid, kBlock,
						  icfg.smem_bytes, R.V, icfg, comm, R.seq_in, m0, fused ? 1 : 0, R.b, R.x_ext,
						  R.g, R.p, R.q, R.diag, R.w, partials, st)
				      : launch_on(false, krylov_init_stream_kernel<false, kL, Comm>, icfg.grid, kBlock,
						  icfg.smem_bytes, R.V, icfg, comm, R.seq_in, m0, fused ? 1 : 0, R.b, R.x_ext,
						  R.g, R.p, R.q, R.diag, R.w, partials, st);
		});
	} else if (jacobi) {
		e = launch_on(false, krylov_init_kernel<true>, igrid, kBlock, 0, N, A->n_slices, A->d_slice_off, A->d_perm,
			      A->d_val, A->d_col, R.b, R.x_ext, R.g, R.p, R.q, R.diag, partials, st);
	} else {
		e = launch_on(false, krylov_init_kernel<false>, igrid, kBlock, 0, N, A->n_slices, A->d_slice_off, A->d_perm,
			      A->d_val, A->d_col, R.b, R.x_ext, R.g, R.p, R.q, R.diag, partials, st);
	}
	NB_CUDA(e);
	if (seq) {
		seq_dot_kernel<<<1, 32, 0, c.stream>>>(N, R.g, R.g, &st->gg[0], jacobi ? R.g : nullptr, R.q, &st->gq[0], st);
		NB_LAUNCHED();
	}

	KrylovProfile &prof = krylov_profile();
	prof.recorded = 0;
	if (prof.on && prof.ev.empty()) {
		prof.ev.resize(4 * kProfIters);
		for (auto &ev : prof.ev)
			NB_CUDA(cudaEventCreate(&ev));
	}
	uint32_t k = 0;
	int slot = 0;
	bool pending[2] = {false, false};
	bool finished = false;
	while (!finished) {
		const uint32_t k_end = (uint32_t)std::min<uint64_t>(R.max_iter, (uint64_t)k + kChunkIters);
		for (; k < k_end; k++) {
			const bool pr = prof.on && k < kProfIters;
			if (pr)
				NB_CUDA(cudaEventRecord(prof.ev[4 * k], c.stream));
			// ---- K1
			if (stream && fused)
				e = by_layout(layout, [&](auto L) {
					constexpr int kL = decltype(L)::value;
					return by_push_warp<Comm, kL>(push_warp, [&](auto PW) {
						constexpr bool kPW = decltype(PW)::value;
						return jacobi ? launch_on(pdl, krylov_fspmv_kernel<true, kL, Comm, kPW>, scfg.grid,
									  k1_threads, scfg.smem_bytes, k, VK, scfg, comm, halo_k(k),
									  msg_k1(k), R.v_ext, (const double *)R.g, R.s, partials, st, tl)
							      : launch_on(pdl, krylov_fspmv_kernel<false, kL, Comm, kPW>, scfg.grid,
									  k1_threads, scfg.smem_bytes, k, VK, scfg, comm, halo_k(k),
									  msg_k1(k), R.v_ext, (const double *)R.g, R.s, partials, st, tl);
					});
				});
			else if (stream)
				e = by_layout(layout, [&](auto L) {
					constexpr int kL = decltype(L)::value;
					return by_push_warp<Comm, kL>(push_warp, [&](auto PW) {
						return launch_on(pdl, krylov_spmv_stream_kernel<kL, Comm, decltype(PW)::value>, scfg.grid,
								 k1_threads, scfg.smem_bytes, k, VK, scfg, comm, halo_k(k), msg_k1(k),
								 R.v_ext, R.w, partials, st, tl);
					});
				});
			else
				e = launch_on(pdl, krylov_spmv_kernel, sgrid, kBlock, 0, k, N, A->n_slices, A->d_slice_off, A->d_perm,
					      A->d_val, A->d_col, (const double *)R.p, R.w, partials, st);
			NB_CUDA(e);
			if (seq) {
				seq_dot_kernel<<<1, 32, 0, c.stream>>>(N, R.p, R.w, &st->pw, nullptr, nullptr, nullptr, st);
				NB_LAUNCHED();
			}
			if (pr)
				NB_CUDA(cudaEventRecord(prof.ev[4 * k + 1], c.stream));
			// ---- K2 (, K3)
			if (fused) {
				e = jacobi ? launch_on(pdl, krylov_fupdate_kernel<true, Comm>, fgrid, kBlock, 0, k, N, comm, msg_k1(k),
						       (const double *)R.s, (const double *)R.diag, R.g, R.q, R.p, R.w, R.x, st,
						       (const double *)partials, n_k1)
					   : launch_on(pdl, krylov_fupdate_kernel<false, Comm>, fgrid, kBlock, 0, k, N, comm, msg_k1(k),
						       (const double *)R.s, (const double *)R.diag, R.g, R.q, R.p, R.w, R.x, st,
						       (const double *)partials, n_k1);
				NB_CUDA(e);
				if (pr) {
					NB_CUDA(cudaEventRecord(prof.ev[4 * k + 2], c.stream));
					NB_CUDA(cudaEventRecord(prof.ev[4 * k + 3], c.stream));
					prof.recorded = k + 1;
				}
				continue;
			}
			e = jacobi ? launch_on(pdl, krylov_update_kernel<true, Comm>, ugrid, kBlock, 0, k, N, comm, msg_k1(k),
					       msg_k2(k), (const double *)R.w, (const double *)R.diag, R.g, R.q, partials, st, n_k1, tl)
				   : launch_on(pdl, krylov_update_kernel<false, Comm>, ugrid, kBlock, 0, k, N, comm, msg_k1(k),
					       msg_k2(k), (const double *)R.w, (const double *)R.diag, R.g, R.q, partials, st, n_k1, tl);
			NB_CUDA(e);
			if (seq) {
				seq_dot_kernel<<<1, 32, 0, c.stream>>>(N, R.g, R.g, &st->gg[(k + 1) % 3u], jacobi ? R.g : nullptr,
								       R.q, &st->gq[(k + 1) & 1], st);
				NB_LAUNCHED();
			}
			if (pr)
				NB_CUDA(cudaEventRecord(prof.ev[4 * k + 2], c.stream));
			NB_CUDA(launch_on(pdl, krylov_dir_kernel<Comm>, dgrid, kBlock, 0, k, N, comm, msg_k2(k),
					  (const double *)R.q, R.p, R.x, st, (const double *)partials, n_k2));
			if (pr) {
				NB_CUDA(cudaEventRecord(prof.ev[4 * k + 3], c.stream));
				prof.recorded = k + 1;
			}
		}
		if (k == R.max_iter) {
			// the loop test that ends the reference's while at k == max_iter (gate only: k >= max_iter)
			NB_CUDA(launch_on(false, krylov_spmv_kernel, 1, kBlock, 0, k, N, A->n_slices, A->d_slice_off, A->d_perm,
					  A->d_val, A->d_col, (const double *)R.p, R.w, partials, st));
		}
		NB_CUDA(cudaMemcpyAsync(&hst[slot], st, sizeof(KrylovState), cudaMemcpyDeviceToHost, c.stream));
		if (R.d_err)
			NB_CUDA(cudaMemcpyAsync(&R.h_err[slot], R.d_err, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
		NB_CUDA(cudaEventRecord(R.poll_ev[slot], c.stream));
		pending[slot] = true;
		const int other = slot ^ 1;
		if (pending[other]) {
			NB_CUDA(cudaEventSynchronize(R.poll_ev[other]));
			if (hst[other].done || (R.h_err && R.h_err[other]))   // converged, or a wait timed out
				finished = true;
		}
		if (k == R.max_iter)
			finished = true;
		slot ^= 1;
	}
	NB_CUDA(cudaMemcpyAsync(&hst[3], st, sizeof(KrylovState), cudaMemcpyDeviceToHost, c.stream));
	if (R.d_err)
		NB_CUDA(cudaMemcpyAsync(&R.h_err[2], R.d_err, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
	NB_CUDA(cudaStreamSynchronize(c.stream));
	if (R.h_err && R.h_err[2]) {
		set_error("distributed solve: a peer did not answer within the timeout");
		return NBGPU_ERR_COMM;
	}
	if (!hst[3].done) {
		set_error("Krylov driver ended without the device gate firing (k=%u)", k);
		return NBGPU_ERR_CUDA;
	}
	if (prof.on) {
		// only iterations that actually ran (kernels past convergence return at once)
		const uint32_t n = std::min(prof.recorded, hst[3].k_final);
		prof.ms[0] = prof.ms[1] = prof.ms[2] = 0;
		for (uint32_t i = 0; i < n; i++)
			for (int j = 0; j < 3; j++) {
				float ms = 0;
				NB_CUDA(cudaEventElapsedTime(&ms, prof.ev[4 * i + j], prof.ev[4 * i + j + 1]));
				prof.ms[j] += ms;
			}
		prof.n = n;
	}
	const uint32_t kf = hst[3].k_final;
	R.msg_seq += 1 + (fused ? 1ull : 2ull) * kf;
	R.halo_seq += kf;
	if (niter)
		*niter = kf;
	if (tol_reached)
		*tol_reached = sqrt(hst[3].gg_final);
	// cg_precond_jacobi.c:86-89 (written so that NaN behaves like the reference's `>`)
	return (hst[3].gg_final > hst[3].tol2) ? NBGPU_NOT_CONVERGED : NBGPU_OK;
}

// which formulation a solve uses: FUSED unless the caller asked for the reference-order dots, the
// classic recurrence (nbgpu_set_pcg_mode / NBGPU_PCG_MODE=classic), or the matrix has no streamed path
bool krylov_want_fused(const nbgpu_matrix_s *A, bool seq_dots);
bool krylov_seq_dots();

}  // namespace nbgpu
