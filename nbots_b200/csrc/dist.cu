// dist.cu -- row-partitioned Jacobi-PCG / CG / SpMV over several B200s of one node.
//
// One process per GPU.  Rank r owns the contiguous rows
// [row_starts[r], row_starts[r+1]) of the global matrix and the matching slices
// of every vector.  Its block of the matrix is a rank-local SELL matrix whose
// column space is "owned columns, then halo columns" (halo = the remote rows its
// rows reference, ascending global id, hence grouped by owner); entries keep the
// ascending GLOBAL column order inside a row, so every row sum is still the
// reference's (sparse.c:405-414).
//
// Exchange: no collective library call sits in the iteration.  Every rank
// exports one device allocation (its "window": a control block plus the two
// vectors that have a halo tail) through CUDA IPC; peers map it and
//   * push the boundary entries of p straight into the neighbours' halo tails
//     (plain stores over NVLink, one CTA per neighbour), and
//   * post their partial dot products into one slot per rank of every peer's
//     control block,
// each followed by a system-scope release store of a sequence number.
// Consumers poll THEIR OWN memory (acquire loads) at the start of the kernel
// that needs the data.  All ranks add the per-rank partials in rank order, so
// alpha, beta and the stopping decision are bit-identical on every rank and the
// iteration count does not depend on the number of GPUs' arrival order.
//
// Why the single-slot / single-buffer scheme is race-free: a rank can only run
// ahead of a peer by less than one reduction.  K2(k) cannot start before every
// rank finished K1(k) (it needs all p.w partials), K3(k) not before every K2(k)
// (all g.q partials), and K1(k+1) reads the halo the neighbours' K3(k) pushed.
// Hence nobody overwrites a slot or a halo tail that a peer has yet to read.
//
// Every wait has a wall-clock timeout (NBGPU_DIST_TIMEOUT_MS, default 10000): on
// expiry the kernel raises the window's error flag, all later kernels of the
// solve return immediately and the host reports NBGPU_ERR_COMM -- a lost peer
// cannot hang the GPU.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "sell_stream.cuh"

using namespace nbgpu;

namespace {

constexpr int kMaxRanks = 16;

struct DistControl {
	unsigned long long halo_seq[kMaxRanks];   // p halo pushed by rank src
	unsigned long long xh_seq[kMaxRanks];     // generic vector halo (x at init, SpMV input)
	unsigned long long xh_ack[kMaxRanks];     // ... consumed by rank dst (flow control for SpMV)
	// Reduction partials travel as self-validating 16-byte messages {value bits, sequence}: one
	// NVLink store per message, no fence on the sender, the receiver polls the pair.
	ulonglong2 pw_msg[kMaxRanks];             // p.w partial of rank src
	ulonglong2 gq_msg[kMaxRanks][2];          // (g.g, g.q) partials of rank src
	int error;
};

struct DistState {
	double gg[3];
	double gq[2];
	double pw;
	double tol2;
	double gg_final;
	uint32_t max_iter;
	uint32_t k_final;
	int32_t done;
	unsigned int ticket;
	unsigned int push_ticket;   // halo pushes (separate from the reductions' ticket: both live in K1)
};

struct PeerTable {
	DistControl *ctrl[kMaxRanks];      // peers' control blocks (own included), in MY address space
	double *p_halo_dst[kMaxRanks];     // where my boundary values of p go in peer d's halo tail
	double *x_halo_dst[kMaxRanks];     // same for the second ext vector
	uint32_t send_ptr[kMaxRanks + 1];  // my send list, grouped by destination
	int world, rank;
	uint32_t n_recv_src;               // ranks I receive a halo from
	int recv_src[kMaxRanks];
	unsigned long long timeout_ns;
	int push_first;                    // A/B: halo push by warp 0 of every CTA (the first scheme)
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
	asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_msg(ulonglong2 *p, double v, unsigned long long seq)
{
	asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"((unsigned long long)__double_as_longlong(v)),
		     "l"(seq)
		     : "memory");
}
__device__ __forceinline__ ulonglong2 ld_msg(const ulonglong2 *p)
{
	ulonglong2 m;
	asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(m.x), "=l"(m.y) : "l"(p) : "memory");
	return m;
}
__device__ __forceinline__ unsigned long long global_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}

// one thread waits until *flag >= seq (or the timeout / another kernel's error)
__device__ __forceinline__ bool wait_seq(const unsigned long long *flag, unsigned long long seq,
					 DistControl *mine, unsigned long long timeout_ns)
{
	if (ld_acquire_sys(flag) >= seq)
		return true;
	const unsigned long long t0 = global_ns();
	for (;;) {
		if (ld_acquire_sys(flag) >= seq)
			return true;
		if (*(volatile int *)&mine->error)
			return false;
		if (global_ns() - t0 > timeout_ns) {
			*(volatile int *)&mine->error = 1;
			return false;
		}
		__nanosleep(64);
	}
}

// one thread waits for the message with sequence `seq`; the value comes back in *v
__device__ __forceinline__ bool wait_msg(const ulonglong2 *slot, unsigned long long seq, double *v, DistControl *mine,
					 unsigned long long timeout_ns)
{
	unsigned long long t0 = 0;
	for (unsigned int spin = 0;; spin++) {
		const ulonglong2 m = ld_msg(slot);
		if (m.y == seq) {
			*v = __longlong_as_double((long long)m.x);
			return true;
		}
		if (spin < 64)
			continue;   // the common case: the message is at most a few microseconds away
		if (t0 == 0)
			t0 = global_ns();
		if (*(volatile int *)&mine->error)
			return false;
		if (global_ns() - t0 > timeout_ns) {
			*(volatile int *)&mine->error = 1;
			return false;
		}
		__nanosleep(32);
	}
}

// CTA-wide, CTA-uniform gather of NV-value messages from every rank, summed in rank order.
// false if the solve is over (done / error) or a wait failed.
template <int NV, typename SlotOf>
__device__ __forceinline__ bool cta_reduce_msgs(int world, SlotOf slot_of, unsigned long long seq, DistControl *mine,
						unsigned long long timeout_ns, const int32_t *done, double (&tot)[NV])
{
	__shared__ double s_val[NV][kMaxRanks];
	int ok = 1;
	if ((int)threadIdx.x < world * NV) {
		// the pollers look at done / error themselves: one barrier for the whole step
		ok = !(done && *(volatile const int32_t *)done) && !*(volatile int *)&mine->error;
		const int r = threadIdx.x / NV, c = threadIdx.x % NV;
		double v = 0.0;
		if (ok)
			ok = wait_msg(slot_of(r, c), seq, &v, mine, timeout_ns) ? 1 : 0;
		s_val[c][r] = v;
	}
	if (!__syncthreads_and(ok))
		return false;
#pragma unroll
	for (int c = 0; c < NV; c++) {
		double t = 0.0;
		for (int r = 0; r < world; r++)
			t += s_val[c][r];
		tot[c] = t;
	}
	return true;
}

// ---- halo push ---------------------------------------------------------------------
// The send list (boundary entries of v, grouped by destination) is split evenly over
// the CTAs of the calling kernel; in each CTA ONE warp copies its piece into the
// neighbours' halo tails with plain NVLink stores, fences at system scope and takes a
// ticket; the last arriver raises the destinations' flags.  Other warps of the CTA
// are not held up.  Cost is independent of the halo size up to ~grid x 32 entries
// per trip (the round-1 first version pushed from a single CTA: 60 us for a 64 KB
// halo).  which: 0 -> p (halo_seq), 1 -> input vector (xh_seq; waits for the
// destination's ack of the previous push first, the input halo is single-buffered).
__device__ __forceinline__ void push_halo_piece(const PeerTable &T, const uint32_t *__restrict__ send_idx,
						const double *v, int which, unsigned long long seq,
						unsigned int *ticket, uint32_t piece, uint32_t n_pieces)
{
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t total = T.send_ptr[T.world];
	DistControl *mine = T.ctrl[T.rank];
	const uint32_t per = (total + n_pieces - 1) / n_pieces;
	const uint32_t b = min(total, piece * per), e = min(total, b + per);
	bool ok = true;
	if (which == 1 && e > b) {
		// destinations touched by [b, e)
		for (int d = 0; d < T.world && ok; d++)
			if (T.send_ptr[d] < e && T.send_ptr[d + 1] > b) {
				int w = 1;
				if (lane == 0)
					w = wait_seq(&mine->xh_ack[d], seq - 1, mine, T.timeout_ns) ? 1 : 0;
				ok = __shfl_sync(0xffffffffu, w, 0) != 0;
			}
	}
	if (ok) {
		for (uint32_t j = b + lane; j < e; j += 32) {
			int d = 0;
			while (j >= T.send_ptr[d + 1])
				d++;
			double *dst = which ? T.x_halo_dst[d] : T.p_halo_dst[d];
			dst[j - T.send_ptr[d]] = v[send_idx[j]];
		}
		if (e > b)
			__threadfence_system();
	}
	__syncwarp();
	if (lane == 0) {
		const unsigned int t = atomicInc(ticket, n_pieces - 1);
		if (t == n_pieces - 1) {
			__threadfence_system();
			for (int d = 0; d < T.world; d++)
				if (T.send_ptr[d + 1] > T.send_ptr[d])
					st_release_sys(which ? &T.ctrl[d]->xh_seq[T.rank] : &T.ctrl[d]->halo_seq[T.rank], seq);
		}
	}
}

__global__ void __launch_bounds__(32)
halo_push_kernel(PeerTable T, const uint32_t *__restrict__ send_idx, const double *__restrict__ v, int which,
		 unsigned long long seq, unsigned int *ticket)
{
	pdl_wait();
	pdl_launch_dependents();
	if (*(volatile int *)&T.ctrl[T.rank]->error)
		return;
	push_halo_piece(T, send_idx, v, which, seq, ticket, blockIdx.x, gridDim.x);
}

// post NV partial values into slot `rank` of every peer (called by one CTA, after its reduction)
__device__ __forceinline__ void post_pw(const PeerTable &T, double v, unsigned long long seq)
{
	if ((int)threadIdx.x < T.world) {
		st_msg(&T.ctrl[threadIdx.x]->pw_msg[T.rank], v, seq);
	}
}
__device__ __forceinline__ void post_gq(const PeerTable &T, double gg, double gq, unsigned long long seq)
{
	if ((int)threadIdx.x < T.world) {
		st_msg(&T.ctrl[threadIdx.x]->gq_msg[T.rank][0], gg, seq);
		st_msg(&T.ctrl[threadIdx.x]->gq_msg[T.rank][1], gq, seq);
	}
}

__device__ __forceinline__ uint32_t gate_slot(uint32_t k) { return k == 0 ? 0u : (k - 1) % 3u; }

__device__ __forceinline__ bool iteration_gate(uint32_t k, DistState *st)
{
	// plain loads issued together: see iteration_gate in krylov.cu (the state line is read by every
	// warp of the grid; volatile loads make it an L2 hot spot)
	const DistState *cs = st;
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

// every warp's lane 0 waits for the halo of the vector it is about to gather from
__device__ __forceinline__ bool warp_wait_halo(const PeerTable &T, int which, unsigned long long seq)
{
	DistControl *mine = T.ctrl[T.rank];
	int ok = 1;
	if ((threadIdx.x & 31) == 0)
		for (uint32_t i = 0; i < T.n_recv_src && ok; i++) {
			const int src = T.recv_src[i];
			ok = wait_seq(which ? &mine->xh_seq[src] : &mine->halo_seq[src], seq, mine, T.timeout_ns) ? 1 : 0;
		}
	return __shfl_sync(0xffffffffu, ok, 0) != 0;
}

// ---- init: g = A x - b, q = g / diag, p = -q; posts the partial (g.g, g.q) ------
template <bool JACOBI, bool BLOCKED>
__global__ void __launch_bounds__(kBlock, kStreamCtas)
dist_init_kernel(SellView A, StreamConfig cfg, PeerTable T, unsigned long long seq, unsigned long long gseq,
		 const double *__restrict__ b,
		 const double *__restrict__ x_ext, double *__restrict__ g, double *__restrict__ p,
		 double *__restrict__ q, double *__restrict__ diag, double *partials, DistState *st)
{
	extern __shared__ __align__(128) unsigned char smem[];
	double dots[2] = {0.0, 0.0};
	bool ok = true;
	sell_stream_rows<BLOCKED, JACOBI, false>(
		A, x_ext, cfg, smem, [] { return true; },
		[&] {
			ok = warp_wait_halo(T, 1, seq);
			return ok;
		},
		[&](uint32_t row, double acc, double d, double) {
			if (row < A.N) {
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
		});
	// a failed wait is raised in the window's error flag; the reduction below still
	// has to be taken by every CTA (ticket), its result is simply not used
	double tot[2];
	if (grid_reduce<2>(dots, partials, &st->ticket, tot)) {
		post_gq(T, tot[0], JACOBI ? tot[1] : tot[0], gseq);
		// consumed the input halo: let the sources push again (SpMV flow control)
		if ((int)threadIdx.x < T.world)
			st_release_sys(&T.ctrl[threadIdx.x]->xh_ack[T.rank], seq);
	}
	(void)ok;
}

// sums the per-rank partials of the init kernel into the state (one warp)
__global__ void dist_init_reduce_kernel(PeerTable T, unsigned long long seq, DistState *st)
{
	pdl_wait();
	pdl_launch_dependents();
	DistControl *mine = T.ctrl[T.rank];
	double tot[2];
	if (!cta_reduce_msgs<2>(T.world, [&](int r, int c) { return &mine->gq_msg[r][c]; }, seq, mine, T.timeout_ns,
				nullptr, tot))
		return;
	if (threadIdx.x == 0) {
		st->gg[0] = tot[0];
		st->gq[0] = tot[1];
	}
}

// ---- K1: halo wait, gate, w = A p, posts the partial p.w ---------------------------
template <bool BLOCKED>
__global__ void __launch_bounds__(kBlock, kStreamCtas)
dist_spmv_kernel(uint32_t k, SellView A, StreamConfig cfg, PeerTable T, unsigned long long base,
		 const uint32_t *__restrict__ send_idx, const double *__restrict__ p_ext, double *__restrict__ w,
		 double *partials, DistState *st)
{
	extern __shared__ __align__(128) unsigned char smem[];
	double dots[1] = {0.0};
	bool active = true;
	sell_stream_rows<BLOCKED, false, false>(
		A, p_ext, cfg, smem,
		[&] {
			const int failed = T.ctrl[T.rank]->error;   // plain: a stale 0 only delays the exit
			active = iteration_gate(k, st) && !failed;
			// The boundary entries of p_k leave at the START of the kernel that consumes p_k, while the
			// other warps already stream the matrix; the neighbours only need them for their last slices
			// (late wait below).  A push costs its warp a system-scope fence (an NVLink round trip), so it
			// is given to the LAST warp of the LAST CTAs, 32 values each: with slices dealt round-robin
			// those warps own one slice fewer than the first ones whenever the division leaves a rest.
			const uint32_t total = T.send_ptr[T.world];
			if (active && total > 0) {
				if (T.push_first) {
					if ((threadIdx.x >> 5) == 0)
						push_halo_piece(T, send_idx, p_ext, 0, base + k + 1, &st->push_ticket, blockIdx.x,
								gridDim.x);
				} else {
					const uint32_t n_push = min(gridDim.x, (total + 31u) / 32u);
					const uint32_t first_cta = gridDim.x - n_push;
					if ((threadIdx.x >> 5) == kStreamWarps - 1 && blockIdx.x >= first_cta)
						push_halo_piece(T, send_idx, p_ext, 0, base + k + 1, &st->push_ticket,
								blockIdx.x - first_cta, n_push);
				}
			}
			return active;
		},
		[&] { return warp_wait_halo(T, 0, base + k + 1); },   // only the slices that read halo columns wait
		[&](uint32_t row, double acc, double, double p_row) {
			if (row < A.N) {
				w[row] = acc;
				dots[0] = __dadd_rn(dots[0], __dmul_rn(p_row, acc));
			}
		});
	// `active` can differ between warps only through a timeout, which also sets the
	// error flag; the ticketed reduction must be taken by all CTAs that got past the gate
	__shared__ int s_gate;
	if (threadIdx.x == 0)
		s_gate = iteration_gate(k, st) ? 1 : 0;
	__syncthreads();
	if (!s_gate)
		return;
	double tot[1];
	if (grid_reduce<1>(dots, partials, &st->ticket, tot))
		post_pw(T, tot[0], base + k + 1);
}

// ---- K2: waits for all p.w, update, posts the partial (g.g, g.q) --------------------
template <bool JACOBI>
__global__ void __launch_bounds__(kBlock)
dist_update_kernel(uint32_t k, uint32_t N, PeerTable T, unsigned long long base,
		   const double *__restrict__ w, const double *__restrict__ diag,
		   double *__restrict__ g, double *__restrict__ q, double *partials, DistState *st)
{
	const uint32_t stride = gridDim.x * blockDim.x;
	const uint32_t base_i = blockIdx.x * blockDim.x + threadIdx.x;
	double g0 = 0, d0 = 1, g1 = 0, d1 = 1;
	if (base_i < N) {
		const uint32_t j1 = base_i + stride < N ? base_i + stride : base_i;
		g0 = g[base_i];
		g1 = g[j1];
		if (JACOBI) {
			d0 = diag[base_i];
			d1 = diag[j1];
		}
	}
	pdl_wait();
	pdl_launch_dependents();
	DistControl *mine = T.ctrl[T.rank];
	// w of the first trip is fetched while the peers' p.w partials are still travelling
	double wf0 = 0, wf1 = 0;
	if (base_i < N) {
		wf0 = w[base_i];
		wf1 = w[base_i + stride < N ? base_i + stride : base_i];
	}
	double pw_tot[1];
	if (!cta_reduce_msgs<1>(T.world, [&](int r, int) { return &mine->pw_msg[r]; }, base + k + 1, mine, T.timeout_ns,
				&st->done, pw_tot))
		return;
	const double pw = pw_tot[0];
	if (blockIdx.x == 0 && threadIdx.x == 0)
		st->pw = pw;
	const double alpha = __ddiv_rn(st->gq[k & 1], pw);
	double dots[2] = {0.0, 0.0};
	for (uint32_t i0 = base_i; i0 < N; i0 += 2 * stride) {
		const uint32_t i1 = i0 + stride;
		const bool has1 = i1 < N;
		const uint32_t j1 = has1 ? i1 : i0;
		if (i0 != base_i) {
			g0 = g[i0];
			g1 = g[j1];
			if (JACOBI) {
				d0 = diag[i0];
				d1 = diag[j1];
			}
		}
		const double w0 = i0 == base_i ? wf0 : w[i0], w1 = i0 == base_i ? wf1 : w[j1];
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
	double tot[2];
	if (grid_reduce<2>(dots, partials, &st->ticket, tot))
		post_gq(T, tot[0], JACOBI ? tot[1] : tot[0], base + k + 2);
}

// ---- K3: waits for all (g.g, g.q), x += alpha p, p = -q + beta p --------------------
// x_first: the x update needs only alpha (known since K2) and can run BEFORE the wait, while the partials
// of the peers travel; p is then read a second time.  Measured neutral (2 GPUs) to slower (1 GPU): off.
__global__ void __launch_bounds__(kBlock)
dist_dir_kernel(uint32_t k, uint32_t N, PeerTable T, unsigned long long base, const double *__restrict__ q,
		double *__restrict__ p, double *__restrict__ x, DistState *st, int x_first)
{
	const uint32_t stride = gridDim.x * blockDim.x;
	const uint32_t base_i = blockIdx.x * blockDim.x + threadIdx.x;
	double p0 = 0, p1 = 0, x0 = 0, x1 = 0;
	if (base_i < N) {
		const uint32_t j1 = base_i + stride < N ? base_i + stride : base_i;
		p0 = p[base_i];
		p1 = p[j1];
		x0 = x[base_i];
		x1 = x[j1];
	}
	pdl_wait();
	pdl_launch_dependents();
	DistControl *mine = T.ctrl[T.rank];
	const int32_t done = st->done;    // plain loads: every thread of the grid reads these two words
	const int failed = mine->error;
	if (done || failed)
		return;   // grid-uniform: set by the gate of K1(k) / by a failed wait earlier in the stream
	const double alpha = __ddiv_rn(st->gq[k & 1], st->pw);   // the update kernel stored pw
	for (uint32_t i0 = base_i; x_first && i0 < N; i0 += 2 * stride) {
		const uint32_t i1 = i0 + stride;
		const bool has1 = i1 < N;
		const uint32_t j1 = has1 ? i1 : i0;
		if (i0 != base_i) {
			p0 = p[i0];
			p1 = p[j1];
			x0 = x[i0];
			x1 = x[j1];
		}
		x[i0] = __dadd_rn(x0, __dmul_rn(alpha, p0));
		if (has1)
			x[i1] = __dadd_rn(x1, __dmul_rn(alpha, p1));
	}
	double tot[2];
	if (!cta_reduce_msgs<2>(T.world, [&](int r, int c) { return &mine->gq_msg[r][c]; }, base + k + 2, mine,
				T.timeout_ns, &st->done, tot))
		return;
	const double gg = tot[0], gq = tot[1];
	const double beta = __ddiv_rn(gq, st->gq[k & 1]);
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		st->gg[(k + 1) % 3u] = gg;
		st->gq[(k + 1) & 1] = gq;
	}
	for (uint32_t i0 = base_i; i0 < N; i0 += 2 * stride) {
		const uint32_t i1 = i0 + stride;
		const bool has1 = i1 < N;
		const uint32_t j1 = has1 ? i1 : i0;
		const double q0 = q[i0], q1 = q[j1];
		// the first trip's p and x were fetched before the dependency wait (the x_first pass reuses
		// those registers for its later trips, so it reloads)
		const bool pre = !x_first && i0 == base_i;
		const double pa = pre ? p0 : p[i0], pb = pre ? p1 : p[j1];
		if (!x_first) {
			const double xa = pre ? x0 : x[i0], xb = pre ? x1 : x[j1];
			x[i0] = __dadd_rn(xa, __dmul_rn(alpha, pa));
			if (has1)
				x[i1] = __dadd_rn(xb, __dmul_rn(alpha, pb));
		}
		p[i0] = __dadd_rn(-q0, __dmul_rn(beta, pa));
		if (has1)
			p[i1] = __dadd_rn(-q1, __dmul_rn(beta, pb));
	}
}

// ---- distributed SpMV (config 3): y = A x with the halo of x exchanged first --------
template <bool BLOCKED>
__global__ void __launch_bounds__(kBlock, kStreamCtas)
dist_plain_spmv_kernel(SellView A, StreamConfig cfg, PeerTable T, unsigned long long seq,
		       const double *__restrict__ x_ext, double *__restrict__ y, unsigned int *ticket)
{
	extern __shared__ __align__(128) unsigned char smem[];
	sell_stream_rows<BLOCKED, false, false>(
		A, x_ext, cfg, smem, [] { return true; }, [&] { return warp_wait_halo(T, 1, seq); },
		[&](uint32_t row, double acc, double, double) {
			if (row < A.N)
				y[row] = acc;
		});
	// last CTA out tells the sources that the halo tail may be overwritten
	__shared__ bool last;
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
	}
	__syncthreads();
	if (last && (int)threadIdx.x < T.world)
		st_release_sys(&T.ctrl[threadIdx.x]->xh_ack[T.rank], seq);
}

template <typename... KArgs, typename... Args>
cudaError_t launch(bool pdl, void (*kernel)(KArgs...), int grid, int block, size_t smem, Args &&...args)
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

}  // namespace

// ------------------------------------------------------------------ host objects --

struct nbgpu_dist_plan_s {
	int rank = 0, world = 1;
	std::vector<uint32_t> row_starts;      // [world + 1]
	uint32_t N_loc = 0, n_halo = 0;
	uint64_t nnz = 0;
	std::vector<uint32_t> halo_global;     // [n_halo] ascending
	std::vector<uint32_t> recv_counts;     // [world]
	std::vector<uint32_t> cols_local;      // [nnz]
	// sends (filled by nbgpu_dist_plan_set_sends)
	bool have_sends = false;
	std::vector<uint32_t> send_ptr;        // [world + 1]
	std::vector<uint32_t> send_local;      // local row ids, grouped by destination
	std::vector<uint32_t> dst_offset;      // [world] where my block starts in the destination's halo
	uint32_t *d_send_idx = nullptr;
	// SpMV visit order: start right after the last slice that reads halo columns, so that every
	// halo-reading slice is visited at the end (visit index >= late_from) -- the halo wait of a
	// kernel is then hidden behind the interior slices
	uint32_t visit_shift = 0, late_from = 0;
};

struct nbgpu_dist_s {
	int rank = 0, world = 1;
	size_t ext_len = 0;                    // capacity of the two ext vectors (owned + halo)
	void *window = nullptr;                // control block | p_ext | x_ext
	size_t window_bytes = 0;
	void *peer_window[kMaxRanks] = {};
	bool connected = false;
	unsigned long long seq_base = 0;       // advances identically on all ranks
	unsigned long long spmv_seq = 0;
	DistState *d_state = nullptr;
	DistState *h_state = nullptr;          // pinned, 4 slots
	unsigned int *d_ticket = nullptr;
	cudaEvent_t poll_ev[2] = {nullptr, nullptr};
	DistControl *ctrl() const { return (DistControl *)window; }
	double *p_ext() const { return (double *)((char *)window + 4096); }
	double *x_ext() const { return p_ext() + ext_len; }
	double *peer_p_ext(int r) const { return (double *)((char *)peer_window[r] + 4096); }
	double *peer_x_ext(int r) const { return peer_p_ext(r) + ext_len_of[r]; }
	size_t ext_len_of[kMaxRanks] = {};
};

static_assert(sizeof(DistControl) <= 4096, "control block must fit its page");

extern "C" {

int nbgpu_dist_plan_create(int rank, int world, const uint32_t *row_starts, const uint32_t *rows_size,
			   const uint32_t *cols_global, nbgpu_dist_plan_t **out)
{
	NB_ARG(out != nullptr && row_starts != nullptr && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world);
	nbgpu_dist_plan_t *P = new nbgpu_dist_plan_t();
	P->rank = rank;
	P->world = world;
	P->row_starts.assign(row_starts, row_starts + world + 1);
	const uint32_t r0 = row_starts[rank], r1 = row_starts[rank + 1];
	P->N_loc = r1 - r0;
	uint64_t nnz = 0;
	for (uint32_t i = 0; i < P->N_loc; i++)
		nnz += rows_size[i];
	P->nnz = nnz;
	// halo = remote columns referenced by my rows, ascending global id (=> grouped by owner)
	std::vector<uint32_t> halo;
	for (uint64_t k = 0; k < nnz; k++)
		if (cols_global[k] < r0 || cols_global[k] >= r1)
			halo.push_back(cols_global[k]);
	std::sort(halo.begin(), halo.end());
	halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
	P->n_halo = (uint32_t)halo.size();
	P->halo_global = halo;
	P->recv_counts.assign(world, 0);
	{
		int owner = 0;
		for (uint32_t h : halo) {
			if (h >= row_starts[world]) {
				delete P;
				set_error("column %u outside the global matrix", h);
				return NBGPU_ERR_ARG;
			}
			while (h >= row_starts[owner + 1])
				owner++;
			P->recv_counts[owner]++;
		}
	}
	// local numbering: owned columns first, then the halo in list order; entry order untouched
	P->cols_local.resize(nnz);
#pragma omp parallel for schedule(static)
	for (int64_t k = 0; k < (int64_t)nnz; k++) {
		const uint32_t c = cols_global[k];
		if (c >= r0 && c < r1)
			P->cols_local[k] = c - r0;
		else
			P->cols_local[k] = P->N_loc + (uint32_t)(std::lower_bound(halo.begin(), halo.end(), c) - halo.begin());
	}
	// largest (circular) run of slices that read no halo column = the interior
	{
		const uint32_t n_slices = (P->N_loc + kSliceRows - 1) / kSliceRows;
		std::vector<uint8_t> reads_halo(n_slices, 0);
		uint64_t k = 0;
		for (uint32_t i = 0; i < P->N_loc; i++)
			for (uint32_t j = 0; j < rows_size[i]; j++, k++)
				if (P->cols_local[k] >= P->N_loc)
					reads_halo[i / kSliceRows] = 1;
		uint32_t best_start = 0, best_len = 0;
		bool any = false;
		for (uint32_t s0 = 0; s0 < n_slices; s0++) {
			if (!reads_halo[s0])
				continue;
			any = true;
			// run of clean slices that starts right after halo slice s0 (circularly)
			uint32_t len = 0;
			while (len < n_slices && !reads_halo[(s0 + 1 + len) % n_slices])
				len++;
			if (len > best_len || best_len == 0) {
				if (len >= best_len) {
					best_len = len;
					best_start = (s0 + 1) % n_slices;
				}
			}
		}
		if (!any) {
			P->visit_shift = 0;
			P->late_from = 0xFFFFFFFFu;   // nothing to wait for
		} else {
			P->visit_shift = best_start;
			P->late_from = best_len;
		}
	}
	*out = P;
	return NBGPU_OK;
}

int nbgpu_dist_plan_info(const nbgpu_dist_plan_t *P, uint32_t *N_loc, uint32_t *n_halo, uint64_t *nnz,
			 uint32_t *recv_counts)
{
	NB_ARG(P != nullptr);
	if (N_loc)
		*N_loc = P->N_loc;
	if (n_halo)
		*n_halo = P->n_halo;
	if (nnz)
		*nnz = P->nnz;
	if (recv_counts)
		memcpy(recv_counts, P->recv_counts.data(), P->world * sizeof(uint32_t));
	return NBGPU_OK;
}

int nbgpu_dist_plan_halo_ids(const nbgpu_dist_plan_t *P, uint32_t *halo_global)
{
	NB_ARG(P != nullptr && (halo_global != nullptr || P->n_halo == 0));
	if (P->n_halo)
		memcpy(halo_global, P->halo_global.data(), (size_t)P->n_halo * sizeof(uint32_t));
	return NBGPU_OK;
}

int nbgpu_dist_plan_local_cols(const nbgpu_dist_plan_t *P, uint32_t *cols_local)
{
	NB_ARG(P != nullptr && (cols_local != nullptr || P->nnz == 0));
	if (P->nnz)
		memcpy(cols_local, P->cols_local.data(), (size_t)P->nnz * sizeof(uint32_t));
	return NBGPU_OK;
}

int nbgpu_dist_plan_set_sends(nbgpu_dist_plan_t *P, const uint32_t *send_counts, const uint32_t *send_global,
			      const uint32_t *dst_offsets)
{
	NB_ARG(P != nullptr && send_counts != nullptr && dst_offsets != nullptr);
	const uint32_t r0 = P->row_starts[P->rank], r1 = P->row_starts[P->rank + 1];
	P->send_ptr.assign(P->world + 1, 0);
	for (int r = 0; r < P->world; r++)
		P->send_ptr[r + 1] = P->send_ptr[r] + send_counts[r];
	const uint32_t total = P->send_ptr[P->world];
	NB_ARG(total == 0 || send_global != nullptr);
	P->send_local.resize(total);
	for (uint32_t j = 0; j < total; j++) {
		NB_ARG(send_global[j] >= r0 && send_global[j] < r1);
		P->send_local[j] = send_global[j] - r0;
	}
	P->dst_offset.assign(dst_offsets, dst_offsets + P->world);
	P->have_sends = true;
	return NBGPU_OK;
}

int nbgpu_dist_plan_destroy(nbgpu_dist_plan_t *P)
{
	if (!P)
		return NBGPU_OK;
	if (P->d_send_idx && ctx().ready)
		cudaFree(P->d_send_idx);
	delete P;
	return NBGPU_OK;
}

int nbgpu_dist_create(int rank, int world, size_t ext_len, void *ipc_handle_out, nbgpu_dist_t **out)
{
	NB_INIT();
	NB_ARG(out != nullptr && ipc_handle_out != nullptr && world >= 1 && world <= kMaxRanks && rank >= 0 &&
	       rank < world);
	nbgpu_dist_t *D = new nbgpu_dist_t();
	D->rank = rank;
	D->world = world;
	D->ext_len = (ext_len + 1) & ~(size_t)1;
	D->window_bytes = 4096 + 2 * D->ext_len * sizeof(double);
	cudaError_t e = cudaMalloc(&D->window, D->window_bytes);
	if (e == cudaSuccess)
		e = cudaMemset(D->window, 0, D->window_bytes);
	if (e == cudaSuccess)
		e = cudaMalloc(&D->d_state, sizeof(DistState));
	if (e == cudaSuccess)
		e = cudaMalloc(&D->d_ticket, 4 * sizeof(unsigned int));
	if (e == cudaSuccess)
		e = cudaMemset(D->d_ticket, 0, 4 * sizeof(unsigned int));
	if (e == cudaSuccess)
		e = cudaMallocHost(&D->h_state, 4 * sizeof(DistState) + 64);
	if (e == cudaSuccess)
		e = cudaEventCreateWithFlags(&D->poll_ev[0], cudaEventDisableTiming);
	if (e == cudaSuccess)
		e = cudaEventCreateWithFlags(&D->poll_ev[1], cudaEventDisableTiming);
	cudaIpcMemHandle_t h;
	if (e == cudaSuccess)
		e = cudaIpcGetMemHandle(&h, D->window);
	if (e != cudaSuccess) {
		set_error("nbgpu_dist_create: %s", cudaGetErrorString(e));
		cudaGetLastError();
		delete D;
		return NBGPU_ERR_COMM;
	}
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	memcpy(ipc_handle_out, &h, 64);
	*out = D;
	return NBGPU_OK;
}

/* all_handles: world x 64 bytes; all_ext_len: the ext_len every rank passed to nbgpu_dist_create */
int nbgpu_dist_connect(nbgpu_dist_t *D, const void *all_handles, const uint64_t *all_ext_len)
{
	NB_INIT();
	NB_ARG(D != nullptr && all_handles != nullptr && all_ext_len != nullptr);
	for (int r = 0; r < D->world; r++) {
		D->ext_len_of[r] = ((size_t)all_ext_len[r] + 1) & ~(size_t)1;
		if (r == D->rank) {
			D->peer_window[r] = D->window;
			continue;
		}
		cudaIpcMemHandle_t h;
		memcpy(&h, (const char *)all_handles + 64 * r, 64);
		cudaError_t e = cudaIpcOpenMemHandle(&D->peer_window[r], h, cudaIpcMemLazyEnablePeerAccess);
		if (e != cudaSuccess) {
			set_error("cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
			cudaGetLastError();
			return NBGPU_ERR_COMM;
		}
	}
	D->connected = true;
	return NBGPU_OK;
}

int nbgpu_dist_destroy(nbgpu_dist_t *D)
{
	if (!D)
		return NBGPU_OK;
	if (ctx().ready) {
		cudaSetDevice(ctx().device);
		cudaStreamSynchronize(ctx().stream);
		for (int r = 0; r < D->world; r++)
			if (r != D->rank && D->peer_window[r])
				cudaIpcCloseMemHandle(D->peer_window[r]);
		cudaFree(D->window);
		cudaFree(D->d_state);
		cudaFree(D->d_ticket);
		cudaFreeHost(D->h_state);
		cudaEventDestroy(D->poll_ev[0]);
		cudaEventDestroy(D->poll_ev[1]);
	}
	delete D;
	return NBGPU_OK;
}

}  // extern "C"

namespace {

int build_peer_table(nbgpu_dist_t *D, nbgpu_dist_plan_t *P, PeerTable *T)
{
	NB_ARG(D->connected && P->have_sends && P->rank == D->rank && P->world == D->world);
	NB_ARG((size_t)P->N_loc + P->n_halo <= D->ext_len);
	memset(T, 0, sizeof(*T));
	T->world = D->world;
	T->rank = D->rank;
	const char *to = getenv("NBGPU_DIST_TIMEOUT_MS");
	T->timeout_ns = (unsigned long long)(to ? atoll(to) : 10000) * 1000000ull;
	for (int r = 0; r < D->world; r++) {
		T->ctrl[r] = (DistControl *)D->peer_window[r];
		const uint32_t n_loc_r = P->row_starts[r + 1] - P->row_starts[r];
		T->p_halo_dst[r] = D->peer_p_ext(r) + n_loc_r + P->dst_offset[r];
		T->x_halo_dst[r] = D->peer_x_ext(r) + n_loc_r + P->dst_offset[r];
		T->send_ptr[r] = P->send_ptr[r];
		if (P->recv_counts[r] > 0)
			T->recv_src[T->n_recv_src++] = r;
	}
	T->send_ptr[D->world] = P->send_ptr[D->world];
	if (!P->d_send_idx) {
		NB_CUDA(cudaMalloc(&P->d_send_idx, std::max<size_t>(1, P->send_local.size()) * sizeof(uint32_t)));
		NB_CUDA(cudaMemcpy(P->d_send_idx, P->send_local.data(), P->send_local.size() * sizeof(uint32_t),
				   cudaMemcpyHostToDevice));
	}
	return NBGPU_OK;
}

int n_destinations(const nbgpu_dist_plan_t *P)
{
	int n = 0;
	for (int r = 0; r < P->world; r++)
		n += P->send_ptr[r + 1] > P->send_ptr[r];
	return n;
}

int dist_solve(nbgpu_dist_t *D, nbgpu_dist_plan_t *P, const nbgpu_matrix_t *A, const double *d_b, double *d_x,
	       uint32_t max_iter, double tol, uint32_t *niter, double *tol_reached, bool jacobi)
{
	NB_INIT();
	NB_ARG(D != nullptr && P != nullptr && A != nullptr && d_b != nullptr && d_x != nullptr);
	NB_ARG(A->N == P->N_loc && A->n_cols == P->N_loc + P->n_halo);
	Context &c = ctx();
	PeerTable T;
	NB_TRY(build_peer_table(D, P, &T));
	const uint32_t N = A->N;
	const size_t Np = ((size_t)N + 1) & ~(size_t)1;
	NB_TRY(ensure_workspace(5 * Np * sizeof(double)));
	double *xw = c.ws, *g = xw + Np, *w = g + Np;
	double *q = jacobi ? w + Np : g, *diag = jacobi ? q + Np : nullptr;
	double *p = D->p_ext(), *x_ext = D->x_ext();
	// No persisting-L2 window here: with p in the IPC window (outside the work-vector block) a window
	// over the other vectors measured SLOWER than the default policy (51.2 vs 49.6 us/iteration, 1 M dof).
	DistState *st = D->d_state, *hst = D->h_state;
	int *herr = (int *)(hst + 4);   // pinned, behind the four state slots
	herr[0] = herr[1] = herr[2] = 0;
	memset(&hst[2], 0, sizeof(DistState));
	hst[2].tol2 = tol * tol;
	hst[2].max_iter = max_iter;
	NB_CUDA(cudaMemcpyAsync(st, &hst[2], sizeof(DistState), cudaMemcpyHostToDevice, c.stream));
	NB_CUDA(cudaMemcpyAsync(xw, d_x, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
	NB_CUDA(cudaMemcpyAsync(x_ext, d_x, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));

	SellView V;
	V.N = N; V.n_slices = A->n_slices; V.slice_off = A->d_slice_off; V.val = A->d_val;
	V.col = A->blocked ? A->d_bcol : A->d_col;
	V.uniform_width = A->uniform_width;
	V.visit_shift = P->visit_shift; V.late_from = P->late_from;
	StreamConfig scfg, icfg;
	const void *sk = A->blocked ? (const void *)dist_spmv_kernel<true> : (const void *)dist_spmv_kernel<false>;
	const void *ik = jacobi ? (A->blocked ? (const void *)dist_init_kernel<true, true>
					      : (const void *)dist_init_kernel<true, false>)
				: (A->blocked ? (const void *)dist_init_kernel<false, true>
					      : (const void *)dist_init_kernel<false, false>);
	if (!stream_config(A, sk, &scfg) || !stream_config(A, ik, &icfg)) {
		set_error("distributed solve needs the streamed SpMV path (slice too wide or NBGPU_SPMV_PATH=reg)");
		return NBGPU_ERR_ARG;
	}
	const int64_t vec_blocks = ((int64_t)N + 2 * kBlock - 1) / (2 * kBlock);
	const int ugrid = jacobi ? resident_grid(dist_update_kernel<true>, vec_blocks)
				 : resident_grid(dist_update_kernel<false>, vec_blocks);
	const int dgrid = resident_grid(dist_dir_kernel, vec_blocks);
	const int n_dst = n_destinations(P);
	const bool pdl = !getenv("NBGPU_NO_PDL");
	// x += alpha p before the wait for the peers' partials (p is then read twice) measured neutral on
	// 2 GPUs and 0.9 us slower on 1; off unless asked for
	const int x_first = getenv("NBGPU_DIST_X_FIRST") ? 1 : 0;
	T.push_first = getenv("NBGPU_DIST_PUSH_FIRST") ? 1 : 0;
	// the x-halo exchange shares its flags with nbgpu_dist_spmv: one counter for both
	const unsigned long long xseq = ++D->spmv_seq;
	const unsigned long long base = D->seq_base;

	// x halo -> init
	const int push_grid = std::max(1, (int)std::min<uint32_t>((P->send_ptr[P->world] + 31) / 32, 1024));
	if (n_dst)
		NB_CUDA(launch(false, halo_push_kernel, push_grid, 32, 0, T, (const uint32_t *)P->d_send_idx,
			       (const double *)x_ext, 1, xseq, D->d_ticket + 1));
	cudaError_t e;
	// the init kernel waits for the x halo (xseq) and posts its dots as sequence base + 1
	if (jacobi && A->blocked)
		e = launch(false, dist_init_kernel<true, true>, icfg.grid, kBlock, icfg.smem_bytes, V, icfg, T, xseq, base + 1, d_b,
			   x_ext, g, p, q, diag, c.partials, st);
	else if (jacobi)
		e = launch(false, dist_init_kernel<true, false>, icfg.grid, kBlock, icfg.smem_bytes, V, icfg, T, xseq, base + 1, d_b,
			   x_ext, g, p, q, diag, c.partials, st);
	else if (A->blocked)
		e = launch(false, dist_init_kernel<false, true>, icfg.grid, kBlock, icfg.smem_bytes, V, icfg, T, xseq, base + 1, d_b,
			   x_ext, g, p, q, diag, c.partials, st);
	else
		e = launch(false, dist_init_kernel<false, false>, icfg.grid, kBlock, icfg.smem_bytes, V, icfg, T, xseq, base + 1, d_b,
			   x_ext, g, p, q, diag, c.partials, st);
	NB_CUDA(e);
	NB_CUDA(launch(false, dist_init_reduce_kernel, 1, 32, 0, T, base + 1, st));
	uint32_t k = 0;
	int slot = 0;
	bool pending[2] = {false, false};
	bool finished = false;
	while (!finished) {
		const uint32_t k_end = (uint32_t)std::min<uint64_t>(max_iter, (uint64_t)k + 32);
		for (; k < k_end; k++) {
			if (A->blocked)
				e = launch(pdl, dist_spmv_kernel<true>, scfg.grid, kBlock, scfg.smem_bytes, k, V, scfg, T, base,
					   (const uint32_t *)P->d_send_idx, (const double *)p, w, c.partials, st);
			else
				e = launch(pdl, dist_spmv_kernel<false>, scfg.grid, kBlock, scfg.smem_bytes, k, V, scfg, T, base,
					   (const uint32_t *)P->d_send_idx, (const double *)p, w, c.partials, st);
			NB_CUDA(e);
			if (jacobi)
				e = launch(pdl, dist_update_kernel<true>, ugrid, kBlock, 0, k, N, T, base, (const double *)w,
					   (const double *)diag, g, q, c.partials, st);
			else
				e = launch(pdl, dist_update_kernel<false>, ugrid, kBlock, 0, k, N, T, base, (const double *)w,
					   (const double *)diag, g, q, c.partials, st);
			NB_CUDA(e);
			NB_CUDA(launch(pdl, dist_dir_kernel, dgrid, kBlock, 0, k, N, T, base, (const double *)q, p, xw, st, x_first));
		}
		if (k == max_iter) {
			if (A->blocked)
				e = launch(false, dist_spmv_kernel<true>, scfg.grid, kBlock, scfg.smem_bytes, k, V, scfg, T, base,
					   (const uint32_t *)P->d_send_idx, (const double *)p, w, c.partials, st);
			else
				e = launch(false, dist_spmv_kernel<false>, scfg.grid, kBlock, scfg.smem_bytes, k, V, scfg, T, base,
					   (const uint32_t *)P->d_send_idx, (const double *)p, w, c.partials, st);
			NB_CUDA(e);
		}
		NB_CUDA(cudaMemcpyAsync(&hst[slot], st, sizeof(DistState), cudaMemcpyDeviceToHost, c.stream));
		NB_CUDA(cudaMemcpyAsync(&herr[slot], &D->ctrl()->error, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
		NB_CUDA(cudaEventRecord(D->poll_ev[slot], c.stream));
		pending[slot] = true;
		const int other = slot ^ 1;
		if (pending[other]) {
			NB_CUDA(cudaEventSynchronize(D->poll_ev[other]));
			if (hst[other].done || herr[other])   // converged, or a wait timed out
				finished = true;
		}
		if (k == max_iter)
			finished = true;
		slot ^= 1;
	}
	NB_CUDA(cudaMemcpyAsync(d_x, xw, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
	NB_CUDA(cudaMemcpyAsync(&hst[3], st, sizeof(DistState), cudaMemcpyDeviceToHost, c.stream));
	NB_CUDA(cudaMemcpyAsync(&herr[2], &D->ctrl()->error, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
	NB_CUDA(cudaStreamSynchronize(c.stream));
	const int err = herr[2];
	if (err || !hst[3].done) {
		set_error("distributed solve: %s", err ? "a peer did not answer within the timeout" : "gate did not fire");
		return NBGPU_ERR_COMM;
	}
	// all ranks saw the same k_final: advance the sequence space identically everywhere
	D->seq_base = base + (unsigned long long)hst[3].k_final + 8;
	if (niter)
		*niter = hst[3].k_final;
	if (tol_reached)
		*tol_reached = sqrt(hst[3].gg_final);
	return (hst[3].gg_final > hst[3].tol2) ? NBGPU_NOT_CONVERGED : NBGPU_OK;
}

}  // namespace

extern "C" {

int nbgpu_dist_pcg_jacobi(nbgpu_dist_t *D, nbgpu_dist_plan_t *P, const nbgpu_matrix_t *A_local, const double *d_b,
			  double *d_x, uint32_t max_iter, double tolerance, uint32_t *niter_performed,
			  double *tolerance_reached)
{
	return dist_solve(D, P, A_local, d_b, d_x, max_iter, tolerance, niter_performed, tolerance_reached, true);
}

int nbgpu_dist_cg(nbgpu_dist_t *D, nbgpu_dist_plan_t *P, const nbgpu_matrix_t *A_local, const double *d_b,
		  double *d_x, uint32_t max_iter, double tolerance, uint32_t *niter_performed,
		  double *tolerance_reached)
{
	return dist_solve(D, P, A_local, d_b, d_x, max_iter, tolerance, niter_performed, tolerance_reached, false);
}

int nbgpu_dist_spmv(nbgpu_dist_t *D, nbgpu_dist_plan_t *P, const nbgpu_matrix_t *A, const double *d_in,
		    double *d_out)
{
	NB_INIT();
	NB_ARG(D != nullptr && P != nullptr && A != nullptr && d_in != nullptr && d_out != nullptr);
	NB_ARG(A->N == P->N_loc && A->n_cols == P->N_loc + P->n_halo);
	Context &c = ctx();
	PeerTable T;
	NB_TRY(build_peer_table(D, P, &T));
	SellView V;
	V.N = A->N; V.n_slices = A->n_slices; V.slice_off = A->d_slice_off; V.val = A->d_val;
	V.col = A->blocked ? A->d_bcol : A->d_col;
	V.uniform_width = A->uniform_width;
	V.visit_shift = P->visit_shift; V.late_from = P->late_from;
	StreamConfig cfg;
	const void *kern = A->blocked ? (const void *)dist_plain_spmv_kernel<true>
				      : (const void *)dist_plain_spmv_kernel<false>;
	if (!stream_config(A, kern, &cfg)) {
		set_error("distributed SpMV needs the streamed path");
		return NBGPU_ERR_ARG;
	}
	const unsigned long long seq = ++D->spmv_seq;
	double *x_ext = D->x_ext();
	if (d_in != x_ext)   // callers that fill nbgpu_dist_input_vector() directly skip this copy
		NB_CUDA(cudaMemcpyAsync(x_ext, d_in, (size_t)A->N * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
	const int n_dst = n_destinations(P);
	const int push_grid = std::max(1, (int)std::min<uint32_t>((P->send_ptr[P->world] + 31) / 32, 1024));
	if (n_dst)
		NB_CUDA(launch(false, halo_push_kernel, push_grid, 32, 0, T, (const uint32_t *)P->d_send_idx,
			       (const double *)x_ext, 1, seq, D->d_ticket + 1));
	cudaError_t e;
	if (A->blocked)
		e = launch(false, dist_plain_spmv_kernel<true>, cfg.grid, kBlock, cfg.smem_bytes, V, cfg, T, seq,
			   (const double *)x_ext, d_out, D->d_ticket);
	else
		e = launch(false, dist_plain_spmv_kernel<false>, cfg.grid, kBlock, cfg.smem_bytes, V, cfg, T, seq,
			   (const double *)x_ext, d_out, D->d_ticket);
	NB_CUDA(e);
	return NBGPU_OK;
}

/* the window's input vector (owned part): SpMV callers may write x here and pass it as d_in */
double *nbgpu_dist_input_vector(nbgpu_dist_t *D)
{
	return D ? D->x_ext() : nullptr;
}

/* communication error raised by a kernel wait (0 = none) */
int nbgpu_dist_error(nbgpu_dist_t *D)
{
	NB_INIT();
	NB_ARG(D != nullptr);
	int err = 0;
	NB_CUDA(cudaMemcpy(&err, &D->ctrl()->error, sizeof(int), cudaMemcpyDeviceToHost));
	return err ? NBGPU_ERR_COMM : NBGPU_OK;
}

}  // extern "C"
