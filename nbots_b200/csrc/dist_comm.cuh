// dist_comm.cuh -- the exchange layer of the row-partitioned solvers: what a kernel needs to
// talk to the other GPUs of the node over NVLink peer memory.  No collective library is involved.
//
// Every rank owns one "window" allocation (control block + the two vectors that have halo parts);
// peers map it (CUDA IPC between processes, peer access inside one process) and
//   * push the boundary entries of the gathered vector straight into the neighbours' halo parts
//     (plain stores over NVLink), followed by a system-scope release of a sequence number;
//   * post their partial dot products (the rank's total, sent by one CTA of the producing or of the
//     consuming kernel) as self-validating 16-byte messages into one slot per rank of every peer's
//     control block.
// Consumers poll THEIR OWN memory.  All ranks add the per-rank partials in rank order, so alpha,
// beta and the stopping decision are bit-identical on every rank.
//
// Message format: {bits(v), bits(v) ^ mix(seq)}.  A reader accepts the pair only when
// word0 ^ word1 == mix(seq): the two 8-byte halves may arrive in any order (PTX gives no 16-byte
// single-copy atomicity), a torn pair (one old, one new half) validates only if old and new value
// are equal, in which case it is right.  Slots alternate with seq & 1, so a fast rank's message
// seq + 1 never overwrites a slot a slow peer still polls for seq (it cannot post seq + 2 before
// every peer consumed seq, because seq + 1 needs every peer's contribution).
//
// Every wait has a wall-clock timeout (NBGPU_DIST_TIMEOUT_MS, default 10000): on expiry the
// window's error flag is raised, all later kernels of the solve return at once and the host
// reports NBGPU_ERR_COMM -- a lost peer cannot hang the GPU.
#pragma once

#include "common.cuh"

namespace nbgpu {

constexpr int kMaxRanks = 16;
constexpr int kMaxMsgValues = 4;
constexpr uint32_t kExtAlign = 16;   // doubles: halo parts start on their own 128-byte lines

struct DistControl {
	unsigned long long halo_seq[kMaxRanks];   // Krylov gather-vector halo pushed by rank src
	unsigned long long xh_seq[kMaxRanks];     // input-vector halo (x at init, SpMV input)
	unsigned long long xh_ack[kMaxRanks];     // ... consumed by rank dst (flow control for SpMV)
	ulonglong2 msg[2][kMaxRanks][kMaxMsgValues];   // reduction partials [seq & 1][src][value]
	int error;
};
static_assert(sizeof(DistControl) <= 4096, "control block must fit its page");

struct PeerTable {
	DistControl *ctrl[kMaxRanks];      // peers' control blocks (own included), in MY address space
	double *v_halo_dst[kMaxRanks];     // where my boundary values of the Krylov vector go in peer d's ext vector
	double *x_halo_dst[kMaxRanks];     // same for the input vector
	uint32_t send_ptr[kMaxRanks + 1];  // my send list, grouped by destination
	int world, rank;
	uint32_t n_recv_src;               // ranks I receive a halo from
	int recv_src[kMaxRanks];
	unsigned long long timeout_ns;
};

#ifdef __CUDACC__
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
__device__ __forceinline__ unsigned long long msg_mix(unsigned long long seq)
{
	return (seq * 0x9E3779B97F4A7C15ull) | 1ull;
}
__device__ __forceinline__ void st_msg(ulonglong2 *p, double v, unsigned long long seq)
{
	const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
	asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(bits), "l"(bits ^ msg_mix(seq)) : "memory");
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
	const unsigned long long want = msg_mix(seq);
	unsigned long long t0 = 0;
	for (unsigned int spin = 0;; spin++) {
		const ulonglong2 m = ld_msg(slot);
		if ((m.x ^ m.y) == want) {
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

// ---- halo push ---------------------------------------------------------------------
// The send list (boundary entries of v, grouped by destination) is split evenly over
// `n_pieces` warps of the calling kernel; each copies its piece into the neighbours' halo
// parts with plain NVLink stores, fences at system scope and takes a ticket; the last
// arriver raises the destinations' flags.  which: 0 -> Krylov vector (halo_seq),
// 1 -> input vector (xh_seq; waits for the destination's ack of the previous push first,
// the input halo is single-buffered).  `v` is indexed by local row (owned part).
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
			double *dst = which ? T.x_halo_dst[d] : T.v_halo_dst[d];
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

// (cold: called once per warp from inside the streaming loop; out of line it costs that loop no registers)
static __device__ __noinline__ bool warp_wait_halo_cold(const PeerTable *T, int which, unsigned long long seq)
{
	return warp_wait_halo(*T, which, seq);
}

// CTA-wide: every rank's message `seq` from the slots in this GPU's own control block, summed in rank
// order.  Kept out of line with scalar arguments only: inlined it costs the streaming kernels ~35
// registers (a CTA per SM), as a member function the exchange policy object would be copied to every
// thread's stack.
template <int NV>
__device__ __noinline__ bool collect_slots(DistControl *mine, int world, unsigned long long timeout_ns,
					   unsigned long long seq, double (&v)[NV])
{
	__shared__ double s_val[NV][kMaxRanks];
	int ok = 1;
	if ((int)threadIdx.x < world * NV) {
		const int r = threadIdx.x / NV, c = threadIdx.x % NV;
		double got = 0.0;
		ok = wait_msg(&mine->msg[seq & 1][r][c], seq, &got, mine, timeout_ns) ? 1 : 0;
		s_val[c][r] = got;
	}
	if (!__syncthreads_and(ok))
		return false;
#pragma unroll
	for (int c = 0; c < NV; c++) {
		double t = 0.0;
		for (int r = 0; r < world; r++)
			t += s_val[c][r];
		v[c] = t;
	}
	return true;
}

// The same for a rank that already holds its own total in v (consumer-posted reductions): waits for the OTHER
// ranks' messages only, adds everything in rank order.
template <int NV>
__device__ __noinline__ bool collect_others(DistControl *mine, int world, int rank, unsigned long long timeout_ns,
					    unsigned long long seq, double (&v)[NV])
{
	__shared__ double s_val[NV][kMaxRanks];
	int ok = 1;
	if ((int)threadIdx.x < world * NV) {
		const int r = threadIdx.x / NV, c = threadIdx.x % NV;
		double got = v[0];
#pragma unroll
		for (int i = 1; i < NV; i++)
			got = (c == i) ? v[i] : got;
		if (r != rank)
			ok = wait_msg(&mine->msg[seq & 1][r][c], seq, &got, mine, timeout_ns) ? 1 : 0;
		s_val[c][r] = got;
	}
	if (!__syncthreads_and(ok))
		return false;
#pragma unroll
	for (int c = 0; c < NV; c++) {
		double t = 0.0;
		for (int r = 0; r < world; r++)
			t += s_val[c][r];
		v[c] = t;
	}
	return true;
}

// ---- the two exchange policies of the solver kernels --------------------------------
// A reduction over the ranks has a PRODUCER side (the CTA that finished the kernel's grid
// reduction) and a CONSUMER side (the next kernel):
//   post()     one fire-and-forget store per rank and value; the kernel exits behind it, so the
//              NVLink flight runs beside the kernel boundary, not in front of it;
//   collect()  every CTA of the consumer polls the slots in its own memory (one L2 round trip when
//              the messages are there -- what reading the reduced scalar costs on one GPU) and adds
//              them in rank order.
//   exchange() the CONSUMER-posted form (CLASSIC default, see PeerComm::exchange): the producer only stores
//              per-CTA partials as on one GPU, every consumer CTA adds them up, CTA 0 sends the rank's total
//              and all CTAs add the peers' totals -- no ticket pass, no NVLink drain at the producer's end.
// Measured alternatives (2 GPUs, 1 M dof per GPU, us per iteration): the producer CTA waiting for the
// peers and storing the global sum for a poll-free consumer 57.7; the producer looking once and the
// consumer polling only when that failed 58.7; post + collect (round 1's scheme) 54.8, 48.4 after K1's
// push warp; exchange() 46.6; every producer CTA sending its partial to every rank 54.7.
// all_reduce() = post + collect by one CTA (init kernel: once per solve).
// NoComm: single GPU, everything compiles away.
struct NoComm {
	static constexpr bool kDist = false;
	__device__ __forceinline__ int failed() const { return 0; }
	__device__ __forceinline__ void push_halo(const double *, int, unsigned long long, unsigned int *) const {}
	__device__ __forceinline__ bool wait_halo(int, unsigned long long) const { return true; }
	template <int NV>
	__device__ __forceinline__ bool all_reduce(double (&)[NV], unsigned long long) const { return true; }
	template <int NV>
	__device__ __forceinline__ void post(const double (&)[NV], unsigned long long) const {}
	template <int NV>
	__device__ __forceinline__ bool collect(unsigned long long, double (&)[NV]) const { return true; }
	__host__ __device__ __forceinline__ bool consumer_posts() const { return false; }
	template <int NV>
	__device__ __forceinline__ bool exchange(unsigned long long, double (&)[NV]) const { return true; }
	__device__ __forceinline__ void ack_input(unsigned long long) const {}
};

// PeerComm: row-partitioned over the GPUs of one node.  What the hot paths need travels by value
// (kernel parameter: no dependent loads in front of the first store or poll); the full peer table
// stays in device memory for the cold paths (a ~600-byte parameter indexed at run time would be
// copied to every thread's stack).
struct PeerComm {
	static constexpr bool kDist = true;
	const PeerTable *T;
	DistControl *mine;
	const uint32_t *send_idx;
	DistControl *ctrl[kMaxRanks];   // peers' control blocks (own included)
	int world, rank;
	uint32_t total_sends;
	int cpost;                      // reductions are posted by the consumer kernel (see exchange())
	unsigned long long timeout_ns;

	__device__ __forceinline__ int failed() const { return mine->error; }   // plain: a stale 0 only delays the exit

	// Called by every thread after the kernel's gate.  The boundary entries leave at the START of the
	// kernel that consumes the vector, while the other warps already stream the matrix; the neighbours
	// only need them for their late slices.  A push costs its warp 5-6 us (index load, value load, NVLink
	// stores, a system-scope fence until NVLink acknowledges, the ticket), and slices are dealt statically,
	// so WHICH warps push decides the kernel's tail.  The pushing warp is the LAST warp of each of the first
	// ceil(sends / 32) CTAs: the extra push-only warp where the kernel is launched with one
	// (krylov_kernels.cuh, PUSH_WARP: 2 GPUs 53.5 -> 48.7 us per iteration), else the last streaming warp.
	// First CTAs rather than last: the CTAs of a programmatically launched grid become resident in blockIdx
	// order as the predecessor's CTAs retire, so the first ones start earliest (53.5 -> 52.3 before the push
	// warp existed).
	__device__ __forceinline__ void push_halo(const double *v_own, int which, unsigned long long seq,
						  unsigned int *ticket) const
	{
		if (total_sends == 0)
			return;
		const uint32_t n_push = min(gridDim.x, (total_sends + 31u) / 32u);
		if ((threadIdx.x >> 5) == (blockDim.x >> 5) - 1 && blockIdx.x < n_push)
			push_piece(v_own, which, seq, ticket, blockIdx.x, n_push);
	}
	// (cold: kept out of line so that it costs the streaming loop no registers)
	__device__ __noinline__ void push_piece(const double *v_own, int which, unsigned long long seq,
						unsigned int *ticket, uint32_t piece, uint32_t n_pieces) const
	{
		push_halo_piece(*T, send_idx, v_own, which, seq, ticket, piece, n_pieces);
	}
	__device__ __forceinline__ bool wait_halo(int which, unsigned long long seq) const
	{
		return warp_wait_halo_cold(T, which, seq);
	}

	// thread (r, c) of the calling CTA posts value c to rank r
	template <int NV>
	__device__ __forceinline__ void post(const double (&v)[NV], unsigned long long seq) const
	{
		static_assert(NV <= kMaxMsgValues, "message too long");
		if ((int)threadIdx.x < world * NV) {
			const int r = threadIdx.x / NV, c = threadIdx.x % NV;
			double mine_c = v[0];
#pragma unroll
			for (int i = 1; i < NV; i++)
				mine_c = (c == i) ? v[i] : mine_c;
			st_msg(&ctrl[r]->msg[seq & 1][rank][c], mine_c, seq);
		}
	}
	// CTA-wide: wait for every rank's message `seq`, sums in rank order.  false: a wait failed
	// (timeout; the error flag is raised).
	template <int NV>
	__device__ __forceinline__ bool collect(unsigned long long seq, double (&v)[NV]) const
	{
		return collect_slots<NV>(mine, world, timeout_ns, seq, v);
	}
	template <int NV>
	__device__ __forceinline__ bool all_reduce(double (&v)[NV], unsigned long long seq) const
	{
		post<NV>(v, seq);
		return collect<NV>(seq, v);
	}
	// Consumer-posted reduction: the producer kernel's CTAs only stored their partials (as on one GPU), every
	// CTA of the consumer has just added them up (v = this rank's total, identical on every CTA).  CTA 0 sends
	// it to the peers, every CTA waits for the peers' totals and adds all in rank order.  Against the
	// ticketed producer this takes the ticket pass (fence, atomic round trip, last CTA re-reading the
	// partials) and the end-of-kernel drain of the NVLink stores off the critical path; the NVLink flight is
	// exposed in exchange.
	__host__ __device__ __forceinline__ bool consumer_posts() const { return cpost != 0; }
	template <int NV>
	__device__ __forceinline__ bool exchange(unsigned long long seq, double (&v)[NV]) const
	{
		if (world == 1)
			return true;
		if (blockIdx.x == 0 && (int)threadIdx.x < world * NV) {
			const int r = threadIdx.x / NV, c = threadIdx.x % NV;
			double mine_c = v[0];
#pragma unroll
			for (int i = 1; i < NV; i++)
				mine_c = (c == i) ? v[i] : mine_c;
			if (r != rank)
				st_msg(&ctrl[r]->msg[seq & 1][rank][c], mine_c, seq);
		}
		return collect_others<NV>(mine, world, rank, timeout_ns, seq, v);
	}
	// the input halo has been consumed: let the sources push again (SpMV flow control)
	__device__ __forceinline__ void ack_input(unsigned long long seq) const
	{
		if ((int)threadIdx.x < world)
			st_release_sys(&ctrl[threadIdx.x]->xh_ack[rank], seq);
	}
};
#endif

}  // namespace nbgpu
