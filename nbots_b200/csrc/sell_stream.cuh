// sell_stream.cuh -- the matrix stream of the SELL-32 kernels, staged through
// shared memory with bulk asynchronous copies (TMA, `cp.async.bulk`, SASS
// UBLKCP) and mbarriers.
//
// Why: SpMV here is a pure HBM stream (12 B per stored entry, used once) plus a
// cached gather.  With plain loads the bytes in flight per SM are bounded by
// registers x resident warps, and the dependent gather sits between two batches
// of streaming loads, so the DRAM pipe idles (ncu, round 1: 48 % DRAM
// utilisation at 40 % achieved occupancy).  Here every warp owns a small ring
// of shared-memory stages; one lane issues the bulk copies of the warp's NEXT
// slices (values and column ids are each one contiguous block per slice) while
// all 32 lanes consume the current one.  The bytes in flight are
// stages x slice size x warps -- independent of registers and occupancy -- and
// the consumers touch global memory only for the gather of x.
//
// A warp is fully self-contained (its own stages, its own mbarriers, producer =
// its lane 0): no block-level synchronisation anywhere in the stream.
//
// BLOCKED layout: the elasticity matrix has 2 dofs per node, so rows 2i, 2i+1
// share their column pattern and columns come in (2c, 2c+1) pairs.  When the
// pattern has that structure (checked at creation) the column stream stores one
// NODE id per 2x2 block (a quarter of the ids: 9 instead of 12 bytes per entry)
// and the gather fetches x[2c], x[2c+1] as one 16-byte load.  The accumulation
// order per row is unchanged.
//
// IDX16: for banded numberings the ids are stored as int16 differences to the
// row's own index (node index when BLOCKED), padding = -32768: 10 / 8.5 bytes
// per entry instead of 12 / 9.
//
// What bounds the kernel (ncu source page, Q1): the per-warp chain
// "ids from shared memory -> gathers of x -> ordered accumulation"; a third of
// all stall samples sit on the first use of a gathered value.  Hence 10 warps
// per CTA rather than 8 with more registers, and no early stage release in the
// blocked variant (NB_EARLY, see the loop).  Not bytes in flight, not L1/LSU
// capacity to spare either: a `prefetch.global.L1` of the next slice's gather
// targets made the kernel 17 % slower (28 KB of L1 left beside the stages).
#pragma once

#include <type_traits>

#include "spmv.cuh"

namespace nbgpu {

// ------------------------------------------------------------------ PTX ----
__device__ __forceinline__ uint32_t smem_addr(const void *p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes)
		     : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t done;
	do {
		asm volatile("{\n"
			     ".reg .pred p;\n"
			     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
			     "selp.u32 %0, 1, 0, p;\n"
			     "}\n"
			     : "=r"(done)
			     : "r"(smem_addr(bar)), "r"(parity)
			     : "memory");
	} while (!done);
}
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
	uint64_t policy;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
	return policy;
}
// global -> shared bulk copy (bytes and both addresses multiples of 16), completion on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
					 uint64_t policy)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
		     "[%0], [%1], %2, [%3], %4;" ::"r"(smem_addr(dst)),
		     "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
		     : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ double2 ld_gather_f64x2(const double *p)
{
	double2 r;
	asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
	return r;
}

// Gathers of a vector whose halo part is written by PEER GPUs while the kernel runs (dist.cu): `.nc`
// is only defined for data that stays read-only for the kernel's lifetime, and a non-coherent L1 line
// could outlive the arrival of a halo.  Slices that read halo columns (visited last, after the
// acquire of the neighbours' flags) therefore gather with `ld.relaxed.gpu` -- a strong load, served
// from L2, the point of coherence for NVLink stores into this GPU's memory; all other slices keep
// the read-only path.  The halo parts of such a vector start on their own 128-byte lines
// (nbgpu_dist_ext_layout), so an interior slice can never pull a stale halo value into L1.
__device__ __forceinline__ double ld_gather_f64_coherent(const double *p)
{
	double r;
	asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(r) : "l"(p));
	return r;
}
__device__ __forceinline__ double2 ld_gather_f64x2_coherent(const double *p)
{
	double2 r;
	asm volatile("ld.relaxed.gpu.global.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
	return r;
}

// ------------------------------------------------------------- geometry ----
constexpr int kStreamWarps = kBlock / 32;      // consumer warps per CTA
#ifndef NB_STREAM_CTAS
#define NB_STREAM_CTAS 2
#endif
constexpr int kStreamCtas = NB_STREAM_CTAS;    // CTAs per SM the streamed kernels are compiled for
constexpr int kStreamMaxStages = 6;
constexpr int kGatherBatch = 9;                // gathers in flight per lane

struct StreamConfig {
	uint32_t cap;          // stage capacity in columns (>= widest slice)
	uint32_t stages;       // ring depth per warp
	uint32_t stage_bytes;  // cap * (256 + id bytes per column): ids are 128 / 32 (blocked) / 64 / 16 (16-bit)
	uint32_t smem_bytes;   // dynamic shared memory per CTA
	int grid;              // resident CTAs
};

// bytes of column ids per stored entry-column of a slice
__host__ __device__ constexpr uint32_t stream_id_bytes(bool blocked, bool idx16)
{
	return blocked ? (idx16 ? 16u : 32u) : (idx16 ? kSliceRows * 2u : kSliceRows * 4u);
}

__host__ __device__ inline uint32_t stream_stage_bytes(uint32_t cap, bool blocked, bool idx16)
{
	return cap * (kSliceRows * 8u) + cap * stream_id_bytes(blocked, idx16);
}

constexpr int kPadDelta = -32768;   // 16-bit column ids: padding marker (valid deltas are -32767 .. 32767)

struct SellView {
	uint32_t N, n_slices;
	const uint32_t *slice_off;
	const double *val;
	const void *col;        // per-entry column ids, or per-block node ids when BLOCKED; uint32, or with
				// IDX16 int16 differences to the row's own index (node index when BLOCKED)
	// Visit order (rank-local blocks of a partitioned matrix): the v-th slice visited is
	// (v + visit_shift) mod n_slices, so that the slices whose rows read halo columns come last,
	// and `late()` -- the wait for the halo -- is only called before visit index late_from.
	uint32_t visit_shift = 0;
	uint32_t late_from = 0xFFFFFFFFu;
	uint32_t late_to = 0xFFFFFFFFu;   // visits [late_from, late_to) are the halo-reading ones
	uint32_t uniform_width = 0;   // != 0: slice s starts at s * uniform_width (no offset loads)
	// rank-local block (HALO kernels): row r is column r + col_shift of the block's column space
	// "lower halo | owned | upper halo" (even, a multiple of 16)
	uint32_t col_shift = 0;
	const uint32_t *perm = nullptr;   // SELL-C-sigma: row stored at position slice * 32 + lane (null: identity)
};

// Runs `body(row, acc, diag, x_row)` for every row of the slices this warp owns
// (slices warp, warp + total_warps, ...), acc = sum_j A[row][j] x[col_j] in
// ascending column order with separately rounded products and sums.  diag is
// A[row][row] (WANT_DIAG) and x_row is x[row]: both are picked up from the
// diagonal entry's column while it passes by (FEM rows always store their
// diagonal), which saves the solvers a second dependent load per row.
//
// The matrix never changes during a solve, so the first stages are requested
// BEFORE the programmatic-dependency wait: the bulk copies of this kernel
// overlap the tail of the previous kernel in the stream.  `gate()` is evaluated
// after the wait (it may read what the predecessor wrote); when it returns
// false the warp only drains its in-flight copies and leaves.  `late()` is
// called once, before the first slice with visit index >= A.late_from (slices
// that need data a peer GPU is still sending); false = give up likewise.
// `pre(row)` runs when a slice's stage has landed, BEFORE its gathers: a load the body needs (one
// more row-indexed vector) is issued there so that its latency hides behind the gathers; its value
// is handed to `body(row, acc, diag, x_row, pre_value)`.
// HALO: x is the extended vector of a rank-local block (col_shift, coherent gathers for the late
// slices, see ld_gather_f64_coherent).  After the gate every CTA releases its programmatic dependents, so
// the next kernel's CTAs become resident (and run their pre-wait loads) as this kernel's CTAs retire.
struct NoPre {
	__device__ __forceinline__ double operator()(uint32_t) const { return 0.0; }
};

template <bool BLOCKED, bool WANT_DIAG, bool IDX16, bool HALO, typename Gate, typename Late, typename Pre,
	  typename Body>
__device__ __forceinline__ void sell_stream_rows(const SellView A, const double *x,
						 const StreamConfig cfg, unsigned char *smem, Gate gate, Late late,
						 Pre pre, Body body)
{
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t total_warps = gridDim.x * kStreamWarps;
	// Slices are dealt round-robin over the warps of the grid.  CTA-major (the 10 warps of a CTA work on
	// 10 consecutive slices) or warp-major (the CTAs interleave): measured, not derived -- the per-entry
	// layout (9-point Laplacian) is 4.5-7.6 % faster warp-major (L4096 SpMV 287 -> 265 us), the blocked
	// layouts are equal (Q1) or 1.4-4 % slower (Q4, ragged triangles), so each keeps its better order.
	const uint32_t first = BLOCKED ? blockIdx.x * kStreamWarps + warp : warp * gridDim.x + blockIdx.x;
	// shared memory: [warps][stages][stage_bytes] | [warps][stages] mbarriers | [warps][stages] {off,width}
	unsigned char *ring = smem + (size_t)warp * cfg.stages * cfg.stage_bytes;
	uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)kStreamWarps * cfg.stages * cfg.stage_bytes) +
			 warp * kStreamMaxStages;
	uint2 *meta = reinterpret_cast<uint2 *>(smem + (size_t)kStreamWarps * cfg.stages * cfg.stage_bytes +
						 kStreamWarps * kStreamMaxStages * sizeof(uint64_t)) +
		      warp * kStreamMaxStages;
	const uint32_t val_bytes_per_col = kSliceRows * 8u;
	const uint32_t col_bytes_per_col = stream_id_bytes(BLOCKED, IDX16);

	uint64_t policy = 0;
	if (lane == 0) {
		for (uint32_t s = 0; s < cfg.stages; s++)
			mbar_init(bars + s, 1);
		fence_mbar_init();
		fence_proxy_async();
		policy = l2_evict_first_policy();
	}
	__syncwarp();

	// producer (lane 0): fill stage `st` with slice `s`
	auto issue = [&](uint32_t st, uint32_t v) {
		uint32_t s = v + A.visit_shift;
		if (s >= A.n_slices)
			s -= A.n_slices;
		uint32_t off, width;
		if (A.uniform_width) {
			off = s * A.uniform_width;
			width = A.uniform_width;
		} else {
			off = __ldg(A.slice_off + s);
			width = __ldg(A.slice_off + s + 1) - off;
		}
		meta[st] = make_uint2(off, width);
		unsigned char *dst = ring + (size_t)st * cfg.stage_bytes;
		const uint32_t vb = width * val_bytes_per_col, cb = width * col_bytes_per_col;
		mbar_arrive_expect_tx(bars + st, vb + cb);
		if (width) {
			bulk_g2s(dst, A.val + (size_t)off * kSliceRows, vb, bars + st, policy);
			const unsigned char *csrc = reinterpret_cast<const unsigned char *>(A.col) +
						    (size_t)off * col_bytes_per_col;
			bulk_g2s(dst + (size_t)cfg.cap * val_bytes_per_col, csrc, cb, bars + st, policy);
		}
	};
	if (lane == 0)
		for (uint32_t st = 0; st < cfg.stages; st++) {
			const uint64_t s = (uint64_t)first + (uint64_t)st * total_warps;
			if (s < A.n_slices)
				issue(st, (uint32_t)s);
		}
	__syncwarp();
	pdl_wait();
	const bool go = gate();
	pdl_launch_dependents();
	if (!go) {
		// a CTA must not exit with bulk copies still landing in its shared memory
		for (uint32_t st = 0; st < cfg.stages; st++)
			if ((uint64_t)first + (uint64_t)st * total_warps < A.n_slices)
				mbar_wait(bars + st, 0);
		return;
	}
	const uint32_t col_shift = HALO ? A.col_shift : 0u;

	uint32_t n = 0;
	bool late_done = false;
	// SELL-C-sigma: the row stored at this lane's position is fetched one slice ahead, so that the
	// load's latency never sits in front of the gathers
	auto slice_of = [&](uint64_t v) {
		uint32_t s = (uint32_t)v + A.visit_shift;
		return s >= A.n_slices ? s - A.n_slices : s;
	};
	uint32_t row_ahead = 0;
	if (A.perm && first < A.n_slices)
		row_ahead = __ldg(A.perm + (size_t)slice_of(first) * kSliceRows + lane);
	// ring position of visit n (n % stages) and its mbarrier phase ((n / stages) & 1), kept incrementally:
	// a division by the run-time ring depth per slice is ~25 instructions of a ~440-instruction slice
	uint32_t st = 0, parity = 0;
	for (uint64_t s64 = first; s64 < A.n_slices; s64 += total_warps, n++) {
		const uint32_t s = slice_of(s64);
		const uint32_t row = A.perm ? row_ahead : s * kSliceRows + lane;
		if (A.perm && s64 + total_warps < A.n_slices)
			row_ahead = __ldg(A.perm + (size_t)slice_of(s64 + total_warps) * kSliceRows + lane);
		if (!late_done && s64 >= A.late_from && s64 < A.late_to) {
			late_done = true;
			if (!late()) {
				// drain the stages that are in flight for visits n .. n + stages - 1
				for (uint32_t a = 0; a < cfg.stages; a++)
					if (s64 + (uint64_t)a * total_warps < A.n_slices)
						mbar_wait(bars + (n + a) % cfg.stages, ((n + a) / cfg.stages) & 1u);
				return;
			}
		}
		mbar_wait(bars + st, parity);
		// The work on one slice, compiled twice for HALO kernels: slices that read halo columns gather
		// coherently (see ld_gather_f64_coherent), all others through the read-only path.  (One body with
		// a run-time choice of the load instruction per gather cost 100-200 bytes of register spills.)
		auto slice_work = [&](auto coh_tag) {
		constexpr bool kCoh = decltype(coh_tag)::value;
		const double pre_value = pre(row);
		const uint32_t crow = row + col_shift;   // this row's index in the column space
		const uint2 m = meta[st];
		const uint32_t width = m.y;
		const unsigned char *stage = ring + (size_t)st * cfg.stage_bytes;
		const double *sval = reinterpret_cast<const double *>(stage) + lane;
		double acc = 0.0, diag = 0.0, x_row = 0.0;
		bool have_row = false;
		// every lane is done with the stage: hand it back to the producer
		auto release = [&]() {
			__syncwarp();
			if (lane == 0) {
				const uint64_t next = s64 + (uint64_t)cfg.stages * total_warps;
				if (next < A.n_slices) {
					fence_proxy_async();
					issue(st, (uint32_t)next);
				}
			}
		};
		// A slice that fits one gather batch (all 2-D linear FEM rows, the 9-point stencil) is moved
		// to registers as a whole while its gathers are in flight and the stage is re-armed BEFORE the
		// arithmetic: the next bulk copy into this stage then overlaps gather latency and math, which
		// keeps two copies per warp in flight instead of one (the kernel is bound by bytes in flight).
		// id k of this lane: entry column k, or block k of the lane's node pair when BLOCKED
		const unsigned char *sids = stage + (size_t)cfg.cap * val_bytes_per_col;
		const uint32_t id_base = BLOCKED ? (crow >> 1) : crow;
		auto load_id = [&](uint32_t k) -> uint32_t {
			constexpr uint32_t kStride = BLOCKED ? 16u : kSliceRows;
			const uint32_t mine = BLOCKED ? (lane >> 1) : lane;
			if (IDX16) {
				const int d = reinterpret_cast<const short *>(sids)[k * kStride + mine];
				return d == kPadDelta ? kPadCol : id_base + (uint32_t)d;
			}
			return reinterpret_cast<const uint32_t *>(sids)[k * kStride + mine];
		};
		auto gather1 = [&](const double *p) { return kCoh ? ld_gather_f64_coherent(p) : ld_gather_f64(p); };
		auto gather2 = [&](const double *p) { return kCoh ? ld_gather_f64x2_coherent(p) : ld_gather_f64x2(p); };
#ifndef NB_EARLY
#define NB_EARLY 2
#endif
		// NB_EARLY: 1 = both layouts, 2 = per-entry layout only (the blocked variant needs 122 registers)
		const bool single = (NB_EARLY == 1 || (NB_EARLY == 2 && !BLOCKED)) &&
				    (BLOCKED ? (width >> 1) : width) <= (uint32_t)kGatherBatch;
		if (!BLOCKED && single) {
			uint32_t cj[kGatherBatch];
			double xj[kGatherBatch], vj[kGatherBatch];
#pragma unroll
			for (int u = 0; u < kGatherBatch; u++)
				cj[u] = (uint32_t)u < width ? load_id(u) : kPadCol;
#pragma unroll
			for (int u = 0; u < kGatherBatch; u++)
				xj[u] = gather1(x + (cj[u] == kPadCol ? 0u : cj[u]));
#pragma unroll
			for (int u = 0; u < kGatherBatch; u++)
				vj[u] = (uint32_t)u < width ? sval[u * kSliceRows] : 0.0;
			release();
#pragma unroll
			for (int u = 0; u < kGatherBatch; u++) {
				const double t = __dmul_rn(vj[u], xj[u]);
				acc = (cj[u] == kPadCol) ? acc : __dadd_rn(acc, t);
				if (cj[u] == crow) {
					diag = vj[u];
					x_row = xj[u];
					have_row = true;
				}
			}
		} else if (BLOCKED && single) {
			const uint32_t nblk = width >> 1;
			const uint32_t my_node = crow >> 1;
			uint32_t cb[kGatherBatch];
			double2 xb[kGatherBatch];
			double v0[kGatherBatch], v1[kGatherBatch];
#pragma unroll
			for (int u = 0; u < kGatherBatch; u++)
				cb[u] = (uint32_t)u < nblk ? load_id(u) : kPadCol;
#pragma unroll
			for (int u = 0; u < kGatherBatch; u++)
				xb[u] = gather2(x + 2 * (size_t)(cb[u] == kPadCol ? 0u : cb[u]));
#pragma unroll
			for (int u = 0; u < kGatherBatch; u++) {
				v0[u] = (uint32_t)u < nblk ? sval[(2 * u) * kSliceRows] : 0.0;
				v1[u] = (uint32_t)u < nblk ? sval[(2 * u + 1) * kSliceRows] : 0.0;
			}
			release();
#pragma unroll
			for (int u = 0; u < kGatherBatch; u++) {
				const bool pad = cb[u] == kPadCol;
				const double t0 = __dmul_rn(v0[u], xb[u].x);
				acc = pad ? acc : __dadd_rn(acc, t0);
				const double t1 = __dmul_rn(v1[u], xb[u].y);
				acc = pad ? acc : __dadd_rn(acc, t1);
				if (cb[u] == my_node) {
					diag = (row & 1) ? v1[u] : v0[u];
					x_row = (row & 1) ? xb[u].y : xb[u].x;
					have_row = true;
				}
			}
		} else if (!BLOCKED) {
			uint32_t j0 = 0;
			// full batches: column ids, then ALL gathers, then (behind a warp barrier that
			// keeps the shared-memory value loads and hence the math from creeping up
			// between the gathers) the accumulation
			for (; j0 + kGatherBatch <= width; j0 += kGatherBatch) {
				uint32_t cj[kGatherBatch];
				double xj[kGatherBatch];
#pragma unroll
				for (int u = 0; u < kGatherBatch; u++)
					cj[u] = load_id(j0 + u);
#pragma unroll
				for (int u = 0; u < kGatherBatch; u++)
					xj[u] = gather1(x + (cj[u] == kPadCol ? 0u : cj[u]));
				__syncwarp();
#pragma unroll
				for (int u = 0; u < kGatherBatch; u++) {
					const double v = sval[(j0 + u) * kSliceRows];
					const double t = __dmul_rn(v, xj[u]);
					acc = (cj[u] == kPadCol) ? acc : __dadd_rn(acc, t);
					if (cj[u] == crow) {
						diag = v;
						x_row = xj[u];
						have_row = true;
					}
				}
			}
			for (; j0 < width; j0++) {
				const uint32_t cj = load_id(j0);
				const double v = sval[j0 * kSliceRows];
				const double xj = gather1(x + (cj == kPadCol ? 0u : cj));
				const double t = __dmul_rn(v, xj);
				acc = (cj == kPadCol) ? acc : __dadd_rn(acc, t);
				if (cj == crow) {
					diag = v;
					x_row = xj;
					have_row = true;
				}
			}
		} else {
			// one node id per 2x2 block; lanes 2k, 2k+1 (the two dofs of a node) share it.
			// The work per block is ~45 instructions and the kernel's time follows its instruction count
			// closely (ncu: 440 per slice, issue slots half used by 20 warps per SM), so two things are kept
			// out of the common case:
			//  * padding: entries are ascending, so a row is padded iff its LAST block id is the marker; a
			//    warp vote picks the loop without any padding selects (uniform meshes: only slices with
			//    boundary rows have padding, 0.3 % at Q1);
			//  * x[row]: picked out of the diagonal block only when the diagonal VALUE is wanted as well
			//    (init kernel); otherwise it is one more coalesced load issued with the gathers.
			const uint32_t nblk = width >> 1;
			const uint32_t my_node = crow >> 1;
			const bool padded = nblk != 0 && load_id(nblk - 1) == kPadCol;
			const bool any_pad = __any_sync(0xffffffffu, padded);
			if (!WANT_DIAG && row < A.N) {
				x_row = __ldg(x + crow);   // (a plain load: dropped by the compiler where the body ignores x[row])
				have_row = true;
			}
			auto blocks = [&](auto pads_tag) {
				constexpr bool kPads = decltype(pads_tag)::value;
				uint32_t b0 = 0;
				for (; b0 + kGatherBatch <= nblk; b0 += kGatherBatch) {
					uint32_t cb[kGatherBatch];
					double2 xb[kGatherBatch];
#pragma unroll
					for (int u = 0; u < kGatherBatch; u++)
						cb[u] = load_id(b0 + u);
#pragma unroll
					for (int u = 0; u < kGatherBatch; u++)
						xb[u] = gather2(x + (uint32_t)(2u * ((kPads && cb[u] == kPadCol) ? 0u : cb[u])));
					__syncwarp();
#pragma unroll
					for (int u = 0; u < kGatherBatch; u++) {
						const double v0 = sval[(2 * (b0 + u)) * kSliceRows];
						const double v1 = sval[(2 * (b0 + u) + 1) * kSliceRows];
						const bool pad = kPads && cb[u] == kPadCol;
						const double t0 = __dmul_rn(v0, xb[u].x);
						acc = pad ? acc : __dadd_rn(acc, t0);
						const double t1 = __dmul_rn(v1, xb[u].y);
						acc = pad ? acc : __dadd_rn(acc, t1);
						if (WANT_DIAG && cb[u] == my_node) {
							diag = (row & 1) ? v1 : v0;
							x_row = (row & 1) ? xb[u].y : xb[u].x;
							have_row = true;
						}
					}
				}
				for (; b0 < nblk; b0++) {
					const uint32_t cb = load_id(b0);
					const double v0 = sval[(2 * b0) * kSliceRows], v1 = sval[(2 * b0 + 1) * kSliceRows];
					const bool pad = kPads && cb == kPadCol;
					const double2 xb = gather2(x + (uint32_t)(2u * (pad ? 0u : cb)));
					const double t0 = __dmul_rn(v0, xb.x);
					acc = pad ? acc : __dadd_rn(acc, t0);
					const double t1 = __dmul_rn(v1, xb.y);
					acc = pad ? acc : __dadd_rn(acc, t1);
					if (WANT_DIAG && cb == my_node) {
						diag = (row & 1) ? v1 : v0;
						x_row = (row & 1) ? xb.y : xb.x;
						have_row = true;
					}
				}
			};
			if (any_pad)
				blocks(std::true_type{});
			else
				blocks(std::false_type{});
		}
		if (!single)
			release();
		if (!have_row && row < A.N)
			x_row = gather1(x + crow);   // row without a stored diagonal
		body(row, acc, diag, x_row, pre_value);
		};   // slice_work
		if (HALO && s64 >= A.late_from && s64 < A.late_to)
			slice_work(std::true_type{});
		else
			slice_work(std::false_type{});
		if (++st == cfg.stages) {
			st = 0;
			parity ^= 1u;
		}
	}
}

// host: WHERE in the visit order the halo-reading slices of a rank-local block go.  The plan puts them last
// (visit indices [late_from, n)).  With slices dealt round-robin over W warps, the warps without a slice in the
// final partial round have one slice of slack: a halo slice costs ~3 us more than an interior one (flag wait
// with a system-scope acquire, gathers served from L2 instead of L1), so the halo slices are moved to the END
// OF THE LAST FULL ROUND, onto exactly those warps -- late enough for the halo to have arrived (tens of
// microseconds after the push at the kernel's start), and off the kernel's critical path.
inline void place_halo_slices(SellView *V, uint32_t total_warps)
{
	const uint32_t n = V->n_slices;
	if (V->late_from >= n || total_warps == 0)
		return;   // no halo
	const uint32_t n_h = n - V->late_from, rem = n % total_warps;
	if (rem == 0 || n < total_warps || n_h > total_warps - rem)
		return;   // no slack to hide them in: leave them last
	const uint32_t v_end = (n / total_warps) * total_warps;
	// last halo slice = visit_shift - 1 (circularly); it gets visit index v_end - 1
	V->visit_shift = (uint32_t)(((uint64_t)V->visit_shift + n - v_end) % n);
	V->late_from = v_end - n_h;
	V->late_to = v_end;
}

// host: choose ring depth / CTAs per SM for a matrix; returns false when the
// widest slice does not fit (the register-path kernels are used instead)
bool stream_config(const nbgpu_matrix_s *A, const void *kernel, StreamConfig *cfg, int block = kBlock);

}  // namespace nbgpu
