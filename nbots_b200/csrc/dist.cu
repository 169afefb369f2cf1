// dist.cu -- row-partitioned Jacobi-PCG / CG / SpMV over several B200s of one node (host side).
//
// One rank per GPU: one process per GPU with CUDA-IPC windows, or one host thread
// per GPU inside a single process with peer access (nbgpu_dist_connect_local).  Rank r owns the
// contiguous rows [row_starts[r], row_starts[r+1]) of the global matrix and the matching slices of
// every vector.  Its block of the matrix is a rank-local SELL matrix whose column space is
//
//       lower halo | pad | owned columns | pad | upper halo          (nbgpu_dist_ext_layout)
//
// i.e. the global order with the remote ranges squeezed out: halo = the remote rows its rows
// reference, ascending global id.  Local ids ascend with global ids, so every row sum is still the
// reference's (sparse.c:405-414), the block keeps the single-GPU storage forms (2x2-blocked and
// 16-bit column ids: a neighbour's boundary line is as close in local numbering as in global), and
// the halo parts start on their own 128-byte lines (owned entries and peer-written entries never
// share a cache line).
//
// The kernels are the single-GPU ones (krylov_kernels.cuh) instantiated with the PeerComm policy
// (dist_comm.cuh): halo push at the start of K1, late halo wait, all-reduce of the dot-product
// partials by the CTA that finishes the grid reduction.  No collective library call sits in the
// iteration.  Per iteration: CLASSIC 1 halo + 2 scalar exchanges, FUSED 1 halo + 1 scalar exchange.
//
// Why single-buffered halo parts are race-free: a rank can only run ahead of a peer by less than
// one reduction.  The kernel after K1(k) needs every rank's K1(k) partials, which a rank posts only
// after ALL its CTAs finished K1(k) -- i.e. finished reading the halo -- and the next halo push
// happens at the start of K1(k+1).  Reduction messages alternate between two slots (see
// dist_comm.cuh) and are numbered continuously across solves.
#include "krylov_kernels.cuh"
#include "dist_plan.cuh"

using namespace nbgpu;

namespace {

__global__ void __launch_bounds__(32)
halo_push_kernel(const PeerTable *T, const uint32_t *__restrict__ send_idx, const double *__restrict__ v_own,
		 int which, unsigned long long seq, unsigned int *ticket)
{
	pdl_wait();
	pdl_launch_dependents();
	if (*(volatile int *)&T->ctrl[T->rank]->error)
		return;
	push_halo_piece(*T, send_idx, v_own, which, seq, ticket, blockIdx.x, gridDim.x);
}

// ---- distributed SpMV (config 3): y = A x with the halo of x exchanged first --------
template <int LAYOUT>
__global__ void __launch_bounds__(kBlock, kStreamCtas)
dist_plain_spmv_kernel(SellView A, StreamConfig cfg, const PeerTable *T, unsigned long long seq,
		       const double *x_ext, double *__restrict__ y, unsigned int *ticket)
{
	extern __shared__ __align__(128) unsigned char smem[];
	sell_stream_rows<(LAYOUT & 1) != 0, false, (LAYOUT & 2) != 0, true>(
		A, x_ext, cfg, smem, [] { return true; }, [&] { return warp_wait_halo(*T, 1, seq); }, NoPre(),
		[&](uint32_t row, double acc, double, double, double) {
			if (row < A.N)
				y[row] = acc;
		});
	// last CTA out tells the sources that the halo part may be overwritten
	__shared__ bool last;
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
	}
	__syncthreads();
	if (last && (int)threadIdx.x < T->world)
		st_release_sys(&T->ctrl[threadIdx.x]->xh_ack[T->rank], seq);
}

uint32_t round_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

}  // namespace

namespace nbgpu {

// SpMV visit order of a rank-local block: the largest (circular) run of slices that read no halo
// column is the interior; the visit starts there, so every halo-reading slice comes last
void plan_visit_order(nbgpu_dist_plan_s *P, const uint32_t *rows_size)
{
	const uint32_t off_own = P->off_own, off_up = P->off_up;
	const uint32_t n_slices = (P->N_loc + kSliceRows - 1) / kSliceRows;
	std::vector<uint8_t> reads_halo(n_slices, 0);
	uint64_t k = 0;
	for (uint32_t i = 0; i < P->N_loc; i++)
		for (uint32_t j = 0; j < rows_size[i]; j++, k++)
			if (P->cols_local[k] < off_own || P->cols_local[k] >= off_up)
				reads_halo[i / kSliceRows] = 1;
	uint32_t best_start = 0, best_len = 0;
	bool any = false;
	for (uint32_t s0 = 0; s0 < n_slices; s0++) {
		if (!reads_halo[s0])
			continue;
		// run of clean slices that starts right after halo slice s0 (circularly)
		any = true;
		uint32_t len = 0;
		while (len < n_slices && !reads_halo[(s0 + 1 + len) % n_slices])
			len++;
		if (len >= best_len) {
			best_len = len;
			best_start = (s0 + 1) % n_slices;
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

}  // namespace nbgpu

// ------------------------------------------------------------------ host objects --

extern "C" {

/* column space of a rank-local block: lower halo | owned | upper halo, each part on its own 128-byte lines */
int nbgpu_dist_ext_layout(uint32_t n_lo, uint32_t N_loc, uint32_t n_hi, uint32_t *off_own, uint32_t *off_up,
			  uint32_t *ext_len)
{
	const uint64_t o = round_up(n_lo, kExtAlign);
	const uint64_t u = (o + N_loc + kExtAlign - 1) / kExtAlign * kExtAlign;
	const uint64_t e = (u + n_hi + kExtAlign - 1) / kExtAlign * kExtAlign;
	if (e > 0xFFFFFFF0ull) {
		set_error("rank-local column space too large");
		return NBGPU_ERR_ARG;
	}
	if (off_own)
		*off_own = (uint32_t)o;
	if (off_up)
		*off_up = (uint32_t)u;
	if (ext_len)
		*ext_len = (uint32_t)std::max<uint64_t>(e, kExtAlign);
	return NBGPU_OK;
}

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
	P->n_lo = (uint32_t)(std::lower_bound(halo.begin(), halo.end(), r0) - halo.begin());
	P->n_hi = P->n_halo - P->n_lo;
	if (nbgpu_dist_ext_layout(P->n_lo, P->N_loc, P->n_hi, &P->off_own, &P->off_up, &P->ext_len) != NBGPU_OK) {
		delete P;
		return NBGPU_ERR_ARG;
	}
	// local numbering: the global order with the remote ranges squeezed out; entry order untouched
	P->cols_local.resize(nnz);
	const uint32_t n_lo = P->n_lo, off_own = P->off_own, off_up = P->off_up;
#pragma omp parallel for schedule(static)
	for (int64_t k = 0; k < (int64_t)nnz; k++) {
		const uint32_t c = cols_global[k];
		if (c >= r0 && c < r1) {
			P->cols_local[k] = off_own + (c - r0);
		} else {
			const uint32_t h = (uint32_t)(std::lower_bound(halo.begin(), halo.end(), c) - halo.begin());
			P->cols_local[k] = h < n_lo ? h : off_up + (h - n_lo);
		}
	}
	plan_visit_order(P, rows_size);
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

int nbgpu_dist_plan_layout(const nbgpu_dist_plan_t *P, uint32_t *n_lo, uint32_t *off_own, uint32_t *off_up,
			   uint32_t *ext_len)
{
	NB_ARG(P != nullptr);
	if (n_lo)
		*n_lo = P->n_lo;
	if (off_own)
		*off_own = P->off_own;
	if (off_up)
		*off_up = P->off_up;
	if (ext_len)
		*ext_len = P->ext_len;
	return NBGPU_OK;
}

int nbgpu_dist_plan_visit_order(const nbgpu_dist_plan_t *P, uint32_t total_warps, uint32_t *visit_shift,
				uint32_t *late_from, uint32_t *late_to)
{
	NB_ARG(P != nullptr && visit_shift != nullptr && late_from != nullptr && late_to != nullptr);
	SellView V;
	V.n_slices = (P->N_loc + kSliceRows - 1) / kSliceRows;
	V.visit_shift = P->visit_shift;
	V.late_from = P->late_from;
	if (total_warps)
		place_halo_slices(&V, total_warps);
	*visit_shift = V.visit_shift;
	*late_from = V.late_from;
	*late_to = std::min(V.late_to, V.n_slices);
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
	NB_ARG(out != nullptr && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world);
	nbgpu_dist_t *D = new nbgpu_dist_t();
	D->rank = rank;
	D->world = world;
	D->ext_len = round_up((uint32_t)std::max<size_t>(ext_len, 2), kExtAlign);
	D->window_bytes = 4096 + 2 * D->ext_len * sizeof(double);
	cudaError_t e = cudaMalloc(&D->window, D->window_bytes);
	if (e == cudaSuccess)
		e = cudaMemset(D->window, 0, D->window_bytes);
	if (e == cudaSuccess)
		e = cudaMalloc(&D->d_state, sizeof(KrylovState));
	if (e == cudaSuccess)
		e = cudaMalloc(&D->d_ticket, 4 * sizeof(unsigned int));
	if (e == cudaSuccess)
		e = cudaMemset(D->d_ticket, 0, 4 * sizeof(unsigned int));
	if (e == cudaSuccess)
		e = cudaMalloc(&D->d_table, sizeof(PeerTable));
	if (e == cudaSuccess)
		e = cudaMallocHost(&D->h_state, 4 * sizeof(KrylovState) + 64);
	if (e == cudaSuccess)
		e = cudaEventCreateWithFlags(&D->poll_ev[0], cudaEventDisableTiming);
	if (e == cudaSuccess)
		e = cudaEventCreateWithFlags(&D->poll_ev[1], cudaEventDisableTiming);
	cudaIpcMemHandle_t h;
	memset(&h, 0, sizeof(h));
	if (e == cudaSuccess && ipc_handle_out)
		e = cudaIpcGetMemHandle(&h, D->window);
	if (e != cudaSuccess) {
		set_error("nbgpu_dist_create: %s", cudaGetErrorString(e));
		cudaGetLastError();
		delete D;
		return NBGPU_ERR_COMM;
	}
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	if (ipc_handle_out)
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
		D->ext_len_of[r] = round_up((uint32_t)std::max<uint64_t>(all_ext_len[r], 2), kExtAlign);
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
		D->peer_is_ipc[r] = true;
	}
	D->connected = true;
	return NBGPU_OK;
}

/* Ranks that live in ONE process (one host thread per GPU): the peers' windows are plain device
 * pointers, reachable once peer access is enabled.  all[r] = rank r's object, device_of[r] its GPU.
 * Called by every rank's own thread (bound to its device). */
int nbgpu_dist_connect_local(nbgpu_dist_t *D, nbgpu_dist_t *const *all, const int *device_of)
{
	NB_INIT();
	NB_ARG(D != nullptr && all != nullptr && device_of != nullptr);
	for (int r = 0; r < D->world; r++) {
		NB_ARG(all[r] != nullptr);
		D->ext_len_of[r] = all[r]->ext_len;
		D->peer_window[r] = all[r]->window;
		if (r == D->rank || device_of[r] == ctx().device)
			continue;
		int can = 0;
		cudaDeviceCanAccessPeer(&can, ctx().device, device_of[r]);
		cudaError_t e = can ? cudaDeviceEnablePeerAccess(device_of[r], 0) : cudaErrorPeerAccessUnsupported;
		if (e == cudaErrorPeerAccessAlreadyEnabled) {
			cudaGetLastError();
			e = cudaSuccess;
		}
		if (e != cudaSuccess) {
			set_error("peer access %d -> %d: %s", ctx().device, device_of[r], cudaGetErrorString(e));
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
			if (r != D->rank && D->peer_window[r] && D->peer_is_ipc[r])
				cudaIpcCloseMemHandle(D->peer_window[r]);
		cudaFree(D->window);
		cudaFree(D->d_state);
		cudaFree(D->d_ticket);
		cudaFree(D->d_table);
		cudaFreeHost(D->h_state);
		cudaEventDestroy(D->poll_ev[0]);
		cudaEventDestroy(D->poll_ev[1]);
	}
	delete D;
	return NBGPU_OK;
}

}  // extern "C"

namespace {

// fills D->d_table for plan P (kept until another plan is used with D)
int build_peer_table(nbgpu_dist_t *D, nbgpu_dist_plan_t *P)
{
	PeerTable table, *T = &table;
	NB_ARG(D->connected && P->have_sends && P->rank == D->rank && P->world == D->world);
	NB_ARG((size_t)P->ext_len <= D->ext_len);
	if (D->broken) {
		set_error("an earlier exchange on this nbgpu_dist_t failed; destroy and re-create it on every rank");
		return NBGPU_ERR_COMM;
	}
	memset(T, 0, sizeof(*T));
	T->world = D->world;
	T->rank = D->rank;
	const char *to = getenv("NBGPU_DIST_TIMEOUT_MS");
	T->timeout_ns = (unsigned long long)(to ? atoll(to) : 10000) * 1000000ull;
	for (int r = 0; r < D->world; r++) {
		T->ctrl[r] = (DistControl *)D->peer_window[r];
		T->v_halo_dst[r] = D->peer_v_ext(r) + P->dst_offset[r];
		T->x_halo_dst[r] = D->peer_x_ext(r) + P->dst_offset[r];
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
	if (D->table_plan != P) {
		NB_CUDA(cudaStreamSynchronize(ctx().stream));
		NB_CUDA(cudaMemcpy(D->d_table, T, sizeof(PeerTable), cudaMemcpyHostToDevice));
		D->table_plan = P;
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

void local_view(const nbgpu_dist_plan_t *P, const nbgpu_matrix_t *A, SellView *V)
{
	V->N = A->N; V->n_slices = A->n_slices; V->slice_off = A->d_slice_off; V->val = A->d_val;
	V->col = A->stream_ids();
	V->uniform_width = A->uniform_width;
	V->visit_shift = P->visit_shift; V->late_from = P->late_from;
	V->col_shift = P->off_own;
}

int dist_solve(nbgpu_dist_t *D, nbgpu_dist_plan_t *P, const nbgpu_matrix_t *A, const double *d_b, double *d_x,
	       uint32_t max_iter, double tol, uint32_t *niter, double *tol_reached, bool jacobi)
{
	NB_INIT();
	NB_ARG(D != nullptr && P != nullptr && A != nullptr && d_b != nullptr && d_x != nullptr);
	NB_ARG(A->N == P->N_loc && A->n_cols == P->ext_len && A->col_shift == P->off_own);
	Context &c = ctx();
	PeerComm comm;
	NB_TRY(build_peer_table(D, P));
	comm.T = D->d_table;
	comm.mine = D->ctrl();
	comm.send_idx = P->d_send_idx;
	comm.world = D->world;
	comm.rank = D->rank;
	comm.total_sends = P->send_ptr[P->world];
	for (int r = 0; r < kMaxRanks; r++)
		comm.ctrl[r] = r < D->world ? (DistControl *)D->peer_window[r] : nullptr;
	{
		const char *to = getenv("NBGPU_DIST_TIMEOUT_MS");
		comm.timeout_ns = (unsigned long long)(to ? atoll(to) : 10000) * 1000000ull;
	}
	const uint32_t N = A->N;
	const size_t Np = ((size_t)N + 1) & ~(size_t)1;
	const bool fused = krylov_want_fused(A, false);
	{
		// Who posts a reduction: the consumer kernel (CLASSIC: 46.6 against 48.4 us per iteration on two GPUs)
		// or the producer's last CTA (FUSED: 46.1 against 47.8).  NBGPU_DIST_REDUCE=ticket|consumer overrides;
		// every rank must make the same choice.
		const char *rd = getenv("NBGPU_DIST_REDUCE");
		comm.cpost = rd ? (rd[0] == 'c' ? 1 : 0) : (fused ? 0 : 1);
	}
	// vectors: the one the SpMV gathers lives in the window (it has halo parts), the rest in the workspace
	const int n_ws = fused ? (jacobi ? 6 : 4) : (jacobi ? 5 : 3);
	NB_TRY(ensure_workspace((size_t)n_ws * Np * sizeof(double)));
	double *v_own = D->v_ext() + P->off_own;
	double *x_own = D->x_ext() + P->off_own;
	KrylovRun R;
	double *next = c.ws;
	auto take = [&]() { double *v = next; next += Np; return v; };
	R.x = take();
	if (fused) {
		if (jacobi) {
			R.q = v_own;
			R.g = take();
			R.diag = take();
		} else {
			R.g = R.q = v_own;
		}
		R.p = take();
		R.w = take();
		R.s = take();
	} else {
		R.p = v_own;
		R.g = take();
		R.w = take();
		if (jacobi) {
			R.q = take();
			R.diag = take();
		} else {
			R.q = R.g;
		}
	}
	// No persisting-L2 window here: with one vector in the window (outside the work-vector block) a window
	// over the others measured SLOWER than the default policy (51.2 vs 49.6 us/iteration, 1 M dof).
	R.st = D->d_state;
	R.hst = D->h_state;
	R.h_err = (int *)(D->h_state + 4);   // pinned, behind the four state slots
	R.d_err = &D->ctrl()->error;
	R.poll_ev[0] = D->poll_ev[0];
	R.poll_ev[1] = D->poll_ev[1];
	NB_CUDA(cudaMemcpyAsync(R.x, d_x, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
	NB_CUDA(cudaMemcpyAsync(x_own, d_x, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
	R.A = A;
	local_view(P, A, &R.V);
	R.jacobi = jacobi;
	R.fused = fused;
	R.pdl = !getenv("NBGPU_NO_PDL");
	R.max_iter = max_iter;
	R.tol = tol;
	R.b = d_b;
	R.v_ext = D->v_ext();
	R.x_ext = D->x_ext();
	R.partials = c.partials;
	// the x-halo exchange shares its flags with nbgpu_dist_spmv: one counter for both
	R.seq_in = ++D->spmv_seq;
	R.msg_seq = D->msg_seq;
	R.halo_seq = D->halo_seq;

	// x halo -> init
	const int push_grid = std::max(1, (int)std::min<uint32_t>((P->send_ptr[P->world] + 31) / 32, 1024));
	if (n_destinations(P))
		NB_CUDA(launch_on(false, halo_push_kernel, push_grid, 32, 0, (const PeerTable *)D->d_table,
				  (const uint32_t *)P->d_send_idx, (const double *)x_own, 1, R.seq_in, D->d_ticket + 1));
	const int status = krylov_run(R, comm, niter, tol_reached);
	if (status != NBGPU_OK && status != NBGPU_NOT_CONVERGED) {
		if (status == NBGPU_ERR_COMM)
			D->broken = true;
		return status;
	}
	// all ranks saw the same k_final: the sequence spaces advance identically everywhere
	D->msg_seq = R.msg_seq;
	D->halo_seq = R.halo_seq;
	NB_CUDA(cudaMemcpyAsync(d_x, R.x, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
	NB_CUDA(cudaStreamSynchronize(c.stream));
	return status;
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
	NB_ARG(A->N == P->N_loc && A->n_cols == P->ext_len && A->col_shift == P->off_own);
	Context &c = ctx();
	NB_TRY(build_peer_table(D, P));
	const PeerTable *T = D->d_table;
	SellView V;
	local_view(P, A, &V);
	StreamConfig cfg;
	const int layout = A->layout();
	const void *kern = by_layout(layout, [](auto L) {
		return (const void *)dist_plain_spmv_kernel<decltype(L)::value>;
	});
	if (!stream_config(A, kern, &cfg)) {
		set_error("distributed SpMV needs the streamed path");
		return NBGPU_ERR_ARG;
	}
	const unsigned long long seq = ++D->spmv_seq;
	double *x_own = D->x_ext() + P->off_own;
	if (d_in != x_own)   // callers that fill nbgpu_dist_input_vector() directly skip this copy
		NB_CUDA(cudaMemcpyAsync(x_own, d_in, (size_t)A->N * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
	const int push_grid = std::max(1, (int)std::min<uint32_t>((P->send_ptr[P->world] + 31) / 32, 1024));
	if (n_destinations(P))
		NB_CUDA(launch_on(false, halo_push_kernel, push_grid, 32, 0, T, (const uint32_t *)P->d_send_idx,
				  (const double *)x_own, 1, seq, D->d_ticket + 1));
	NB_CUDA(by_layout(layout, [&](auto L) {
		return launch_on(false, dist_plain_spmv_kernel<decltype(L)::value>, cfg.grid, kBlock, cfg.smem_bytes, V, cfg, T,
				 seq, (const double *)D->x_ext(), d_out, D->d_ticket);
	}));
	return NBGPU_OK;
}

/* the window's input vector (owned part): SpMV callers may write x here and pass it as d_in */
double *nbgpu_dist_input_vector(nbgpu_dist_t *D, const nbgpu_dist_plan_t *P)
{
	return (D && P) ? D->x_ext() + P->off_own : nullptr;
}

#ifdef NB_TIMELINE
/* diagnostic builds only: the %globaltimer stamps of the last row-partitioned solve, [512][10] */
int nbgpu_dist_timeline(unsigned long long *out)
{
	NB_INIT();
	NB_CUDA(cudaStreamSynchronize(ctx().stream));
	NB_CUDA(cudaMemcpyFromSymbol(out, g_timeline, sizeof(unsigned long long) * kTlIters * kTlSlots));
	return NBGPU_OK;
}
#endif

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
