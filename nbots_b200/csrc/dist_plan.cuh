// dist_plan.cuh -- host objects of the row-partitioned path (dist.cu, dist_fem.cu).
#pragma once

#include <vector>

#include "dist_comm.cuh"

namespace nbgpu {
struct KrylovState;
}

struct nbgpu_dist_plan_s {
	int rank = 0, world = 1;
	std::vector<uint32_t> row_starts;      // [world + 1]
	uint32_t N_loc = 0, n_halo = 0;
	uint32_t n_lo = 0, n_hi = 0;           // halo columns below / above the owned range
	uint32_t off_own = 0, off_up = 0, ext_len = 0;   // column space: [0,n_lo) | [off_own, +N_loc) | [off_up, +n_hi)
	uint64_t nnz = 0;
	std::vector<uint32_t> halo_global;     // [n_halo] ascending
	std::vector<uint32_t> recv_counts;     // [world]
	std::vector<uint32_t> cols_local;      // [nnz]
	// sends (filled by nbgpu_dist_plan_set_sends)
	bool have_sends = false;
	std::vector<uint32_t> send_ptr;        // [world + 1]
	std::vector<uint32_t> send_local;      // local row ids, grouped by destination
	std::vector<uint32_t> dst_offset;      // [world] where my block starts in the destination's ext vector
	uint32_t *d_send_idx = nullptr;
	// SpMV visit order: start right after the last slice that reads halo columns, so that every
	// halo-reading slice is visited at the end (visit index >= late_from) -- the halo wait of a
	// kernel is then hidden behind the interior slices
	uint32_t visit_shift = 0, late_from = 0;
};

struct nbgpu_dist_s {
	int rank = 0, world = 1;
	size_t ext_len = 0;                    // capacity of the two ext vectors
	void *window = nullptr;                // control block | v_ext | x_ext
	size_t window_bytes = 0;
	void *peer_window[nbgpu::kMaxRanks] = {};
	bool peer_is_ipc[nbgpu::kMaxRanks] = {};
	bool connected = false;
	bool broken = false;                   // an exchange failed: sequence numbers no longer agree
	unsigned long long msg_seq = 0;        // reduction messages, numbered continuously (identical on all ranks)
	unsigned long long halo_seq = 0;       // Krylov-vector halo pushes
	unsigned long long spmv_seq = 0;       // input-vector halo pushes
	nbgpu::KrylovState *d_state = nullptr;
	nbgpu::KrylovState *h_state = nullptr;        // pinned, 4 slots + 3 error words
	unsigned int *d_ticket = nullptr;
	nbgpu::PeerTable *d_table = nullptr;          // the peer table of the plan last used, in device memory
	const nbgpu_dist_plan_s *table_plan = nullptr;
	cudaEvent_t poll_ev[2] = {nullptr, nullptr};
	nbgpu::DistControl *ctrl() const { return (nbgpu::DistControl *)window; }
	double *v_ext() const { return (double *)((char *)window + 4096); }
	double *x_ext() const { return v_ext() + ext_len; }
	double *peer_v_ext(int r) const { return (double *)((char *)peer_window[r] + 4096); }
	double *peer_x_ext(int r) const { return peer_v_ext(r) + ext_len_of[r]; }
	size_t ext_len_of[nbgpu::kMaxRanks] = {};
};


namespace nbgpu {
void plan_visit_order(nbgpu_dist_plan_s *P, const uint32_t *rows_size);
}
