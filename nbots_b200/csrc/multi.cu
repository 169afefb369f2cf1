// multi.cu -- several GPUs driven from ONE host process: the caller's single thread hands a solve (or
// the whole static-elasticity pipeline) to one worker thread per GPU; every worker binds its own
// device context (nbgpu_thread_bind_device) and runs the unchanged one-rank-per-GPU code of dist.cu /
// dist_fem.cu, with peer access between the windows instead of CUDA IPC.
//
// This is how the reference-named entry points (shim/nb_shim.c, NBGPU_DEVICES=N) use N GPUs from a
// program that knows nothing about ranks: nb_sparse_solve_CG_precond_Jacobi / _conjugate_gradient
// (cg_precond_jacobi.h:8-15, conjugate_gradient.h:8-15) and nb_fem_compute_2D_Solid_Mechanics
// (static_elasticity2D.c:31-97) keep their single-process, blocking semantics.
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>

#include "dist_plan.cuh"
#include "matrix.cuh"

using namespace nbgpu;

namespace {

class Barrier {
public:
	explicit Barrier(int n) : n_(n) {}
	void wait()
	{
		std::unique_lock<std::mutex> lk(m_);
		const unsigned gen = gen_;
		if (++count_ == n_) {
			count_ = 0;
			gen_++;
			cv_.notify_all();
		} else {
			cv_.wait(lk, [&] { return gen != gen_; });
		}
	}

private:
	std::mutex m_;
	std::condition_variable cv_;
	int n_, count_ = 0;
	unsigned gen_ = 0;
};

// first failure of any worker: status + message, shown to the caller's thread afterwards
struct Failure {
	std::mutex m;
	int status = NBGPU_OK;
	std::string text;
	void note(int st)
	{
		if (st == NBGPU_OK)
			return;
		std::lock_guard<std::mutex> lk(m);
		if (status == NBGPU_OK) {
			status = st;
			text = nbgpu_last_error();
		}
	}
	bool any()
	{
		std::lock_guard<std::mutex> lk(m);
		return status != NBGPU_OK;
	}
};

int usable_devices(int wanted)
{
	const int have = nbgpu_device_count();
	return std::max(1, std::min({wanted, have, (int)kMaxRanks}));
}

}  // namespace

extern "C" {

/* NBGPU_DEVICES=N: how many GPUs the reference-named entry points may use (default 1) */
int nbgpu_devices_from_env(void)
{
	const char *env = getenv("NBGPU_DEVICES");
	const int want = env ? atoi(env) : 1;
	return want > 1 ? usable_devices(want) : 1;
}

int nbgpu_fem_static_elasticity2d_lists_multi(int n_devices, const nbgpu_mesh_desc_t *md,
					      const nbgpu_elem_tables_t *tables, const double D[4], double density,
					      uint32_t n_neu, const uint32_t *neu_dof, const double *neu_add,
					      uint32_t n_dir, const uint32_t *dir_dof, const double *dir_val,
					      int self_weight, const double gravity[2], double thickness,
					      const uint8_t *enabled, double solver_tol, double *displacement,
					      double *strain, nbgpu_fem_report_t *report)
{
	NB_ARG(md != nullptr && displacement != nullptr && D != nullptr);
	const int world = usable_devices(n_devices);
	std::vector<uint32_t> node_starts(world + 1);
	NB_TRY(nbgpu_partition_nodes(md->N_nod, world, 8, node_starts.data()));
	for (int r = 0; r < world; r++)
		if (node_starts[r] == node_starts[r + 1]) {
			set_error("mesh too small for %d devices", world);
			return NBGPU_ERR_ARG;
		}
	std::vector<nbgpu_dist_fem_t *> fem(world, nullptr);
	std::vector<int> device_of(world);
	for (int r = 0; r < world; r++)
		device_of[r] = r;
	Barrier bar(world);
	Failure fail;
	std::vector<int> asm_status(world, 0);
	std::vector<uint32_t> iters(world, 0);
	std::vector<double> resid(world, 0.0);
	std::vector<int> solve_status(world, 0);
	std::vector<double> ms_phase(4 * (size_t)world, 0.0);   // setup, assembly, solve, download per rank
	auto now = [] {
		return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
	};
	auto worker = [&](int r) {
		int st = nbgpu_thread_bind_device(device_of[r]);
		double t0 = now();
		if (st == NBGPU_OK)
			st = nbgpu_dist_fem_create(md, r, world, node_starts.data(), tables, D, density, n_neu, neu_dof, neu_add,
						   n_dir, dir_dof, dir_val, self_weight, gravity, thickness, nullptr, &fem[r]);
		fail.note(st);
		bar.wait();
		if (!fail.any()) {
			st = nbgpu_dist_fem_connect_local(fem[r], fem.data(), device_of.data());
			fail.note(st);
		}
		ms_phase[4 * r] = now() - t0;
		bar.wait();
		if (!fail.any()) {
			t0 = now();
			st = nbgpu_dist_fem_assemble(fem[r], enabled, nullptr, nullptr);
			if (st == NBGPU_DISTORTED_ELEMENT) {
				asm_status[r] = 1;   // static_elasticity2D.c:62-65
				st = NBGPU_OK;
			}
			if (st == NBGPU_OK)
				st = nbgpu_sync();
			fail.note(st);
			ms_phase[4 * r + 1] = now() - t0;
		}
		bar.wait();
		bool distorted = false;
		for (int q = 0; q < world; q++)
			distorted |= asm_status[q] != 0;
		if (!fail.any() && !distorted) {
			t0 = now();
			st = nbgpu_dist_fem_solve(fem[r], 0, 0, solver_tol, &iters[r], &resid[r]);
			if (st == NBGPU_OK || st == NBGPU_NOT_CONVERGED) {
				solve_status[r] = st;   // both accepted, static_elasticity2D.c:92
				st = NBGPU_OK;
			}
			fail.note(st);
			ms_phase[4 * r + 2] = now() - t0;
			if (st == NBGPU_OK) {
				t0 = now();
				st = nbgpu_dist_fem_results(fem[r], displacement + 2 * (size_t)node_starts[r]);
				fail.note(st);
				ms_phase[4 * r + 3] = now() - t0;
			}
		}
		bar.wait();   // nobody tears its window down while a peer may still store into it
		nbgpu_dist_fem_destroy(fem[r]);
		nbgpu_thread_bind_device(-1);
	};
	std::vector<std::thread> threads;
	for (int r = 0; r < world; r++)
		threads.emplace_back(worker, r);
	for (auto &t : threads)
		t.join();
	if (fail.status != NBGPU_OK) {
		set_error("%s", fail.text.c_str());
		return fail.status;
	}
	int status = 0;
	for (int r = 0; r < world; r++)
		status |= asm_status[r];
	if (report) {
		memset(report, 0, sizeof(*report));
		report->N = 2 * md->N_nod;
		report->solver_iters = iters[0];
		report->solver_status = solve_status[0];
		report->solver_residual = resid[0];
		for (int r = 0; r < world; r++) {
			report->ms_upload = std::max(report->ms_upload, ms_phase[4 * r]);
			report->ms_assembly = std::max(report->ms_assembly, ms_phase[4 * r + 1]);
			report->ms_solve = std::max(report->ms_solve, ms_phase[4 * r + 2]);
			report->ms_post = std::max(report->ms_post, ms_phase[4 * r + 3]);
		}
	}
	if (status)
		return NBGPU_DISTORTED_ELEMENT;
	if (strain) {
		// strain at the Gauss points (pipeline.c:266-319) from the gathered displacement: one pass over
		// the elements on the caller's device
		nbgpu_mesh_t *mesh = nullptr;
		nbgpu_elem_tables_t tab;
		if (tables)
			tab = *tables;
		else
			NB_TRY(nbgpu_elem_tables_default(md->nodes_per_elem, &tab));
		const uint32_t n_gp = md->nodes_per_elem == 4 ? 4 : 1;
		const size_t n_strain = (size_t)3 * n_gp * md->N_elems, N = 2 * (size_t)md->N_nod;
		double *d_buf = nullptr;
		int st = nbgpu_mesh_create(md->N_nod, md->nod, md->N_elems, md->nodes_per_elem, md->adj, &mesh);
		if (st == NBGPU_OK)
			st = nbgpu_malloc((void **)&d_buf, (N + n_strain + 2) * sizeof(double));
		if (st == NBGPU_OK)
			st = nbgpu_copy_h2d(d_buf, displacement, N * sizeof(double));
		if (st == NBGPU_OK)
			st = nbgpu_compute_strain(mesh, &tab, d_buf, d_buf + N);
		if (st == NBGPU_OK)
			st = nbgpu_copy_d2h(strain, d_buf + N, n_strain * sizeof(double));
		nbgpu_free(d_buf);
		nbgpu_mesh_destroy(mesh);
		NB_TRY(st);
	}
	return NBGPU_OK;
}

/* nb_sparse_solve_CG_precond_Jacobi / nb_sparse_solve_conjugate_gradient on a host nb_sparse_t (its three
 * arrays) over n_devices GPUs: contiguous row blocks, one worker thread per GPU. */
int nbgpu_solve_rows_multi(int n_devices, int jacobi, uint32_t N, const uint32_t *rows_size,
			   uint32_t *const *rows_index, double *const *rows_values, const double *b, double *x,
			   uint32_t max_iter, double tolerance, uint32_t *niter_performed, double *tolerance_reached)
{
	NB_ARG(rows_size != nullptr && rows_index != nullptr && rows_values != nullptr && b != nullptr && x != nullptr);
	const int world = usable_devices(n_devices);
	// row blocks of (almost) equal entry counts, cut at even rows (2 dofs per node stay together)
	std::vector<uint64_t> ptr((size_t)N + 1, 0);
	for (uint32_t i = 0; i < N; i++)
		ptr[i + 1] = ptr[i] + rows_size[i];
	std::vector<uint32_t> row_starts(world + 1, 0);
	for (int r = 1; r < world; r++) {
		const uint64_t want = ptr[N] * r / world;
		uint32_t cut = (uint32_t)(std::lower_bound(ptr.begin(), ptr.end(), want) - ptr.begin());
		cut &= ~1u;
		row_starts[r] = std::max(cut, row_starts[r - 1]);
	}
	row_starts[world] = N;
	for (int r = 0; r < world; r++)
		if (row_starts[r] == row_starts[r + 1]) {
			set_error("matrix too small for %d devices", world);
			return NBGPU_ERR_ARG;
		}
	std::vector<nbgpu_dist_plan_t *> plan(world, nullptr);
	std::vector<nbgpu_dist_t *> dist(world, nullptr);
	std::vector<int> device_of(world);
	for (int r = 0; r < world; r++)
		device_of[r] = r;
	std::vector<uint32_t> iters(world, 0);
	std::vector<double> resid(world, 0.0);
	std::vector<int> solve_status(world, 0);
	Barrier bar(world);
	Failure fail;
	auto worker = [&](int r) {
		const uint32_t r0 = row_starts[r], r1 = row_starts[r + 1], n_loc = r1 - r0;
		const uint64_t nnz = ptr[r1] - ptr[r0];
		nbgpu_matrix_t *A = nullptr;
		double *d_b = nullptr;
		int st = nbgpu_thread_bind_device(device_of[r]);
		std::vector<uint32_t> cols(nnz);
		std::vector<double> vals(nnz);
		for (uint32_t i = r0; i < r1; i++) {
			memcpy(cols.data() + (ptr[i] - ptr[r0]), rows_index[i], rows_size[i] * sizeof(uint32_t));
			memcpy(vals.data() + (ptr[i] - ptr[r0]), rows_values[i], rows_size[i] * sizeof(double));
		}
		if (st == NBGPU_OK)
			st = nbgpu_dist_plan_create(r, world, row_starts.data(), rows_size + r0, cols.data(), &plan[r]);
		fail.note(st);
		bar.wait();
		if (!fail.any()) {
			// what every other rank needs from me: its halo entries inside my row range, in its order
			nbgpu_dist_plan_t *P = plan[r];
			std::vector<uint32_t> send_counts(world, 0), dst_offsets(world, 0), send_global;
			for (int d = 0; d < world; d++) {
				if (d == r)
					continue;
				const std::vector<uint32_t> &H = plan[d]->halo_global;
				const size_t a = std::lower_bound(H.begin(), H.end(), r0) - H.begin();
				const size_t e = std::lower_bound(H.begin(), H.end(), r1) - H.begin();
				send_counts[d] = (uint32_t)(e - a);
				send_global.insert(send_global.end(), H.begin() + a, H.begin() + e);
				const uint32_t pos = (uint32_t)a;
				dst_offsets[d] = pos < plan[d]->n_lo ? pos : plan[d]->off_up + (pos - plan[d]->n_lo);
			}
			st = nbgpu_dist_plan_set_sends(P, send_counts.data(), send_global.empty() ? nullptr : send_global.data(),
						       dst_offsets.data());
			if (st == NBGPU_OK)
				st = nbgpu_matrix_create_local(n_loc, P->ext_len, P->off_own, rows_size + r0, P->cols_local.data(),
							       vals.data(), &A);
			if (st == NBGPU_OK)
				st = nbgpu_dist_create(r, world, P->ext_len, nullptr, &dist[r]);
			fail.note(st);
		}
		cols.clear();
		cols.shrink_to_fit();
		vals.clear();
		vals.shrink_to_fit();
		bar.wait();
		if (!fail.any()) {
			st = nbgpu_dist_connect_local(dist[r], dist.data(), device_of.data());
			if (st == NBGPU_OK)
				st = nbgpu_malloc((void **)&d_b, 2 * ((size_t)n_loc + 2) * sizeof(double));
			if (st == NBGPU_OK)
				st = upload_vector(d_b, b + r0, n_loc);
			if (st == NBGPU_OK)
				st = upload_vector(d_b + n_loc + 2, x + r0, n_loc);
			fail.note(st);
		}
		bar.wait();
		if (!fail.any()) {
			double *d_x = d_b + n_loc + 2;
			st = jacobi ? nbgpu_dist_pcg_jacobi(dist[r], plan[r], A, d_b, d_x, max_iter, tolerance, &iters[r], &resid[r])
				    : nbgpu_dist_cg(dist[r], plan[r], A, d_b, d_x, max_iter, tolerance, &iters[r], &resid[r]);
			if (st == NBGPU_OK || st == NBGPU_NOT_CONVERGED) {
				solve_status[r] = st;
				st = download_vector(x + r0, d_x, n_loc);
			}
			fail.note(st);
		}
		bar.wait();
		nbgpu_free(d_b);
		nbgpu_matrix_destroy(A);
		nbgpu_dist_destroy(dist[r]);
		bar.wait();   // the plans are read by the peers until here
		nbgpu_dist_plan_destroy(plan[r]);
		nbgpu_thread_bind_device(-1);
	};
	std::vector<std::thread> threads;
	for (int r = 0; r < world; r++)
		threads.emplace_back(worker, r);
	for (auto &t : threads)
		t.join();
	if (fail.status != NBGPU_OK) {
		set_error("%s", fail.text.c_str());
		return fail.status;
	}
	if (niter_performed)
		*niter_performed = iters[0];
	if (tolerance_reached)
		*tolerance_reached = resid[0];
	return solve_status[0];
}

}  // extern "C"
