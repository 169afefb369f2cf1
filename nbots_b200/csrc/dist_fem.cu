// dist_fem.cu -- the static-elasticity FEM path on a row-partitioned mesh: every rank assembles ITS
// rows of K straight into its rank-local SELL block on the device (K never visits the host) and the
// ranks solve together (dist.cu).
//
// Reference call shape: nb_fem_compute_2D_Solid_Mechanics
//   (sources/nb/pde_bot/finite_element/solid_mechanics/static_elasticity2D.c:31-97): pattern ->
//   pipeline_assemble_system -> nb_fem_set_bconditions -> nb_sparse_solve_CG_precond_Jacobi.
//
// Partition: contiguous NODE ranges in the mesh's own numbering (both dofs of a node
// on one rank); structured meshes pass ranges that are whole grid lines (slabs), any other mesh any
// contiguous ranges (nbgpu_partition_nodes balances the node counts).  Rank r integrates the sub-mesh
// of every element that touches one of its nodes -- elements on a cut are integrated on both sides,
// so assembly needs no communication -- and keeps the rows of its own nodes only.
//
// The sub-mesh is numbered like the block's column space (dist.cu): lower ghost nodes | owned nodes |
// upper ghost nodes, ascending global id inside each part, padded to whole 128-byte lines.  Hence a
// column id of the block IS 2 * sub-mesh node + dof, local element ids ascend with global ones, and
// the row-parallel GATHER kernel adds every entry's contributions in the reference's element order:
// the rank-local rows are bit-identical to the rows of the single-GPU (and the reference's) matrix.
//
// Every rank is handed the whole mesh description (what the reference's caller holds anyway), so it
// derives all halo and send lists itself: the only thing the ranks exchange before the solve is the
// 64-byte IPC handle of their window (or nothing at all when they live in one process).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>

#include "dist_plan.cuh"
#include "matrix.cuh"

using namespace nbgpu;

struct nbgpu_dist_fem_s {
	int rank = 0, world = 1;
	uint32_t n0 = 0, n1 = 0;               // owned nodes [n0, n1)
	uint32_t N_nod_global = 0, N_elems_global = 0, npe = 0;
	std::vector<uint32_t> elems;           // global ids of the sub-mesh's elements, ascending
	std::vector<uint32_t> ghosts;          // global ids of the ghost nodes, ascending
	uint32_t n_lo_nodes = 0;               // ghosts below n0
	nbgpu_dist_plan_t *plan = nullptr;
	nbgpu_dist_t *dist = nullptr;
	nbgpu_matrix_t *K = nullptr;
	nbgpu_mesh_t *mesh = nullptr;
	nbgpu_dirichlet_t *dirichlet = nullptr;
	uint32_t n_neu = 0;
	uint32_t *d_neu_dof = nullptr;
	double *d_neu_add = nullptr;
	double *d_F = nullptr, *d_x = nullptr;
	nbgpu_elem_tables_t tables;
	nbgpu_assembly_params_t ap;
	std::vector<uint8_t> en_loc;
	std::vector<double> scale_loc;
	bool have_x = false, assembled = false;
	double ms_setup = 0;
};

namespace {

double now_ms()
{
	using namespace std::chrono;
	return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

inline int owner_of(const uint32_t *node_starts, int world, uint32_t node)
{
	return (int)(std::upper_bound(node_starts, node_starts + world + 1, node) - node_starts) - 1;
}

}  // namespace

extern "C" {

/* contiguous node ranges of (almost) equal size; boundaries are multiples of `align` nodes where
 * the mesh is large enough (align = nodes per grid line gives slabs of whole lines) */
int nbgpu_partition_nodes(uint32_t N_nod, int world, uint32_t align, uint32_t *node_starts)
{
	NB_ARG(node_starts != nullptr && world >= 1 && world <= kMaxRanks);
	if (align == 0)
		align = 1;
	const uint64_t units = N_nod / align;   // whole units; the remainder goes to the last rank
	node_starts[0] = 0;
	for (int r = 1; r < world; r++) {
		uint64_t u = units >= (uint64_t)world ? (units * r) / world : 0;
		uint64_t b = units >= (uint64_t)world ? u * align : ((uint64_t)N_nod * r) / world;
		node_starts[r] = (uint32_t)std::max<uint64_t>(b, node_starts[r - 1]);
	}
	node_starts[world] = N_nod;
	return NBGPU_OK;
}

int nbgpu_dist_fem_destroy(nbgpu_dist_fem_t *S)
{
	if (!S)
		return NBGPU_OK;
	if (ctx().ready) {
		cudaSetDevice(ctx().device);
		cudaStreamSynchronize(ctx().stream);
		nbgpu::dfree(S->d_F);
		nbgpu::dfree(S->d_neu_add);
	}
	nbgpu_dirichlet_destroy(S->dirichlet);
	nbgpu_matrix_destroy(S->K);
	nbgpu_mesh_destroy(S->mesh);
	nbgpu_dist_destroy(S->dist);
	nbgpu_dist_plan_destroy(S->plan);
	delete S;
	return NBGPU_OK;
}

}  // extern "C"

namespace {

// Host part of nbgpu_dist_fem_create: this rank's sub-mesh (elements, ghost nodes, column-space numbering),
// the partition plan with its send lists, and the pattern of the owned rows.  No device is touched.
struct SubMesh {
	std::vector<uint32_t> adj_loc;     // [npe n_el] column-space node ids
	std::vector<double> nod_loc;       // [2 n_sub_nodes]
	std::vector<uint32_t> rows_size;   // [N_loc]
	uint32_t n_sub_nodes = 0;
};

int plan_sub_mesh(const nbgpu_mesh_desc_t *md, int rank, int world, const uint32_t *node_starts, nbgpu_dist_fem_t *S,
		  SubMesh *M)
{
	NB_ARG(md != nullptr && node_starts != nullptr);
	NB_ARG(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world);
	NB_ARG(md->nodes_per_elem == 3 || md->nodes_per_elem == 4);
	NB_ARG(node_starts[0] == 0 && node_starts[world] == md->N_nod);
	for (int r = 0; r < world; r++)
		NB_ARG(node_starts[r] <= node_starts[r + 1]);
	const uint32_t npe = md->nodes_per_elem, n0 = node_starts[rank], n1 = node_starts[rank + 1];
	S->rank = rank;
	S->world = world;
	S->n0 = n0;
	S->n1 = n1;
	S->N_nod_global = md->N_nod;
	S->N_elems_global = md->N_elems;
	S->npe = npe;

	// ---- one pass over the elements: my sub-mesh, and EVERY rank's ghost nodes (cut elements only)
	std::vector<std::vector<uint32_t>> ghost_of(world);
	for (uint32_t e = 0; e < md->N_elems; e++) {
		const uint32_t *v = md->adj + (size_t)e * npe;
		uint32_t lo = v[0], hi = v[0];
		for (uint32_t i = 1; i < npe; i++) {
			lo = std::min(lo, v[i]);
			hi = std::max(hi, v[i]);
		}
		if (hi >= md->N_nod) {
			set_error("element %u references node %u >= N_nod", e, hi);
			return NBGPU_ERR_ARG;
		}
		if (hi >= n0 && lo < n1) {
			bool mine = false;
			for (uint32_t i = 0; i < npe; i++)
				mine |= v[i] >= n0 && v[i] < n1;
			if (mine)
				S->elems.push_back(e);
		}
		const int r_lo = owner_of(node_starts, world, lo), r_hi = owner_of(node_starts, world, hi);
		if (r_lo == r_hi)
			continue;
		int own[4];
		for (uint32_t i = 0; i < npe; i++)
			own[i] = owner_of(node_starts, world, v[i]);
		for (uint32_t i = 0; i < npe; i++)
			for (uint32_t j = 0; j < npe; j++)
				if (own[i] != own[j])
					ghost_of[own[i]].push_back(v[j]);   // v[j] is a ghost node of the owner of v[i]
	}
	for (auto &g : ghost_of) {
		std::sort(g.begin(), g.end());
		g.erase(std::unique(g.begin(), g.end()), g.end());
	}
	S->ghosts = ghost_of[rank];
	const std::vector<uint32_t> &G = S->ghosts;
	const uint32_t n_lo_nodes = (uint32_t)(std::lower_bound(G.begin(), G.end(), n0) - G.begin());
	S->n_lo_nodes = n_lo_nodes;

	// ---- the plan, filled from the node lists (no pattern scan, no list exchange between the ranks)
	nbgpu_dist_plan_t *P = new nbgpu_dist_plan_t();
	S->plan = P;
	P->rank = rank;
	P->world = world;
	P->row_starts.resize(world + 1);
	for (int r = 0; r <= world; r++)
		P->row_starts[r] = 2 * node_starts[r];
	P->N_loc = 2 * (n1 - n0);
	P->n_halo = 2 * (uint32_t)G.size();
	P->n_lo = 2 * n_lo_nodes;
	P->n_hi = P->n_halo - P->n_lo;
	if (nbgpu_dist_ext_layout(P->n_lo, P->N_loc, P->n_hi, &P->off_own, &P->off_up, &P->ext_len) != NBGPU_OK)
		return NBGPU_ERR_ARG;
	P->halo_global.resize(P->n_halo);
	P->recv_counts.assign(world, 0);
	for (size_t i = 0; i < G.size(); i++) {
		P->halo_global[2 * i] = 2 * G[i];
		P->halo_global[2 * i + 1] = 2 * G[i] + 1;
		P->recv_counts[owner_of(node_starts, world, G[i])] += 2;
	}
	// what every other rank d needs from me, and where it goes in d's column space
	P->send_ptr.assign(world + 1, 0);
	P->dst_offset.assign(world, 0);
	for (int d = 0; d < world; d++) {
		uint32_t cnt = 0;
		if (d != rank) {
			const std::vector<uint32_t> &Gd = ghost_of[d];
			const size_t a = std::lower_bound(Gd.begin(), Gd.end(), n0) - Gd.begin();
			const size_t b = std::lower_bound(Gd.begin(), Gd.end(), n1) - Gd.begin();
			for (size_t i = a; i < b; i++) {
				P->send_local.push_back(2 * (Gd[i] - n0));
				P->send_local.push_back(2 * (Gd[i] - n0) + 1);
			}
			cnt = 2 * (uint32_t)(b - a);
			const uint32_t n_lo_d = 2 * (uint32_t)(std::lower_bound(Gd.begin(), Gd.end(), node_starts[d]) - Gd.begin());
			const uint32_t n_loc_d = 2 * (node_starts[d + 1] - node_starts[d]);
			uint32_t off_own_d = 0, off_up_d = 0;
			nbgpu_dist_ext_layout(n_lo_d, n_loc_d, 2 * (uint32_t)Gd.size() - n_lo_d, &off_own_d, &off_up_d, nullptr);
			const uint32_t pos = 2 * (uint32_t)a;   // position of my block in d's halo list
			P->dst_offset[d] = pos < n_lo_d ? pos : off_up_d + (pos - n_lo_d);
		}
		P->send_ptr[d + 1] = P->send_ptr[d] + cnt;
	}
	P->have_sends = true;

	// ---- sub-mesh in column-space numbering
	const uint32_t n_sub_nodes = P->ext_len / 2;
	const uint32_t own_node0 = P->off_own / 2, up_node0 = P->off_up / 2;
	auto local_node = [&](uint32_t g) -> uint32_t {
		if (g >= n0 && g < n1)
			return own_node0 + (g - n0);
		const uint32_t h = (uint32_t)(std::lower_bound(G.begin(), G.end(), g) - G.begin());
		if (h >= G.size() || G[h] != g)
			return 0xFFFFFFFFu;
		return h < n_lo_nodes ? h : up_node0 + (h - n_lo_nodes);
	};
	const uint32_t n_el = (uint32_t)S->elems.size();
	std::vector<uint32_t> &adj_loc = M->adj_loc;
	std::vector<double> &nod_loc = M->nod_loc;
	adj_loc.assign((size_t)npe * n_el, 0);
	nod_loc.assign(2 * (size_t)n_sub_nodes, 0.0);
	M->n_sub_nodes = n_sub_nodes;
#pragma omp parallel for schedule(static)
	for (int64_t t = 0; t < (int64_t)n_el; t++) {
		const uint32_t *v = md->adj + (size_t)S->elems[t] * npe;
		for (uint32_t i = 0; i < npe; i++)
			adj_loc[(size_t)t * npe + i] = local_node(v[i]);
	}
	for (uint32_t g = n0; g < n1; g++) {
		nod_loc[2 * (size_t)(own_node0 + g - n0)] = md->nod[2 * (size_t)g];
		nod_loc[2 * (size_t)(own_node0 + g - n0) + 1] = md->nod[2 * (size_t)g + 1];
	}
	for (size_t h = 0; h < G.size(); h++) {
		const uint32_t l = h < n_lo_nodes ? (uint32_t)h : up_node0 + ((uint32_t)h - n_lo_nodes);
		nod_loc[2 * (size_t)l] = md->nod[2 * (size_t)G[h]];
		nod_loc[2 * (size_t)l + 1] = md->nod[2 * (size_t)G[h] + 1];
	}

	// ---- pattern of the owned rows (nb_mesh2D_load_graph NODES_LINKED_BY_ELEMS + nb_sparse_create,
	// load_graph.c:230-328, sparse.c:20-60): a node is linked to every node it shares an element with;
	// columns ascending.  Built per owned node from its element list, already in column-space ids.
	const uint32_t n_own = n1 - n0;
	std::vector<uint32_t> e_ptr((size_t)n_own + 1, 0);
	for (uint32_t t = 0; t < n_el; t++)
		for (uint32_t i = 0; i < npe; i++) {
			const uint32_t l = adj_loc[(size_t)t * npe + i];
			if (l >= own_node0 && l < own_node0 + n_own)
				e_ptr[l - own_node0 + 1]++;
		}
	for (uint32_t i = 0; i < n_own; i++)
		e_ptr[i + 1] += e_ptr[i];
	std::vector<uint32_t> e_of(e_ptr[n_own]);
	{
		std::vector<uint32_t> next(e_ptr.begin(), e_ptr.end() - 1);
		for (uint32_t t = 0; t < n_el; t++)
			for (uint32_t i = 0; i < npe; i++) {
				const uint32_t l = adj_loc[(size_t)t * npe + i];
				if (l >= own_node0 && l < own_node0 + n_own)
					e_of[next[l - own_node0]++] = t;
			}
	}
	std::vector<uint32_t> &rows_size = M->rows_size;
	rows_size.assign((size_t)P->N_loc, 0);
	std::vector<uint64_t> nb_ptr((size_t)n_own + 1, 0);
	// two passes: count, then fill
	auto neighbours = [&](uint32_t i, uint32_t *buf) -> uint32_t {
		uint32_t n = 0;
		buf[n++] = own_node0 + i;
		for (uint32_t k = e_ptr[i]; k < e_ptr[i + 1]; k++)
			for (uint32_t j = 0; j < npe; j++)
				buf[n++] = adj_loc[(size_t)e_of[k] * npe + j];
		std::sort(buf, buf + n);
		return (uint32_t)(std::unique(buf, buf + n) - buf);
	};
	uint32_t max_deg = 0;
	for (uint32_t i = 0; i < n_own; i++)
		max_deg = std::max(max_deg, e_ptr[i + 1] - e_ptr[i]);
	const size_t buf_len = (size_t)max_deg * npe + 1;
#pragma omp parallel
	{
		std::vector<uint32_t> buf(buf_len);
#pragma omp for schedule(static)
		for (int64_t i = 0; i < (int64_t)n_own; i++) {
			const uint32_t n = neighbours((uint32_t)i, buf.data());
			rows_size[2 * (size_t)i] = rows_size[2 * (size_t)i + 1] = 2 * n;
			nb_ptr[i + 1] = n;
		}
	}
	for (uint32_t i = 0; i < n_own; i++)
		nb_ptr[i + 1] += nb_ptr[i];
	P->nnz = 4 * nb_ptr[n_own];
	P->cols_local.resize(P->nnz);
#pragma omp parallel
	{
		std::vector<uint32_t> buf(buf_len);
#pragma omp for schedule(static)
		for (int64_t i = 0; i < (int64_t)n_own; i++) {
			const uint32_t n = neighbours((uint32_t)i, buf.data());
			uint32_t *r0 = P->cols_local.data() + 4 * nb_ptr[i], *r1 = r0 + 2 * (size_t)n;
			for (uint32_t k = 0; k < n; k++) {
				r0[2 * k] = r1[2 * k] = 2 * buf[k];
				r0[2 * k + 1] = r1[2 * k + 1] = 2 * buf[k] + 1;
			}
		}
	}
	plan_visit_order(P, rows_size.data());

	return NBGPU_OK;
}

}  // namespace

extern "C" {

int nbgpu_dist_fem_create(const nbgpu_mesh_desc_t *md, int rank, int world, const uint32_t *node_starts,
			  const nbgpu_elem_tables_t *tables, const double D[4], double density, uint32_t n_neu,
			  const uint32_t *neu_dof, const double *neu_add, uint32_t n_dir, const uint32_t *dir_dof,
			  const double *dir_val, int self_weight, const double gravity[2], double thickness,
			  void *ipc_handle_out, nbgpu_dist_fem_t **out)
{
	NB_INIT();
	NB_ARG(md != nullptr && out != nullptr && D != nullptr && node_starts != nullptr);
	const double t0 = now_ms();
	nbgpu_dist_fem_t *S = new nbgpu_dist_fem_t();
	SubMesh sub;
	{
		const int pst = plan_sub_mesh(md, rank, world, node_starts, S, &sub);
		if (pst != NBGPU_OK) {
			nbgpu_dist_fem_destroy(S);
			return pst;
		}
	}
	nbgpu_dist_plan_t *P = S->plan;
	const uint32_t npe = S->npe, n0 = S->n0, n1 = S->n1, n_lo_nodes = S->n_lo_nodes;
	const std::vector<uint32_t> &G = S->ghosts;
	const uint32_t own_node0 = P->off_own / 2, up_node0 = P->off_up / 2, n_sub_nodes = sub.n_sub_nodes;
	const uint32_t n_el = (uint32_t)S->elems.size();
	std::vector<uint32_t> &adj_loc = sub.adj_loc, &rows_size = sub.rows_size;
	std::vector<double> &nod_loc = sub.nod_loc;
	auto local_node = [&](uint32_t g) -> uint32_t {
		if (g >= n0 && g < n1)
			return own_node0 + (g - n0);
		const uint32_t h = (uint32_t)(std::lower_bound(G.begin(), G.end(), g) - G.begin());
		if (h >= G.size() || G[h] != g)
			return 0xFFFFFFFFu;
		return h < n_lo_nodes ? h : up_node0 + (h - n_lo_nodes);
	};

	// ---- device objects
	int st = nbgpu_matrix_create_local(P->N_loc, P->ext_len, P->off_own, rows_size.data(), P->cols_local.data(),
					   nullptr, &S->K);
	if (st == NBGPU_OK)
		st = nbgpu_mesh_create(n_sub_nodes, nod_loc.data(), n_el, npe, adj_loc.data(), &S->mesh);
	if (st == NBGPU_OK)
		st = nbgpu_dist_create(rank, world, P->ext_len, ipc_handle_out, &S->dist);
	if (st == NBGPU_OK) {
		cudaError_t e = nbgpu::dmalloc(&S->d_F, 2 * (size_t)std::max<uint32_t>(P->N_loc, 2) * sizeof(double));
		if (e != cudaSuccess) {
			set_error("dist_fem vectors: %s", cudaGetErrorString(e));
			cudaGetLastError();
			st = NBGPU_ERR_NOMEM;
		} else {
			S->d_x = S->d_F + std::max<uint32_t>(P->N_loc, 2);
			cudaMemsetAsync(S->d_F, 0, 2 * (size_t)std::max<uint32_t>(P->N_loc, 2) * sizeof(double), ctx().stream);
		}
	}
	// ---- boundary conditions restricted to this rank (list order kept)
	if (st == NBGPU_OK) {
		std::vector<uint32_t> ldof;
		std::vector<double> lval;
		for (uint32_t k = 0; k < n_neu; k++) {
			const uint32_t g = neu_dof[k] >> 1;
			if (g >= n0 && g < n1) {
				ldof.push_back(neu_dof[k] - 2 * n0);
				lval.push_back(neu_add[k]);
			}
		}
		S->n_neu = (uint32_t)ldof.size();
		if (S->n_neu) {
			void *buf = nullptr;
			cudaError_t e = nbgpu::dmalloc(&buf, (size_t)S->n_neu * (sizeof(uint32_t) + sizeof(double)));
			if (e == cudaSuccess) {
				S->d_neu_add = (double *)buf;
				S->d_neu_dof = (uint32_t *)(S->d_neu_add + S->n_neu);
				e = cudaMemcpy(S->d_neu_add, lval.data(), (size_t)S->n_neu * sizeof(double), cudaMemcpyHostToDevice);
			}
			if (e == cudaSuccess)
				e = cudaMemcpy(S->d_neu_dof, ldof.data(), (size_t)S->n_neu * sizeof(uint32_t), cudaMemcpyHostToDevice);
			if (e != cudaSuccess) {
				set_error("dist_fem boundary lists: %s", cudaGetErrorString(e));
				cudaGetLastError();
				st = NBGPU_ERR_CUDA;
			}
		}
		ldof.clear();
		lval.clear();
		for (uint32_t k = 0; k < n_dir && st == NBGPU_OK; k++) {
			const uint32_t l = local_node(dir_dof[k] >> 1);
			if (l == 0xFFFFFFFFu)
				continue;   // neither owned nor a ghost: no entry of my rows refers to it
			ldof.push_back(2 * l + (dir_dof[k] & 1));
			lval.push_back(dir_val[k]);
		}
		if (st == NBGPU_OK)
			st = nbgpu_dirichlet_create(P->ext_len, (uint32_t)ldof.size(), ldof.data(), lval.data(), &S->dirichlet);
	}
	if (st != NBGPU_OK) {
		nbgpu_dist_fem_destroy(S);
		return st;
	}
	if (tables)
		S->tables = *tables;
	else
		nbgpu_elem_tables_default(npe, &S->tables);
	memset(&S->ap, 0, sizeof(S->ap));
	memcpy(S->ap.D, D, sizeof(S->ap.D));
	for (int k = 0; k < 4; k++)
		S->ap.D_void[k] = 1e-6;              /* pipeline.c:93 */
	S->ap.density = density;
	S->ap.density_void = 1e-6;               /* pipeline.c:94 */
	S->ap.thickness = thickness;
	S->ap.self_weight = self_weight;
	if (gravity) {
		S->ap.gravity[0] = gravity[0];
		S->ap.gravity[1] = gravity[1];
	}
	S->ap.mode = NBGPU_ASSEMBLY_GATHER;
	S->ms_setup = now_ms() - t0;
	*out = S;
	return NBGPU_OK;
}

/* Host logic only (no device): the partition plan nbgpu_dist_fem_create derives from the mesh -- halo list,
 * column-space layout, send lists, local column ids of the owned rows (nbgpu_dist_plan_* read it back).
 * rows_size may be NULL or [2 * owned nodes]. */
int nbgpu_dist_plan_from_mesh(const nbgpu_mesh_desc_t *md, int rank, int world, const uint32_t *node_starts,
			      uint32_t *rows_size, nbgpu_dist_plan_t **out)
{
	NB_ARG(out != nullptr);
	nbgpu_dist_fem_t S;
	SubMesh sub;
	const int st = plan_sub_mesh(md, rank, world, node_starts, &S, &sub);
	if (st != NBGPU_OK) {
		nbgpu_dist_plan_destroy(S.plan);
		return st;
	}
	if (rows_size)
		memcpy(rows_size, sub.rows_size.data(), sub.rows_size.size() * sizeof(uint32_t));
	*out = S.plan;
	S.plan = nullptr;
	return NBGPU_OK;
}

/* the plan's send side: counts per destination, the global rows sent (grouped by destination) and where each block
 * lands in the destination's column space; send_global may be NULL to get the counts first */
int nbgpu_dist_plan_sends(const nbgpu_dist_plan_t *P, uint32_t *send_counts, uint32_t *send_global, uint32_t *dst_offsets)
{
	NB_ARG(P != nullptr && P->have_sends);
	for (int r = 0; r < P->world; r++) {
		if (send_counts)
			send_counts[r] = P->send_ptr[r + 1] - P->send_ptr[r];
		if (dst_offsets)
			dst_offsets[r] = P->dst_offset[r];
	}
	if (send_global)
		for (size_t j = 0; j < P->send_local.size(); j++)
			send_global[j] = P->send_local[j] + P->row_starts[P->rank];
	return NBGPU_OK;
}

/* sizes of this rank's block; any pointer may be NULL */
int nbgpu_dist_fem_info(const nbgpu_dist_fem_t *S, uint32_t *N_loc, uint64_t *nnz_loc, uint32_t *n_halo,
			uint32_t *n_elems_loc, uint64_t *ext_len, double *ms_setup)
{
	NB_ARG(S != nullptr);
	if (N_loc)
		*N_loc = S->plan->N_loc;
	if (nnz_loc)
		*nnz_loc = S->plan->nnz;
	if (n_halo)
		*n_halo = S->plan->n_halo;
	if (n_elems_loc)
		*n_elems_loc = (uint32_t)S->elems.size();
	if (ext_len)
		*ext_len = S->plan->ext_len;
	if (ms_setup)
		*ms_setup = S->ms_setup;
	return NBGPU_OK;
}

/* the objects behind the session, for callers that drive nbgpu_dist_* / nbgpu_matrix_* themselves */
nbgpu_matrix_t *nbgpu_dist_fem_matrix(nbgpu_dist_fem_t *S) { return S ? S->K : nullptr; }
nbgpu_dist_plan_t *nbgpu_dist_fem_plan(nbgpu_dist_fem_t *S) { return S ? S->plan : nullptr; }
nbgpu_dist_t *nbgpu_dist_fem_dist(nbgpu_dist_fem_t *S) { return S ? S->dist : nullptr; }
double *nbgpu_dist_fem_rhs(nbgpu_dist_fem_t *S) { return S ? S->d_F : nullptr; }
double *nbgpu_dist_fem_solution(nbgpu_dist_fem_t *S) { return S ? S->d_x : nullptr; }

int nbgpu_dist_fem_connect(nbgpu_dist_fem_t *S, const void *all_handles, const uint64_t *all_ext_len)
{
	NB_ARG(S != nullptr);
	return nbgpu_dist_connect(S->dist, all_handles, all_ext_len);
}

int nbgpu_dist_fem_connect_local(nbgpu_dist_fem_t *S, nbgpu_dist_fem_t *const *all, const int *device_of)
{
	NB_ARG(S != nullptr && all != nullptr);
	nbgpu_dist_t *d[kMaxRanks] = {};
	for (int r = 0; r < S->world; r++) {
		NB_ARG(all[r] != nullptr);
		d[r] = all[r]->dist;
	}
	return nbgpu_dist_connect_local(S->dist, d, device_of);
}

/* pipeline_assemble_system + nb_fem_set_bconditions for this rank's rows; enabled / elem_scale are the
 * GLOBAL per-element arrays (NULL = all enabled / factor 1).  Returns 0, or 1 with the lowest distorted
 * GLOBAL element id of this rank's sub-mesh in *first_bad. */
int nbgpu_dist_fem_assemble(nbgpu_dist_fem_t *S, const uint8_t *enabled, const double *elem_scale,
			    uint32_t *first_bad)
{
	NB_INIT();
	NB_ARG(S != nullptr);
	const uint32_t n_el = (uint32_t)S->elems.size();
	const uint8_t *en = nullptr;
	const double *sc = nullptr;
	if (enabled) {
		S->en_loc.resize(n_el);
		for (uint32_t t = 0; t < n_el; t++)
			S->en_loc[t] = enabled[S->elems[t]];
		en = S->en_loc.data();
	}
	if (elem_scale) {
		S->scale_loc.resize(n_el);
		for (uint32_t t = 0; t < n_el; t++)
			S->scale_loc[t] = elem_scale[S->elems[t]];
		sc = S->scale_loc.data();
	}
	uint32_t bad = 0xFFFFFFFFu;
	const int st = nbgpu_assemble_elasticity2d(S->K, S->mesh, &S->tables, &S->ap, en, sc, S->d_F, &bad);
	if (first_bad)
		*first_bad = bad < n_el ? S->elems[bad] : 0xFFFFFFFFu;
	if (st != NBGPU_OK)
		return st;
	NB_TRY(nbgpu_vector_add_entries_dev(S->d_F, S->n_neu, S->d_neu_dof, S->d_neu_add));
	NB_TRY(nbgpu_dirichlet_apply(S->K, S->d_F, S->dirichlet));
	S->assembled = true;
	return NBGPU_OK;
}

/* Jacobi-PCG over all ranks (collective).  warm_start: keep the previous solution as x0, else x0 = 0
 * (static_elasticity2D.c:87).  max_iter 0 = global N, tolerance <= 0 = 1e-8 (:88-90). */
int nbgpu_dist_fem_solve(nbgpu_dist_fem_t *S, int warm_start, uint32_t max_iter, double tolerance,
			 uint32_t *niter, double *tol_reached)
{
	NB_INIT();
	NB_ARG(S != nullptr && S->assembled);
	if (!warm_start || !S->have_x)
		NB_CUDA(cudaMemsetAsync(S->d_x, 0, (size_t)S->plan->N_loc * sizeof(double), ctx().stream));
	if (max_iter == 0)
		max_iter = 2 * S->N_nod_global;
	if (tolerance <= 0)
		tolerance = 1e-8;
	const int st = nbgpu_dist_pcg_jacobi(S->dist, S->plan, S->K, S->d_F, S->d_x, max_iter, tolerance, niter, tol_reached);
	S->have_x = st == NBGPU_OK || st == NBGPU_NOT_CONVERGED;
	return st;
}

/* this rank's displacements: 2 * (n1 - n0) values for the nodes [n0, n1) */
int nbgpu_dist_fem_results(nbgpu_dist_fem_t *S, double *displacement_owned)
{
	NB_INIT();
	NB_ARG(S != nullptr && displacement_owned != nullptr);
	return download_vector(displacement_owned, S->d_x, S->plan->N_loc);
}

}  // extern "C"
