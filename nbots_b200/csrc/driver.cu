// driver.cu -- host orchestration of the static-elasticity FEM pipeline with
// every heavy object resident on the device.
//
// Reference: nb_fem_compute_2D_Solid_Mechanics and its `solver`
//   (sources/nb/pde_bot/finite_element/solid_mechanics/static_elasticity2D.c:31-97),
//   nb_fem_set_bconditions (solid_mechanics/set_bconditions.c:52-262).
#include <chrono>
#include <cmath>
#include <cstring>

#include "common.cuh"

using namespace nbgpu;

namespace {

double now_ms()
{
	using namespace std::chrono;
	return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// mesh2D.c:650-674: segment and sub-segment lengths
double node_dist(const double *nod, uint32_t a, uint32_t b)
{
	const double dx = nod[2 * a] - nod[2 * b], dy = nod[2 * a + 1] - nod[2 * b + 1];
	return sqrt(dx * dx + dy * dy);
}

void bc_value(const nbgpu_bcond_t &c, const double *xy, double val[2])
{
	if (c.fval) {
		double x[2] = {xy[0], xy[1]};
		val[0] = val[1] = 0.0;
		c.fval(x, 0.0, val);   // bcond_iter get_val with t = 0
	} else {
		val[0] = c.val[0];
		val[1] = c.val[1];
	}
}

struct DofList {
	uint32_t n = 0;
	uint32_t *dof;
	double *val;
	void push(uint32_t d, double v)
	{
		if (dof) {
			dof[n] = d;
			val[n] = v;
		}
		n++;
	}
};

}  // namespace

extern "C" {

int nbgpu_bcond_flatten(const double *nod, const uint32_t *vtx, uint32_t N_sgm, const uint32_t *sgm_sizes,
			const uint32_t *sgm_nodes, uint32_t N_bc, const nbgpu_bcond_t *bc, double factor,
			uint32_t *n_neumann, uint32_t *neumann_dof, double *neumann_add,
			uint32_t *n_dirichlet, uint32_t *dirichlet_dof, double *dirichlet_val)
{
	NB_ARG(nod != nullptr && (N_bc == 0 || bc != nullptr));
	NB_ARG((neumann_dof == nullptr) == (neumann_add == nullptr));
	NB_ARG((dirichlet_dof == nullptr) == (dirichlet_val == nullptr));
	std::vector<uint64_t> sgm_off((size_t)N_sgm + 1, 0);
	for (uint32_t s = 0; s < N_sgm; s++)
		sgm_off[s + 1] = sgm_off[s] + sgm_sizes[s];
	DofList neu{0, neumann_dof, neumann_add}, dir{0, dirichlet_dof, dirichlet_val};
	// set_bconditions.c:57-60: Neumann on segments, Neumann on vertices,
	// Dirichlet on segments, Dirichlet on vertices; each queue in push order
	static const int order[4][2] = {{1, 1}, {1, 0}, {0, 1}, {0, 0}};
	for (int pass = 0; pass < 4; pass++) {
		for (uint32_t b = 0; b < N_bc; b++) {
			const nbgpu_bcond_t &c = bc[b];
			if (c.kind != order[pass][0] || c.where != order[pass][1])
				continue;
			const uint32_t *sn = nullptr;
			uint32_t ns = 0;
			if (c.where) {
				NB_ARG(c.id < N_sgm && sgm_sizes != nullptr && sgm_nodes != nullptr);
				sn = sgm_nodes + sgm_off[c.id];
				ns = sgm_sizes[c.id];
			} else {
				NB_ARG(vtx != nullptr);
			}
			if (c.kind == 1 && c.where == 1 && c.fval) {
				// :87-131 trapezoid rule per sub-segment, half to each end
				if (ns == 0)
					continue;
				uint32_t v1 = sn[0];
				double val1[2], val2[2];
				bc_value(c, nod + 2 * (size_t)v1, val1);
				for (uint32_t i = 0; i + 1 < ns; i++) {
					const uint32_t v2 = sn[i + 1];
					const double len = node_dist(nod, sn[i], v2);
					bc_value(c, nod + 2 * (size_t)v2, val2);
					for (int j = 0; j < 2; j++) {
						if (!c.mask[j])
							continue;
						const double val = 0.5 * (val1[j] + val2[j]) * len;
						neu.push(2 * v1 + j, factor * val * 0.5);
						neu.push(2 * v2 + j, factor * val * 0.5);
					}
					v1 = v2;
					val1[0] = val2[0];
					val1[1] = val2[1];
				}
			} else if (c.kind == 1 && c.where == 1) {
				// :133-170 the value is the segment's total load; every
				// sub-segment takes len_sub/len_sgm of it, half per end node
				if (ns == 0)
					continue;
				const double total = node_dist(nod, sn[0], sn[ns - 1]);
				for (uint32_t i = 0; i + 1 < ns; i++) {
					const double share = node_dist(nod, sn[i], sn[i + 1]) / total;
					const double f = factor * share * 0.5;
					for (int e = 0; e < 2; e++)
						for (int j = 0; j < 2; j++)
							if (c.mask[j])
								neu.push(2 * sn[i + e] + j, f * c.val[j]);
				}
			} else if (c.kind == 1) {
				// :172-188 point load on an input vertex
				const uint32_t v = vtx[c.id];
				for (int j = 0; j < 2; j++)
					if (c.mask[j])
						neu.push(2 * v + j, factor * c.val[j]);
			} else {
				// :190-262 prescribed displacements, node by node, dof by dof
				const uint32_t cnt = c.where ? ns : 1;
				for (uint32_t i = 0; i < cnt; i++) {
					const uint32_t v = c.where ? sn[i] : vtx[c.id];
					double val[2];
					bc_value(c, nod + 2 * (size_t)v, val);
					for (int j = 0; j < 2; j++)
						if (c.mask[j])
							dir.push(2 * v + j, factor * val[j]);
				}
			}
		}
	}
	if (n_neumann)
		*n_neumann = neu.n;
	if (n_dirichlet)
		*n_dirichlet = dir.n;
	return NBGPU_OK;
}

int nbgpu_fem_static_elasticity2d(const nbgpu_mesh_desc_t *md, const nbgpu_elem_tables_t *tables, double E,
				  double poisson, double density, uint32_t N_bc, const nbgpu_bcond_t *bc,
				  int self_weight, const double gravity[2], int analysis2D, double thickness,
				  const uint8_t *enabled, int assembly_mode, double solver_tol,
				  double *displacement, double *strain, nbgpu_fem_report_t *report)
{
	NB_ARG(md != nullptr);
	uint32_t n_neu = 0, n_dir = 0;
	NB_TRY(nbgpu_bcond_flatten(md->nod, md->vtx, md->N_sgm, md->sgm_sizes, md->sgm_nodes, N_bc, bc, 1.0, &n_neu,
				   nullptr, nullptr, &n_dir, nullptr, nullptr));
	std::vector<uint32_t> neu_dof(n_neu + 1), dir_dof(n_dir + 1);
	std::vector<double> neu_add(n_neu + 1), dir_val(n_dir + 1);
	NB_TRY(nbgpu_bcond_flatten(md->nod, md->vtx, md->N_sgm, md->sgm_sizes, md->sgm_nodes, N_bc, bc, 1.0, &n_neu,
				   neu_dof.data(), neu_add.data(), &n_dir, dir_dof.data(), dir_val.data()));
	double D[4];
	NB_TRY(nbgpu_constitutive_matrix(E, poisson, analysis2D, D));
	return nbgpu_fem_static_elasticity2d_lists(md, tables, D, density, n_neu, neu_dof.data(),
						   neu_add.data(), n_dir, dir_dof.data(), dir_val.data(),
						   self_weight, gravity, analysis2D, thickness, enabled,
						   assembly_mode, solver_tol, displacement, strain, report);
}

/* ---- device-resident session: pattern, matrix and mesh are built once; every step
 * re-assembles (optionally with an enabled mask / per-element stiffness factors),
 * re-applies the boundary conditions and solves, warm-started from the previous
 * displacement if asked.  This is the call pattern of the reference's repeated
 * assembly + PCG loop (static_damage2D.c:300-460, void-material hook
 * pipeline.c:93-98) and of a SIMP topology-optimisation loop (BASELINE.json
 * configs[4]); the one-shot driver below is a session with a single step. */
struct nbgpu_fem_session_s {
	uint32_t N = 0, n_gp = 0, N_elems = 0;
	uint64_t nnz = 0;
	nbgpu_matrix_t *K = nullptr;
	nbgpu_mesh_t *mesh = nullptr;
	double *d_vec = nullptr, *d_F = nullptr, *d_x = nullptr, *d_strain = nullptr;
	nbgpu_elem_tables_t tables;
	nbgpu_assembly_params_t ap;
	uint32_t n_neu = 0;
	uint32_t *d_neu_dof = nullptr;        // boundary-condition lists, resident on the device
	double *d_neu_add = nullptr;
	nbgpu_dirichlet_t *dirichlet = nullptr;
	double ms_pattern = 0, ms_upload = 0;
	bool have_x = false;
};

int nbgpu_fem_session_destroy(nbgpu_fem_session_t *S)
{
	if (!S)
		return NBGPU_OK;
	nbgpu_free(S->d_vec);
	nbgpu_free(S->d_neu_add);
	nbgpu_dirichlet_destroy(S->dirichlet);
	nbgpu_mesh_destroy(S->mesh);
	nbgpu_matrix_destroy(S->K);
	delete S;
	return NBGPU_OK;
}

int nbgpu_fem_session_create(const nbgpu_mesh_desc_t *md, const nbgpu_elem_tables_t *tables, const double D[4],
			     double density, uint32_t n_neu, const uint32_t *neu_dof, const double *neu_add,
			     uint32_t n_dir, const uint32_t *dir_dof, const double *dir_val, int self_weight,
			     const double gravity[2], double thickness, int assembly_mode, nbgpu_fem_session_t **out)
{
	NB_INIT();
	NB_ARG(md != nullptr && out != nullptr && D != nullptr);
	NB_ARG(md->nodes_per_elem == 3 || md->nodes_per_elem == 4);
	NB_ARG(!self_weight || gravity != nullptr);
	nbgpu_fem_session_t *S = new nbgpu_fem_session_t();
	S->N = 2 * md->N_nod;
	S->n_gp = md->nodes_per_elem == 4 ? 4 : 1;
	S->N_elems = md->N_elems;
	int st = NBGPU_OK;
	if (tables)
		S->tables = *tables;
	else
		st = nbgpu_elem_tables_default(md->nodes_per_elem, &S->tables);
	memset(&S->ap, 0, sizeof(S->ap));
	memcpy(S->ap.D, D, sizeof(S->ap.D));
	for (int k = 0; k < 4; k++)
		S->ap.D_void[k] = 1e-6;          // pipeline.c:93
	S->ap.density = density;
	S->ap.density_void = 1e-6;               // pipeline.c:94
	S->ap.thickness = thickness;
	S->ap.self_weight = self_weight != 0;
	if (self_weight) {
		S->ap.gravity[0] = gravity[0];
		S->ap.gravity[1] = gravity[1];
	}
	S->ap.mode = assembly_mode;
	S->n_neu = n_neu;

	// (1) mesh on the device, then graph + sparsity pattern (static_elasticity2D.c:45-50): built on the device
	// straight into the SELL arrays when the pattern qualifies (pattern_dev.cu), else on the host (pattern.cu)
	double t0 = now_ms();
	if (st == NBGPU_OK)
		st = nbgpu_mesh_create(md->N_nod, md->nod, md->N_elems, md->nodes_per_elem, md->adj, &S->mesh);
	S->ms_upload = now_ms() - t0;
	t0 = now_ms();
	if (st == NBGPU_OK)
		st = nbgpu_matrix_create_from_mesh(S->mesh, md->N_edg, md->edg, &S->K);
	if (st == NBGPU_OK && !S->K) {
		std::vector<uint32_t> rows_size(S->N), cols;
		st = nbgpu_pattern_from_mesh(md->N_nod, md->N_elems, md->nodes_per_elem, md->adj, md->N_edg, md->edg, 2,
					     rows_size.data(), nullptr, &S->nnz);
		if (st == NBGPU_OK) {
			cols.resize(S->nnz);
			st = nbgpu_pattern_from_mesh(md->N_nod, md->N_elems, md->nodes_per_elem, md->adj, md->N_edg, md->edg, 2,
						     rows_size.data(), cols.data(), &S->nnz);
		}
		if (st == NBGPU_OK)
			st = nbgpu_matrix_create_from_csr(S->N, rows_size.data(), cols.data(), nullptr, &S->K);
	}
	if (st == NBGPU_OK)
		st = nbgpu_matrix_info(S->K, nullptr, &S->nnz, nullptr, nullptr);
	S->ms_pattern = now_ms() - t0;

	// (2) the other device objects
	t0 = now_ms();
	const size_t n_strain = (size_t)3 * S->n_gp * md->N_elems;
	if (st == NBGPU_OK)
		st = nbgpu_malloc((void **)&S->d_vec, (2 * (size_t)S->N + n_strain) * sizeof(double));
	if (st == NBGPU_OK && n_neu) {
		st = nbgpu_malloc((void **)&S->d_neu_add, (size_t)n_neu * (sizeof(double) + sizeof(uint32_t)));
		if (st == NBGPU_OK) {
			S->d_neu_dof = (uint32_t *)(S->d_neu_add + n_neu);
			st = nbgpu_copy_h2d(S->d_neu_add, neu_add, (size_t)n_neu * sizeof(double));
		}
		if (st == NBGPU_OK)
			st = nbgpu_copy_h2d(S->d_neu_dof, neu_dof, (size_t)n_neu * sizeof(uint32_t));
	}
	if (st == NBGPU_OK)
		st = nbgpu_dirichlet_create(S->N, n_dir, dir_dof, dir_val, &S->dirichlet);
	if (st == NBGPU_OK) {
		S->d_F = S->d_vec;
		S->d_x = S->d_vec + S->N;
		S->d_strain = S->d_vec + 2 * (size_t)S->N;
		st = nbgpu_sync();
	}
	S->ms_upload += now_ms() - t0;
	if (st != NBGPU_OK) {
		nbgpu_fem_session_destroy(S);
		return st;
	}
	*out = S;
	return NBGPU_OK;
}

int nbgpu_fem_session_step(nbgpu_fem_session_t *S, const uint8_t *enabled, const double *elem_scale, int warm_start,
			   uint32_t max_iter, double solver_tol, nbgpu_fem_report_t *report)
{
	NB_INIT();
	NB_ARG(S != nullptr);
	nbgpu_fem_report_t rep;
	memset(&rep, 0, sizeof(rep));
	rep.N = S->N;
	rep.nnz = S->nnz;
	rep.ms_pattern = S->ms_pattern;
	rep.ms_upload = S->ms_upload;
	// (3) assembly (static_elasticity2D.c:58-65)
	double t0 = now_ms();
	int status = 0;
	int st = nbgpu_assemble_elasticity2d(S->K, S->mesh, &S->tables, &S->ap, enabled, elem_scale, S->d_F, nullptr);
	if (st == NBGPU_DISTORTED_ELEMENT) {
		status = 1;                   // static_elasticity2D.c:62-65
		st = NBGPU_OK;
	}
	if (st == NBGPU_OK)
		st = nbgpu_sync();
	rep.ms_assembly = now_ms() - t0;
	// (4) boundary conditions (static_elasticity2D.c:67)
	if (st == NBGPU_OK && status == 0) {
		t0 = now_ms();
		st = nbgpu_vector_add_entries_dev(S->d_F, S->n_neu, S->d_neu_dof, S->d_neu_add);
		if (st == NBGPU_OK)
			st = nbgpu_dirichlet_apply(S->K, S->d_F, S->dirichlet);
		if (st == NBGPU_OK)
			st = nbgpu_sync();
		rep.ms_bcond = now_ms() - t0;
	}
	// (5) solver(): x0 = 0 (or the previous displacement), abs tol; 0 and 1 both accepted (:83-97)
	if (st == NBGPU_OK && status == 0) {
		t0 = now_ms();
		if (!warm_start || !S->have_x)
			st = nbgpu_memset(S->d_x, 0, (size_t)S->N * sizeof(double));
		if (st == NBGPU_OK) {
			int sst = nbgpu_pcg_jacobi(S->K, S->d_F, S->d_x, max_iter ? max_iter : S->N,
						   solver_tol > 0 ? solver_tol : 1e-8, &rep.solver_iters, &rep.solver_residual);
			if (sst == NBGPU_OK || sst == NBGPU_NOT_CONVERGED) {
				rep.solver_status = sst;
				S->have_x = true;
			} else {
				st = sst;
			}
		}
		rep.ms_solve = now_ms() - t0;
	}
	if (report)
		*report = rep;
	return st != NBGPU_OK ? st : status;
}

/* displacement[N] and, if not NULL, strain[3 N_gp N_elems] of the last step (pipeline.c:266-319) */
int nbgpu_fem_session_results(nbgpu_fem_session_t *S, double *displacement, double *strain)
{
	NB_INIT();
	NB_ARG(S != nullptr && S->have_x);
	int st = NBGPU_OK;
	if (strain)
		st = nbgpu_compute_strain(S->mesh, &S->tables, S->d_x, S->d_strain);
	if (st == NBGPU_OK && displacement)
		st = nbgpu_copy_d2h(displacement, S->d_x, (size_t)S->N * sizeof(double));
	if (st == NBGPU_OK && strain)
		st = nbgpu_copy_d2h(strain, S->d_strain, (size_t)3 * S->n_gp * S->N_elems * sizeof(double));
	return st;
}

int nbgpu_fem_static_elasticity2d_lists(const nbgpu_mesh_desc_t *md, const nbgpu_elem_tables_t *tables,
					const double D[4], double density, uint32_t n_neu, const uint32_t *neu_dof,
					const double *neu_add, uint32_t n_dir, const uint32_t *dir_dof,
					const double *dir_val, int self_weight, const double gravity[2],
					int analysis2D, double thickness, const uint8_t *enabled,
					int assembly_mode, double solver_tol, double *displacement,
					double *strain, nbgpu_fem_report_t *report)
{
	NB_ARG(displacement != nullptr);
	(void)analysis2D;   // D already carries the (always plane-stress, formulas.c:38-45) material law
	nbgpu_fem_session_t *S = nullptr;
	NB_TRY(nbgpu_fem_session_create(md, tables, D, density, n_neu, neu_dof, neu_add, n_dir, dir_dof, dir_val,
					self_weight, gravity, thickness, assembly_mode, &S));
	nbgpu_fem_report_t rep;
	memset(&rep, 0, sizeof(rep));
	int st = nbgpu_fem_session_step(S, enabled, nullptr, 0, 0, solver_tol, &rep);
	if (st == NBGPU_OK) {
		// (6) strain at the Gauss points (:75) and results back to the caller
		const double t0 = now_ms();
		st = nbgpu_fem_session_results(S, displacement, strain);
		rep.ms_post = now_ms() - t0;
	}
	nbgpu_fem_session_destroy(S);
	if (report)
		*report = rep;
	return st;
}

}  // extern "C"
