/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Flat (ctypes-friendly) entry points around the UNMODIFIED reference library.
 * It is compiled only by oracle/Makefile, together with the reference sources
 * where they lie under /root/reference, into oracle/_ref/libnbots_ref.so.
 * Nothing from the reference is copied here: the harness only CALLS the
 * reference through its public headers plus three private headers it includes
 * from the reference tree at build time (struct nb_sparse_s, struct
 * nb_mshquad_s / nb_msh3trg_s, pipeline.h / set_bconditions.h).
 *
 * Used for: pinning the C restatement in oracle/port, generating the golden
 * fixtures under tests/golden, and the `cpu_baseline.kind = "reference"` leg
 * of bench.py.  The product (nbots_b200/) never links or loads this.
 */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include <stdbool.h>
#include <math.h>

#include "nb/memory_bot.h"
#include "nb/graph_bot.h"
#include "nb/solver_bot.h"
#include "nb/geometric_bot.h"
#include "nb/pde_bot.h"

/* private headers of the reference, resolved with -I at build time */
#include "sparse_struct.h"                  /* sources/nb/solver_bot/sparse */
#include "mshquad_struct.h"                 /* .../mesh2D/elements2D        */
#include "msh3trg_struct.h"
#include "mesh2D_struct.h"                  /* sizeof(struct nb_mesh2D_s)   */
#include "pipeline.h"                       /* FEM solid_mechanics          */
#include "set_bconditions.h"

/* --------------------------------------------------------------- meshes -- */

typedef struct {
	nb_mesh2D_t *mesh;       /* vtable wrapper                          */
	void *msh;               /* struct nb_mshquad_s / nb_msh3trg_s      */
	int kind;                /* 0 trg, 1 quad                           */
	int owns_arrays;         /* built from flat arrays by this harness  */
} refh_mesh_t;

static void *dup_mem(const void *src, size_t bytes)
{
	void *p = malloc(bytes ? bytes : 1);
	if (bytes)
		memcpy(p, src, bytes);
	return p;
}

/* Wrap flat arrays into the reference's own mesh structs
 * (mshquad_struct.h:6-26, msh3trg_struct.h:6-23) and its vtable
 * (nb_mesh2D_init_from_msh, mesh2D.c:89-95). */
void *refh_mesh_create(int kind, uint32_t N_nod, const double *nod,
		       uint32_t N_edg, const uint32_t *edg,
		       uint32_t N_elems, const uint32_t *adj,
		       uint32_t N_vtx, const uint32_t *vtx,
		       uint32_t N_sgm, const uint32_t *N_nod_x_sgm,
		       const uint32_t *nod_x_sgm_flat)
{
	refh_mesh_t *h = calloc(1, sizeof(*h));
	h->kind = kind;
	h->owns_arrays = 1;
	uint32_t npe = kind ? 4 : 3;
	uint32_t **sgm = malloc((N_sgm ? N_sgm : 1) * sizeof(*sgm));
	uint32_t off = 0;
	for (uint32_t s = 0; s < N_sgm; s++) {
		sgm[s] = dup_mem(nod_x_sgm_flat + off,
				 N_nod_x_sgm[s] * sizeof(uint32_t));
		off += N_nod_x_sgm[s];
	}
	uint32_t *ngb = malloc((size_t)npe * N_elems * sizeof(uint32_t) + 4);
	for (size_t i = 0; i < (size_t)npe * N_elems; i++)
		ngb[i] = N_elems;   /* "no neighbour"; unused on the FEM path */
	if (kind) {
		struct nb_mshquad_s *q = calloc(1, sizeof(*q));
		q->N_nod = N_nod;
		q->nod = dup_mem(nod, 2 * (size_t)N_nod * sizeof(double));
		q->N_edg = N_edg;
		q->edg = dup_mem(edg, 2 * (size_t)N_edg * sizeof(uint32_t));
		q->N_elems = N_elems;
		q->type = calloc(N_elems ? N_elems : 1, 1);
		q->adj = dup_mem(adj, 4 * (size_t)N_elems * sizeof(uint32_t));
		q->ngb = ngb;
		q->N_vtx = N_vtx;
		q->vtx = dup_mem(vtx, N_vtx * sizeof(uint32_t));
		q->N_sgm = N_sgm;
		q->N_nod_x_sgm = dup_mem(N_nod_x_sgm, N_sgm * sizeof(uint32_t));
		q->nod_x_sgm = sgm;
		h->msh = q;
	} else {
		struct nb_msh3trg_s *t = calloc(1, sizeof(*t));
		t->N_nod = N_nod;
		t->nod = dup_mem(nod, 2 * (size_t)N_nod * sizeof(double));
		t->N_edg = N_edg;
		t->edg = dup_mem(edg, 2 * (size_t)N_edg * sizeof(uint32_t));
		t->N_elems = N_elems;
		t->adj = dup_mem(adj, 3 * (size_t)N_elems * sizeof(uint32_t));
		t->ngb = ngb;
		t->N_vtx = N_vtx;
		t->vtx = dup_mem(vtx, N_vtx * sizeof(uint32_t));
		t->N_sgm = N_sgm;
		t->N_nod_x_sgm = dup_mem(N_nod_x_sgm, N_sgm * sizeof(uint32_t));
		t->nod_x_sgm = sgm;
		h->msh = t;
	}
	h->mesh = calloc(1, sizeof(struct nb_mesh2D_s));
	nb_mesh2D_init_from_msh(h->mesh, h->msh, kind ? NB_QUAD : NB_TRIAN);
	return h;
}

/* Triangle mesh produced by the reference's own mesher from a PSLG model,
 * the way the reference's FEM test does (utest/.../static_elasticity2D.c,
 * get_mesh): MAX_VTX size constraint + MAX_EDGE_LENGTH = NB_GEOMETRIC_TOL. */
void *refh_mesh_from_model(uint32_t N, const double *vertex,
			   uint32_t M, const uint32_t *edge,
			   uint32_t H, const double *holes, uint32_t max_vtx)
{
	nb_model_t *model = nb_model_create();
	model->N = N;
	model->vertex = dup_mem(vertex, 2 * N * sizeof(double));
	model->M = M;
	model->edge = dup_mem(edge, 2 * M * sizeof(uint32_t));
	model->H = H;
	model->holes = H ? dup_mem(holes, 2 * H * sizeof(double)) : NULL;

	nb_tessellator2D_t *t2d = nb_tessellator2D_create();
	nb_tessellator2D_set_size_constraint(t2d,
			NB_MESH_SIZE_CONSTRAINT_MAX_VTX, max_vtx);
	nb_tessellator2D_set_geometric_constraint(t2d,
			NB_MESH_GEOM_CONSTRAINT_MAX_EDGE_LENGTH,
			NB_GEOMETRIC_TOL);
	nb_tessellator2D_generate_from_model(t2d, model);

	refh_mesh_t *h = calloc(1, sizeof(*h));
	h->kind = 0;
	h->owns_arrays = 0;
	h->mesh = nb_mesh2D_create(NB_TRIAN);
	nb_mesh2D_load_from_tessellator2D(h->mesh, t2d);
	nb_tessellator2D_destroy(t2d);
	nb_model_destroy(model);
	return h;
}

/* counts[0..5] = N_nod, N_edg, N_elems, N_vtx, N_sgm, total nodes in sgm */
void refh_mesh_counts(const void *hp, uint32_t counts[6])
{
	const refh_mesh_t *h = hp;
	const nb_mesh2D_t *m = h->mesh;
	counts[0] = nb_mesh2D_get_N_nodes(m);
	counts[1] = nb_mesh2D_get_N_edges(m);
	counts[2] = nb_mesh2D_get_N_elems(m);
	counts[3] = nb_mesh2D_get_N_invtx(m);
	counts[4] = nb_mesh2D_get_N_insgm(m);
	uint32_t tot = 0;
	for (uint32_t s = 0; s < counts[4]; s++)
		tot += nb_mesh2D_insgm_get_N_nodes(m, s);
	counts[5] = tot;
}

void refh_mesh_export(const void *hp, double *nod, uint32_t *edg,
		      uint32_t *adj, uint32_t *vtx, uint32_t *N_nod_x_sgm,
		      uint32_t *nod_x_sgm_flat)
{
	const refh_mesh_t *h = hp;
	const nb_mesh2D_t *m = h->mesh;
	uint32_t npe = h->kind ? 4 : 3;
	uint32_t N_nod = nb_mesh2D_get_N_nodes(m);
	for (uint32_t i = 0; i < N_nod; i++) {
		nod[2 * i] = nb_mesh2D_node_get_x(m, i);
		nod[2 * i + 1] = nb_mesh2D_node_get_y(m, i);
	}
	uint32_t N_edg = nb_mesh2D_get_N_edges(m);
	for (uint32_t i = 0; i < N_edg; i++) {
		edg[2 * i] = nb_mesh2D_edge_get_1n(m, i);
		edg[2 * i + 1] = nb_mesh2D_edge_get_2n(m, i);
	}
	uint32_t N_el = nb_mesh2D_get_N_elems(m);
	for (uint32_t i = 0; i < N_el; i++)
		for (uint32_t j = 0; j < npe; j++)
			adj[npe * i + j] = nb_mesh2D_elem_get_adj(m, i, j);
	uint32_t N_vtx = nb_mesh2D_get_N_invtx(m);
	for (uint32_t i = 0; i < N_vtx; i++)
		vtx[i] = nb_mesh2D_get_invtx(m, i);
	uint32_t N_sgm = nb_mesh2D_get_N_insgm(m);
	uint32_t off = 0;
	for (uint32_t s = 0; s < N_sgm; s++) {
		N_nod_x_sgm[s] = nb_mesh2D_insgm_get_N_nodes(m, s);
		for (uint32_t i = 0; i < N_nod_x_sgm[s]; i++)
			nod_x_sgm_flat[off++] =
				nb_mesh2D_insgm_get_node(m, s, i);
	}
}

void refh_mesh_destroy(void *hp)
{
	refh_mesh_t *h = hp;
	if (!h->owns_arrays) {
		nb_mesh2D_destroy(h->mesh);
		free(h);
		return;
	}
	if (h->kind) {
		struct nb_mshquad_s *q = h->msh;
		for (uint32_t s = 0; s < q->N_sgm; s++)
			free(q->nod_x_sgm[s]);
		free(q->nod_x_sgm); free(q->N_nod_x_sgm); free(q->vtx);
		free(q->ngb); free(q->adj); free(q->type); free(q->edg);
		free(q->nod); free(q);
	} else {
		struct nb_msh3trg_s *t = h->msh;
		for (uint32_t s = 0; s < t->N_sgm; s++)
			free(t->nod_x_sgm[s]);
		free(t->nod_x_sgm); free(t->N_nod_x_sgm); free(t->vtx);
		free(t->ngb); free(t->adj); free(t->edg);
		free(t->nod); free(t);
	}
	free(h->mesh);
	free(h);
}

/* --------------------------------------------------------------- sparse -- */

/* static_elasticity2D.c:45-50 : graph by elems -> nb_sparse_create(,NULL,2) */
void *refh_sparse_from_mesh(const void *hp)
{
	const refh_mesh_t *h = hp;
	nb_graph_t *graph = malloc(nb_graph_get_memsize());
	nb_graph_init(graph);
	nb_mesh2D_load_graph(h->mesh, graph, NB_NODES_LINKED_BY_ELEMS);
	nb_sparse_t *K = nb_sparse_create(graph, NULL, 2);
	nb_graph_finish(graph);
	free(graph);
	return K;
}

/* number of directed adjacencies of the NODES_LINKED_BY_ELEMS graph */
uint64_t refh_graph_from_mesh(const void *hp, uint32_t *N_adj,
			      uint32_t *adj_flat /* NULL = count only */)
{
	const refh_mesh_t *h = hp;
	nb_graph_t *graph = malloc(nb_graph_get_memsize());
	nb_graph_init(graph);
	nb_mesh2D_load_graph(h->mesh, graph, NB_NODES_LINKED_BY_ELEMS);
	uint64_t tot = 0;
	for (uint32_t i = 0; i < graph->N; i++) {
		if (N_adj)
			N_adj[i] = graph->N_adj[i];
		for (uint32_t j = 0; j < graph->N_adj[i]; j++) {
			if (adj_flat)
				adj_flat[tot] = graph->adj[i][j];
			tot++;
		}
	}
	nb_graph_finish(graph);
	free(graph);
	return tot;
}

void *refh_sparse_from_graph(uint32_t N, const uint32_t *N_adj,
			     const uint32_t *adj_flat, uint32_t vars_per_node)
{
	nb_graph_t g;
	g.N = N;
	g.N_adj = (uint32_t *)N_adj;
	g.adj = malloc((size_t)(N ? N : 1) * sizeof(*g.adj));
	g.wi = NULL;
	g.wij = NULL;
	uint64_t off = 0;
	for (uint32_t i = 0; i < N; i++) {
		g.adj[i] = (uint32_t *)adj_flat + off;
		off += N_adj[i];
	}
	nb_sparse_t *A = nb_sparse_create(&g, NULL, vars_per_node);
	free(g.adj);
	return A;
}

uint32_t refh_sparse_N(const void *A) { return ((const nb_sparse_t *)A)->N; }

uint64_t refh_sparse_nnz(const void *Ap)
{
	const nb_sparse_t *A = Ap;
	uint64_t nnz = 0;
	for (uint32_t i = 0; i < A->N; i++)
		nnz += A->rows_size[i];
	return nnz;
}

void refh_sparse_export(const void *Ap, uint32_t *rows_size,
			uint32_t *cols, double *vals)
{
	const nb_sparse_t *A = Ap;
	uint64_t off = 0;
	for (uint32_t i = 0; i < A->N; i++) {
		uint32_t n = A->rows_size[i];
		if (rows_size)
			rows_size[i] = n;
		if (cols)
			memcpy(cols + off, A->rows_index[i], n * sizeof(uint32_t));
		if (vals)
			memcpy(vals + off, A->rows_values[i], n * sizeof(double));
		off += n;
	}
}

void refh_sparse_import_values(void *Ap, const double *vals)
{
	nb_sparse_t *A = Ap;
	uint64_t off = 0;
	for (uint32_t i = 0; i < A->N; i++) {
		uint32_t n = A->rows_size[i];
		memcpy(A->rows_values[i], vals + off, n * sizeof(double));
		off += n;
	}
}

/* raw struct view, for handing a genuine nb_sparse_t to the drop-in shims */
void refh_sparse_raw(const void *Ap, double ***rows_values,
		     uint32_t ***rows_index, uint32_t **rows_size, uint32_t *N)
{
	const nb_sparse_t *A = Ap;
	*rows_values = A->rows_values;
	*rows_index = A->rows_index;
	*rows_size = A->rows_size;
	*N = A->N;
}

void refh_sparse_destroy(void *A) { nb_sparse_destroy(A); }

void refh_spmv(const void *A, const double *in, double *out, uint32_t threads)
{
	nb_sparse_multiply_vector(A, in, out, threads);
}

int refh_pcg_jacobi(const void *A, const double *b, double *x,
		    uint32_t max_iter, double tol, uint32_t *iters,
		    double *tol_reached, uint32_t threads)
{
	return nb_sparse_solve_CG_precond_Jacobi(A, b, x, max_iter, tol,
						 iters, tol_reached, threads);
}

int refh_cg(const void *A, const double *b, double *x,
	    uint32_t max_iter, double tol, uint32_t *iters,
	    double *tol_reached, uint32_t threads)
{
	return nb_sparse_solve_conjugate_gradient(A, b, x, max_iter, tol,
						  iters, tol_reached, threads);
}

/* the reference's OTHER caller of the two Krylov entry points that can be run as it is:
 * inverse power iteration (eigen/inv_power.c:75 Jacobi-PCG, :80 plain CG).  The model regulariser
 * (geometric_bot/model/modules2D/regularizer.c:24-57) cannot: nb_model_load_vtx_graph (model2D.c:276,288)
 * hands the container ARRAY to nb_container_init / _finish instead of its i-th element and crashes before
 * the solver is reached; tests/test_gpu_dropin.py restates its system (regularizer.c:74-108) instead. */
int refh_inv_power(const void *A, int use_jacobi, int h, double mu, double *eigenvecs /* [h][N] */,
		   double *eigenvals, int *it, double tolerance, uint32_t threads)
{
	const uint32_t N = ((const nb_sparse_t *)A)->N;
	double **vecs = malloc((size_t)h * sizeof(*vecs));
	for (int i = 0; i < h; i++)
		vecs[i] = eigenvecs + (size_t)i * N;
	/* any value other than the two direct solvers and CGJ selects plain CG (inv_power.c:79-83) */
	int st = nb_sparse_eigen_ipower(A, use_jacobi ? NB_SOLVER_CGJ : (nb_solver_t)99, h, mu, vecs, eigenvals,
					it, tolerance, threads);
	free(vecs);
	return st;
}

void refh_dirichlet(void *A, double *rhs, uint32_t idx, double value)
{
	nb_sparse_set_Dirichlet_condition(A, rhs, idx, value);
}

/* ------------------------------------------------------------------ FEM -- */

static nb_material_t *make_material(double E, double nu, double density)
{
	nb_material_t *mat = nb_material_create();
	nb_material_set_elasticity_module(mat, E);
	nb_material_set_poisson_module(mat, nu);
	nb_material_set_density(mat, density);
	return mat;
}

void refh_elem_tables(int elem_type, uint32_t *N_nodes, uint32_t *N_gp,
		      double *w, double *Ni, double *dpsi, double *deta)
{
	nb_fem_elem_t *e = nb_fem_elem_create(elem_type ? NB_QUAD_LINEAR
							: NB_TRG_LINEAR);
	uint32_t n = nb_fem_elem_get_N_nodes(e);
	uint32_t g = nb_fem_elem_get_N_gpoints(e);
	*N_nodes = n;
	*N_gp = g;
	for (uint32_t p = 0; p < g; p++)
		w[p] = nb_fem_elem_weight_gp(e, p);
	for (uint32_t i = 0; i < n; i++)
		for (uint32_t p = 0; p < g; p++) {
			Ni[i * g + p] = nb_fem_elem_Ni(e, i, p);
			dpsi[i * g + p] = nb_fem_elem_dNi_dpsi(e, i, p);
			deta[i * g + p] = nb_fem_elem_dNi_deta(e, i, p);
		}
	nb_fem_elem_destroy(e);
}

void refh_constitutive(double E, double nu, int analysis, double D[4])
{
	nb_material_t *mat = make_material(E, nu, 0.0);
	nb_pde_get_constitutive_matrix(D, mat, (nb_analysis2D_t)analysis);
	nb_material_destroy(mat);
}

static int assemble_with(void *K, double *M, double *F, const void *hp,
			 int elem_type, double E, double nu, double density,
			 int self_weight, double gx, double gy, int analysis,
			 double thickness, const uint8_t *enabled);

int refh_assemble(void *K, double *F, const void *hp, int elem_type,
		  double E, double nu, double density, int self_weight,
		  double gx, double gy, int analysis, double thickness,
		  const uint8_t *enabled /* NULL = all */)
{
	return assemble_with(K, NULL, F, hp, elem_type, E, nu, density,
			     self_weight, gx, gy, analysis, thickness, enabled);
}

/* the same call with the lumped mass vector wanted (M: 2 N_nod doubles) */
int refh_assemble_mass(void *K, double *M, double *F, const void *hp,
		       int elem_type, double E, double nu, double density,
		       int self_weight, double gx, double gy, int analysis,
		       double thickness, const uint8_t *enabled)
{
	return assemble_with(K, M, F, hp, elem_type, E, nu, density,
			     self_weight, gx, gy, analysis, thickness, enabled);
}

static int assemble_with(void *K, double *M, double *F, const void *hp,
			 int elem_type, double E, double nu, double density,
			 int self_weight, double gx, double gy, int analysis,
			 double thickness, const uint8_t *enabled)
{
	const refh_mesh_t *h = hp;
	nb_fem_elem_t *e = nb_fem_elem_create(elem_type ? NB_QUAD_LINEAR
							: NB_TRG_LINEAR);
	nb_material_t *mat = make_material(E, nu, density);
	nb_analysis2D_params params;
	memset(&params, 0, sizeof(params));
	params.thickness = thickness;
	double gravity[2] = {gx, gy};
	uint32_t N_el = nb_mesh2D_get_N_elems(h->mesh);
	bool *en = NULL;
	if (enabled) {
		en = malloc(N_el * sizeof(bool) + 1);
		for (uint32_t i = 0; i < N_el; i++)
			en[i] = enabled[i] != 0;
	}
	int status = pipeline_assemble_system(K, M, F, h->mesh, e, mat,
					      self_weight != 0, gravity,
					      (nb_analysis2D_t)analysis,
					      &params, en);
	free(en);
	nb_material_destroy(mat);
	nb_fem_elem_destroy(e);
	return status;
}

/* static_damage2D.c:474-569, made visible by the Makefile (objcopy); the
 * prototype is the file's own (static_damage2D.c:52-63) */
void DMG_pipeline_assemble_system(nb_sparse_t *K, double *M, double *F,
				  const nb_mesh2D_t *const part,
				  const nb_fem_elem_t *const elem,
				  const nb_material_t *const material,
				  bool enable_self_weight, double gravity[2],
				  nb_analysis2D_t analysis2D,
				  nb_analysis2D_params *params2D,
				  bool enable_computing_damage,
				  double *damage_elem, bool *elements_enabled);

/* the damage driver's assembly: D of Gauss point j of element k is scaled by
 * (1 - damage[k * N_gp + j]); damage NULL = no scaling */
void refh_assemble_damage(void *K, double *F, const void *hp, int elem_type,
			  double E, double nu, double density, int self_weight,
			  double gx, double gy, int analysis, double thickness,
			  const double *damage, const uint8_t *enabled)
{
	const refh_mesh_t *h = hp;
	nb_fem_elem_t *e = nb_fem_elem_create(elem_type ? NB_QUAD_LINEAR
							: NB_TRG_LINEAR);
	nb_material_t *mat = make_material(E, nu, density);
	nb_analysis2D_params params;
	memset(&params, 0, sizeof(params));
	params.thickness = thickness;
	double gravity[2] = {gx, gy};
	uint32_t N_el = nb_mesh2D_get_N_elems(h->mesh);
	bool *en = NULL;
	if (enabled) {
		en = malloc(N_el * sizeof(bool) + 1);
		for (uint32_t i = 0; i < N_el; i++)
			en[i] = enabled[i] != 0;
	}
	DMG_pipeline_assemble_system(K, NULL, F, h->mesh, e, mat,
				     self_weight != 0, gravity,
				     (nb_analysis2D_t)analysis, &params,
				     damage != NULL, (double *)damage, en);
	free(en);
	nb_material_destroy(mat);
	nb_fem_elem_destroy(e);
}

void *refh_bcond_create(void) { return nb_bcond_create(2); }
void refh_bcond_destroy(void *bc) { nb_bcond_destroy(bc); }

/* type: 0 Dirichlet 1 Neumann ; where: 0 point(vtx) 1 segment */
void refh_bcond_push(void *bc, int type, int where, uint32_t id,
		     int mask_x, int mask_y, double vx, double vy)
{
	bool mask[2] = {mask_x != 0, mask_y != 0};
	double val[2] = {vx, vy};
	nb_bcond_push(bc, type ? NB_NEUMANN : NB_DIRICHLET,
		      where ? NB_BC_ON_SEGMENT : NB_BC_ON_POINT,
		      id, mask, val);
}

/* Kirsch plate-with-hole tractions (a = 0.5, far-field tx = 1e4), the
 * analytic field the reference FEM test pushes as function-valued Neumann
 * conditions on two segments.  Standard closed-form solution. */
static void kirsch(double x, double y, double s[3])
{
	double a = 0.5, tx = 1e4;
	double r2 = x * x + y * y;
	double th = atan2(y, x);
	double q = a * a / r2;
	double q2x = 1.5 * q * q;
	double c2 = cos(2 * th), c4 = cos(4 * th);
	double s2 = sin(2 * th), s4 = sin(4 * th);
	s[0] = tx * (1.0 - q * (1.5 * c2 + c4) + q2x * c4);
	s[1] = tx * (-q * (0.5 * c2 - c4) - q2x * c4);
	s[2] = tx * (-q * (0.5 * s2 + s4) + q2x * s4);
}
static void kirsch_face_x(const double *x, double t, double *out)
{
	double s[3];
	(void)t;
	kirsch(x[0], x[1], s);
	out[0] = s[0];
	out[1] = s[2];
}
static void kirsch_face_y(const double *x, double t, double *out)
{
	double s[3];
	(void)t;
	kirsch(x[0], x[1], s);
	out[0] = s[2];
	out[1] = s[1];
}
void refh_kirsch_stress(double x, double y, double s[3]) { kirsch(x, y, s); }

/* which: 0 -> traction on a face with normal +x, 1 -> normal +y */
void refh_bcond_push_kirsch(void *bc, uint32_t sgm_id, int which)
{
	bool mask[2] = {true, true};
	nb_bcond_push_function(bc, NB_NEUMANN, NB_BC_ON_SEGMENT, sgm_id, mask,
			       which ? kirsch_face_y : kirsch_face_x);
}

void refh_set_bconditions(const void *hp, void *K, double *F,
			  const void *bc, double factor)
{
	const refh_mesh_t *h = hp;
	nb_fem_set_bconditions(h->mesh, K, F, bc, factor);
}

int refh_fem_static(const void *hp, int elem_type, double E, double nu,
		    double density, const void *bc, int self_weight,
		    double gx, double gy, int analysis, double thickness,
		    const uint8_t *enabled, double *disp, double *strain)
{
	const refh_mesh_t *h = hp;
	nb_fem_elem_t *e = nb_fem_elem_create(elem_type ? NB_QUAD_LINEAR
							: NB_TRG_LINEAR);
	nb_material_t *mat = make_material(E, nu, density);
	nb_analysis2D_params params;
	memset(&params, 0, sizeof(params));
	params.thickness = thickness;
	double gravity[2] = {gx, gy};
	uint32_t N_el = nb_mesh2D_get_N_elems(h->mesh);
	bool *en = NULL;
	if (enabled) {
		en = malloc(N_el * sizeof(bool) + 1);
		for (uint32_t i = 0; i < N_el; i++)
			en[i] = enabled[i] != 0;
	}
	int status = nb_fem_compute_2D_Solid_Mechanics(h->mesh, e, mat, bc,
						       self_weight != 0,
						       gravity,
						       (nb_analysis2D_t)analysis,
						       &params, en, disp,
						       strain);
	free(en);
	nb_material_destroy(mat);
	nb_fem_elem_destroy(e);
	return status;
}

void refh_compute_strain(const void *hp, int elem_type, double *disp,
			 double *strain)
{
	const refh_mesh_t *h = hp;
	nb_fem_elem_t *e = nb_fem_elem_create(elem_type ? NB_QUAD_LINEAR
							: NB_TRG_LINEAR);
	pipeline_compute_strain(strain, h->mesh, disp, e);
	nb_fem_elem_destroy(e);
}

void refh_stress_from_strain(uint32_t N_elems, int elem_type, double E,
			     double nu, int analysis, double *strain,
			     const uint8_t *enabled, double *stress)
{
	nb_fem_elem_t *e = nb_fem_elem_create(elem_type ? NB_QUAD_LINEAR
							: NB_TRG_LINEAR);
	nb_material_t *mat = make_material(E, nu, 0.0);
	bool *en = NULL;
	if (enabled) {
		en = malloc(N_elems * sizeof(bool) + 1);
		for (uint32_t i = 0; i < N_elems; i++)
			en[i] = enabled[i] != 0;
	}
	nb_fem_compute_stress_from_strain(N_elems, e, mat,
					  (nb_analysis2D_t)analysis, strain,
					  en, stress);
	free(en);
	nb_material_destroy(mat);
	nb_fem_elem_destroy(e);
}

/* NBT mesh file format (mesh2D/file_format_nbt.c), called as it is */
int refh_mesh_save_nbt(const void *hp, const char *path)
{
	return nb_mesh2D_save_nbt(((const refh_mesh_t *)hp)->mesh, path);
}
int refh_mesh_read_type_nbt(const char *path, int *type)
{
	nb_mesh2D_type t = NB_TRIAN;
	int st = nb_mesh2D_read_type_nbt(path, &t);
	*type = (int)t;
	return st;
}
int refh_mesh_read_nbt(void *hp, const char *path)
{
	return nb_mesh2D_read_nbt(((refh_mesh_t *)hp)->mesh, path);
}

/* the reference mesh pointer itself, for the drop-in shim tests */
/* gaussp_to_nodes.c:50 and formulas.c:65-77, called as they are */
int refh_gp_to_nodes(const void *hp, int elem_type, uint32_t N_comp,
		     const double *gp_values, double *nodal_values)
{
	const refh_mesh_t *h = hp;
	nb_fem_elem_t *e = nb_fem_elem_create(elem_type ? NB_QUAD_LINEAR
							: NB_TRG_LINEAR);
	int st = nb_fem_interpolate_from_gpoints_to_nodes(h->mesh, e, N_comp,
							  gp_values,
							  nodal_values);
	nb_fem_elem_destroy(e);
	return st;
}

void refh_vm_stress(uint32_t n, const double *stress, double *vm)
{
	for (uint32_t i = 0; i < n; i++)
		vm[i] = nb_pde_get_vm_stress(stress[3 * i], stress[3 * i + 1],
					     stress[3 * i + 2]);
}

void refh_main_stress(uint32_t n, const double *stress, double *main_stress)
{
	for (uint32_t i = 0; i < n; i++)
		nb_pde_get_main_stress(stress[3 * i], stress[3 * i + 1],
				       stress[3 * i + 2], main_stress + 2 * i);
}

/* on-disk formats written by the reference itself (sparse.c:97-112,
 * matlab_v4.c:184-250 and :537-564, mesh2D_file_format_vtk.c:23-89) */
void refh_sparse_save(const void *A, const char *path) { nb_sparse_save(A, path); }

void refh_sparse_save_mat4(const void *A, const char *path, const char *label)
{
	char buf[64];
	snprintf(buf, sizeof(buf), "%s", label);
	nb_sparse_save_mat4(A, path, buf);
}

void refh_mat4_save_vec(const char *path, const char *label, const double *x,
			uint32_t N)
{
	char buf[64];
	snprintf(buf, sizeof(buf), "%s", label);
	nb_mat4_save_vec(path, buf, x, N);
}

int refh_mesh_save_vtk(const void *hp, const char *path)
{
	return nb_mesh2D_save_vtk(((const refh_mesh_t *)hp)->mesh, path);
}

void *refh_mesh_ptr(const void *hp) { return ((const refh_mesh_t *)hp)->mesh; }
