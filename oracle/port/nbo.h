/*
 * oracle/port/nbo.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement ("port") of the reference algorithms on the hot
 * path, over flat CSR arrays instead of the reference's jagged nb_sparse_s.
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  The port is PINNED: tests/test_oracle_*.py check it
 * (a) against the unmodified reference compiled into oracle/_ref (when that
 * library is present) and (b) against the golden fixtures under tests/golden
 * that were generated from the reference by oracle/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may call this library.
 */
#ifndef NBO_H
#define NBO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- graph + pattern ---------------------------------------------------- */

/* load_graph.c:230-328 (NB_NODES_LINKED_BY_ELEMS).  Two-pass: call with
 * adj_flat == NULL to get counts (N_adj filled) and the total, then again. */
uint64_t nbo_graph_nodes_by_elems(uint32_t N_nod, uint32_t N_edg,
				  const uint32_t *edg, uint32_t N_elems,
				  uint32_t npe, const uint32_t *adj,
				  uint32_t *N_adj, uint32_t *adj_flat);

/* sparse.c:20-60 with perm == NULL.  rows_size[N*vars]; cols may be NULL on
 * the counting pass.  Returns nnz. */
uint64_t nbo_sparse_pattern(uint32_t N_graph, const uint32_t *N_adj,
			    const uint32_t *adj_flat, uint32_t vars,
			    uint32_t *rows_size, uint32_t *cols);

/* ---- SpMV / Krylov ------------------------------------------------------ */

void nbo_row_ptr(uint32_t N, const uint32_t *rows_size, uint64_t *row_ptr);

/* sparse.c:405-414 */
void nbo_spmv(uint32_t N, const uint64_t *row_ptr, const uint32_t *cols,
	      const double *vals, const double *in, double *out,
	      uint32_t threads);

/* cg_precond_jacobi.c:13-90 */
int nbo_pcg_jacobi(uint32_t N, const uint64_t *row_ptr, const uint32_t *cols,
		   const double *vals, const double *b, double *x,
		   uint32_t max_iter, double tol, uint32_t *iters,
		   double *tol_reached, uint32_t threads);

/* conjugate_gradient.c:13-77 */
int nbo_cg(uint32_t N, const uint64_t *row_ptr, const uint32_t *cols,
	   const double *vals, const double *b, double *x,
	   uint32_t max_iter, double tol, uint32_t *iters,
	   double *tol_reached, uint32_t threads);

/* sparse.c:416-430 */
void nbo_dirichlet(const uint64_t *row_ptr, const uint32_t *cols,
		   double *vals, double *rhs, uint32_t idx, double value);

/* ---- FEM ---------------------------------------------------------------- */

/* element.c:54-121 : elem_type 0 = 3-node triangle, 1 = 4-node quad.
 * Tables are indexed [node * N_gp + gp] (element.c:144-160). */
void nbo_elem_tables(int elem_type, uint32_t *N_nodes, uint32_t *N_gp,
		     double *w, double *Ni, double *dpsi, double *deta);

/* formulas.c:32-63 including the switch fall-through (plane stress always) */
void nbo_constitutive(double E, double nu, int analysis, double D[4]);

/* pipeline.c:42-264 + utils.c:9-60.  Returns 0, or 1 at the first element
 * with detJ < 0 (K/F then hold the contributions of the elements before it,
 * like the reference). */
int nbo_assemble(uint32_t N_nod, const double *nod, uint32_t N_elems,
		 int elem_type, const uint32_t *adj, double E, double nu,
		 double density, int self_weight, double gx, double gy,
		 int analysis, double thickness, const uint8_t *enabled,
		 const uint64_t *row_ptr, const uint32_t *cols, double *vals,
		 double *F);

/* the damage driver's assembly loop (static_damage2D.c:474-569): D of Gauss
 * point j of element k times (1 - gp_damage[k N_gp + j]) */
int nbo_assemble_damage(uint32_t N_nod, const double *nod, uint32_t N_elems,
			int elem_type, const uint32_t *adj, double E, double nu,
			double density, int self_weight, double gx, double gy,
			int analysis, double thickness, const uint8_t *enabled,
			const double *gp_damage,
			const uint64_t *row_ptr, const uint32_t *cols, double *vals,
			double *F);

/* the lumped mass vector pipeline_assemble_system fills when M != NULL
 * (pipeline.c:56-57, :216-222, :256-259); 0, or 1 at the first distorted element */
int nbo_lumped_mass(uint32_t N_nod, const double *nod, uint32_t N_elems,
		    int elem_type, const uint32_t *adj, double density,
		    double thickness, const uint8_t *enabled, double *M);

/* nbo_assemble with the product's per-element stiffness factor (not a reference
 * feature; checker for the SIMP-style hook only) */
int nbo_assemble_scaled(uint32_t N_nod, const double *nod, uint32_t N_elems,
			int elem_type, const uint32_t *adj, double E, double nu,
			double density, int self_weight, double gx, double gy,
			int analysis, double thickness, const uint8_t *enabled,
			const double *elem_scale,
			const uint64_t *row_ptr, const uint32_t *cols, double *vals,
			double *F);

/* Boundary-condition record, one per nb_bcond_push (bcond.c:153-165) */
typedef struct {
	int32_t kind;      /* 0 Dirichlet, 1 Neumann                         */
	int32_t where;     /* 0 input vertex, 1 input segment                */
	uint32_t id;       /* vertex / segment id                            */
	int32_t mask[2];
	int32_t fn;        /* 0 constant; 1/2 Kirsch traction on +x/+y face  */
	double val[2];
} nbo_bc_t;

/* set_bconditions.c:52-61 (order: Neumann sgm, Neumann vtx, Dirichlet sgm,
 * Dirichlet vtx; each queue in push order). */
void nbo_set_bconditions(const double *nod, const uint32_t *vtx,
			 const uint32_t *sgm_sizes, const uint32_t *sgm_nodes,
			 uint32_t N_bc, const nbo_bc_t *bc, double factor,
			 const uint64_t *row_ptr, const uint32_t *cols,
			 double *vals, double *F);

/* pipeline.c:266-319 */
int nbo_compute_strain(const double *nod, uint32_t N_elems, int elem_type,
		       const uint32_t *adj, const double *disp,
		       double *strain);

/* static_elasticity2D.c:99-127 */
void nbo_stress_from_strain(uint32_t N_elems, int elem_type, double E,
			    double nu, int analysis, const double *strain,
			    const uint8_t *enabled, double *stress);

void nbo_kirsch_stress(double x, double y, double s[3]);

/* gaussp_to_nodes.c:50-218 */
int nbo_gp_to_nodes(uint32_t N_nod, const double *nod, uint32_t N_elems,
		    int elem_type, const uint32_t *adj, uint32_t N_comp,
		    const double *gp_values, double *nodal_values);

/* formulas.c:65-77 */
double nbo_vm_stress(double sxx, double syy, double sxy);
void nbo_main_stress(double sxx, double syy, double sxy, double main_stress[2]);

#ifdef __cplusplus
}
#endif
#endif
