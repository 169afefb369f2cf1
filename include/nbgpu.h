/*
 * nbgpu.h -- C ABI of the B200 (sm_100a) implementation of the NBOTS hot path:
 * Jacobi-PCG / CG / SpMV over the Solver bot's nb_sparse_t and the PDE bot's
 * 2-D linear-elastic FEM assembly that feeds it.
 *
 * This header is the drop-in boundary.  Everything is
 * `extern "C"`, plain pointers and sizes; no CUDA or torch types appear in a
 * signature.  Each entry point names the reference interface it replaces
 * (paths relative to the reference tree).  The reference-named wrappers
 * (nb_sparse_solve_CG_precond_Jacobi, ... ) that a maintainer links in front
 * of libnbots live in nbots_b200/csrc/shim/ and are described in
 * INTEGRATION.md.
 *
 * Conventions kept from the reference (README_DEVELOPERS.md:52,82-90):
 *   - int return codes, 0 = success;
 *   - solver: 0 converged, 1 = max_iter reached (cg_precond_jacobi.c:86-89);
 *   - assembly: 0 ok, 1 = distorted element (pipeline.c:53-72, utils.c:44-47);
 *   - new failure classes use codes >= 10 (callers only test `0 != status`);
 *   - caller owns every buffer it passes; `x` is in (initial guess) / out.
 * Pointers named d_* are DEVICE pointers (from nbgpu_malloc); all others are
 * host pointers.  There is no CPU fallback: without a usable CUDA device every
 * compute call fails with NBGPU_ERR_CUDA.
 */
#ifndef NBGPU_H
#define NBGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBGPU_OK                 0
#define NBGPU_NOT_CONVERGED      1   /* solver: max_iter reached            */
#define NBGPU_DISTORTED_ELEMENT  1   /* assembly: detJ < 0 somewhere        */
#define NBGPU_ERR_CUDA          10   /* CUDA runtime / no device            */
#define NBGPU_ERR_ARG           11   /* invalid argument                    */
#define NBGPU_ERR_PATTERN       12   /* entry not in the sparsity pattern
                                        (reference: printf + exit(1),
                                        sparse.c:213-217)                   */
#define NBGPU_ERR_NOMEM         13
#define NBGPU_ERR_COMM          14   /* multi-GPU exchange failed           */

typedef struct nbgpu_matrix_s nbgpu_matrix_t;   /* device-resident nb_sparse_t */
typedef struct nbgpu_mesh_s nbgpu_mesh_t;       /* device-resident FEM mesh    */

/* ------------------------------------------------------------ context -- */

/* Bind this process to `device` (-1: LOCAL_RANK env, else 0) and create the
 * work stream.  Idempotent; compute calls do it lazily with -1. */
int nbgpu_init(int device);
int nbgpu_finalize(void);
int nbgpu_device_count(void);
int nbgpu_sync(void);
/* text of the last error raised on the calling thread ("" if none) */
const char *nbgpu_last_error(void);
/* the cudaStream_t every kernel of this library is launched on */
void *nbgpu_stream(void);
/* one-line JSON description of the bound device (SMs, L2, carve-outs, HBM) */
int nbgpu_device_info(char *buf, size_t len);
/* number of kernels this library has launched so far (for `gpu_launches`) */
uint64_t nbgpu_launch_count(void);
/* Give the CALLING host thread a context of its own on `device` (its stream, pool,
 * workspace): the single-process multi-GPU mode runs one such thread per GPU.
 * device < 0 returns the thread to the process-wide context. */
int nbgpu_thread_bind_device(int device);

int nbgpu_malloc(void **d_ptr, size_t bytes);
int nbgpu_free(void *d_ptr);
int nbgpu_memset(void *d_ptr, int byte, size_t bytes);
int nbgpu_copy_h2d(void *d_dst, const void *src, size_t bytes);
int nbgpu_copy_d2h(void *dst, const void *d_src, size_t bytes);
int nbgpu_copy_d2d(void *d_dst, const void *d_src, size_t bytes);
/* pinned host memory (H2D/D2H at full PCIe rate) */
int nbgpu_host_alloc(void **ptr, size_t bytes);
int nbgpu_host_free(void *ptr);
/* stream-ordered timing with CUDA events on the library's stream */
int nbgpu_timer_start(void);
int nbgpu_timer_stop(float *elapsed_ms);

/* ------------------------------------------------------------- matrix -- */
/* Replaces the storage of struct nb_sparse_s
 * (sources/nb/solver_bot/sparse/sparse_struct.h:6-11): per-row heap blocks
 * become one SELL-32 (sliced ELLPACK, slice = one warp of rows) block in HBM.
 * The pattern (rows_size, ascending unique columns) is kept bit for bit. */

/* from the jagged arrays of an nb_sparse_t (what nb_sparse_create returns,
 * sparse.c:20-60); rows_values == NULL means all-zero values */
int nbgpu_matrix_create_from_rows(uint32_t N, const uint32_t *rows_size,
				  uint32_t *const *rows_index,
				  double *const *rows_values,
				  nbgpu_matrix_t **out);
/* from flat CSR arrays (rows concatenated); vals == NULL means zeros */
int nbgpu_matrix_create_from_csr(uint32_t N, const uint32_t *rows_size,
				 const uint32_t *cols, const double *vals,
				 nbgpu_matrix_t **out);
int nbgpu_matrix_destroy(nbgpu_matrix_t *A);
/* N, nnz, slices, stored (padded) entries */
int nbgpu_matrix_info(const nbgpu_matrix_t *A, uint32_t *N, uint64_t *nnz,
		      uint32_t *n_slices, uint64_t *stored_entries);
/* how the rows are stored: sigma = sorting window of the SELL-32-sigma layout
 * (1 = rows in natural order), uniform_width != 0 when every slice has that
 * width, blocked != 0 when the 2x2 node-block column ids are in use, idx16 != 0
 * when the kernels read 16-bit column differences instead of 32-bit ids */
int nbgpu_matrix_layout(const nbgpu_matrix_t *A, uint32_t *sigma,
			uint32_t *uniform_width, uint32_t *max_width,
			int *blocked, int *idx16);
int nbgpu_matrix_set_values_rows(nbgpu_matrix_t *A, double *const *rows_values);
int nbgpu_matrix_set_values_csr(nbgpu_matrix_t *A, const double *vals);
int nbgpu_matrix_get_values_rows(const nbgpu_matrix_t *A, double *const *rows_values);
int nbgpu_matrix_get_values_csr(const nbgpu_matrix_t *A, double *vals);
/* pattern back out, CSR order (for the bit-exactness checks) */
int nbgpu_matrix_get_pattern_csr(const nbgpu_matrix_t *A, uint32_t *rows_size,
				 uint32_t *cols);
/* nb_sparse_reset (sparse.c:127-132) */
int nbgpu_matrix_reset(nbgpu_matrix_t *A);

/* ---------------------------------------------------------------- SpMV -- */
/* nb_sparse_multiply_vector (sparse.c:405-414): out = A in.  Each row is
 * summed in ascending column order with separately rounded products and sums,
 * i.e. the same floating-point result as the reference loop. */
/* d_in must be 16-byte aligned (nbgpu_malloc blocks are) when the matrix uses the
 * 2x2-blocked layout; otherwise NBGPU_ERR_ARG. */
int nbgpu_spmv(const nbgpu_matrix_t *A, const double *d_in, double *d_out);
int nbgpu_spmv_host(const nbgpu_matrix_t *A, const double *in, double *out);

/* -------------------------------------------------------------- Krylov -- */
/* nb_sparse_solve_CG_precond_Jacobi (solvers/cg_precond_jacobi.c:13-90,
 * header solvers/cg_precond_jacobi.h:8-15) and
 * nb_sparse_solve_conjugate_gradient (solvers/conjugate_gradient.c:13-77).
 * Same stopping rule: absolute, `while (g.g > tol^2 && k < max_iter)` with the
 * residual of the iterate BEFORE the last update; tol_reached = sqrt of that
 * same quantity.  niter / tol_reached may be NULL. */
int nbgpu_pcg_jacobi(const nbgpu_matrix_t *A, const double *d_b, double *d_x,
		     uint32_t max_iter, double tolerance,
		     uint32_t *niter_performed, double *tolerance_reached);
int nbgpu_cg(const nbgpu_matrix_t *A, const double *d_b, double *d_x,
	     uint32_t max_iter, double tolerance,
	     uint32_t *niter_performed, double *tolerance_reached);
/* host-buffer forms: upload b and x, solve, download x */
int nbgpu_pcg_jacobi_host(const nbgpu_matrix_t *A, const double *b, double *x,
			  uint32_t max_iter, double tolerance,
			  uint32_t *niter_performed, double *tolerance_reached);
int nbgpu_cg_host(const nbgpu_matrix_t *A, const double *b, double *x,
		  uint32_t max_iter, double tolerance,
		  uint32_t *niter_performed, double *tolerance_reached);

/* Order of the solvers' dot-product reductions.
 *   0 (default)  deterministic parallel tree, fused into the kernels;
 *   1            verification mode: one thread sums in index order, exactly as
 *                the reference's single-threaded loops do (the FEM driver runs
 *                the solver with omp_parallel_threads = 1,
 *                static_elasticity2D.c:90).  All other arithmetic of the solver
 *                already rounds like the reference, so a solve in this mode is
 *                bit-identical to the reference's: same iterates, same iteration
 *                count, same tolerance_reached.  Slow; for parity checks. */
int nbgpu_set_reduction_order(int mode);

/* Formulation of the iteration (parallel-tree mode only; reduction order 1 always
 * runs the classic recurrence):
 *   0 (default)  CLASSIC: the reference's recurrence (cg_precond_jacobi.c:45-76),
 *                three kernels, two dependent reductions.
 *   1            FUSED (opt-in): the matrix is applied to q, w = A p is carried by
 *                recurrence and g.q, q.Aq, g.g are reduced together (Chronopoulos &
 *                Gear): two kernels and ONE reduction per iteration -- one exchange
 *                between GPUs instead of two.  Same iterates in exact arithmetic and
 *                the same converged field (1e-10), identical iteration counts on the
 *                cantilever workloads, but not the same rounding: on ill-conditioned
 *                systems the count can leave the +-2 % band (void-material fixture:
 *                359 vs 385).  Pays for systems whose work vectors stay in L2
 *                (250 k dof: -15 % time) and for latency-bound multi-GPU runs; costs
 *                +12 % vector traffic beyond that (4 M dof: +13 % time).
 *  -1            back to the default / NBGPU_PCG_MODE=classic|fused. */
int nbgpu_set_pcg_mode(int mode);

/* Per-kernel timing of the solvers: when enabled, CUDA events bracket the three
 * kernels of each of the first 256 iterations of the next solves, on the
 * library's stream.  _get returns the summed durations [SpMV+dot, update,
 * direction] in ms and the number of iterations they cover (last solve). */
int nbgpu_krylov_profile(int enable);
int nbgpu_krylov_profile_get(double ms_total[3], uint32_t *n_iters);

/* ------------------------------------------------------- FEM assembly -- */

/* node graph + pattern on the host, bit-exact with
 * nb_mesh2D_load_graph(NB_NODES_LINKED_BY_ELEMS) (mesh2D/load_graph.c:230-328)
 * followed by nb_sparse_create(graph, NULL, vars_per_node) (sparse.c:20-60).
 * edg may be NULL (the edges of a conforming mesh are element sides).  Call
 * with cols == NULL to size the output (rows_size is filled, nnz returned). */
int nbgpu_pattern_from_mesh(uint32_t N_nod, uint32_t N_elems,
			    uint32_t nodes_per_elem, const uint32_t *adj,
			    uint32_t N_edg, const uint32_t *edg,
			    uint32_t vars_per_node, uint32_t *rows_size,
			    uint32_t *cols, uint64_t *nnz);

/* mesh arrays as in struct nb_mshquad_s / nb_msh3trg_s (mshquad_struct.h:6-26):
 * nod[2*N_nod] x,y interleaved; adj[nodes_per_elem*N_elems] CCW connectivity;
 * nodes_per_elem is 3 or 4 for the whole mesh */
int nbgpu_mesh_create(uint32_t N_nod, const double *nod, uint32_t N_elems,
		      uint32_t nodes_per_elem, const uint32_t *adj,
		      nbgpu_mesh_t **out);
int nbgpu_mesh_destroy(nbgpu_mesh_t *mesh);

/* The same pattern built ON THE DEVICE from a device-resident mesh, straight into
 * the SELL-32 column arrays (2 dofs per node, values zero): nothing is built on
 * the host or uploaded.  edg (host array) may be NULL; if given,
 * every edge must be an element side.  Returns NBGPU_OK with *out == NULL when
 * the pattern does not qualify for the device path (ragged rows that want the
 * sigma-sorted layout, an edge that is no element side, or one of the layout
 * override switches): use nbgpu_pattern_from_mesh + nbgpu_matrix_create_from_csr
 * then. */
int nbgpu_matrix_create_from_mesh(const nbgpu_mesh_t *mesh, uint32_t N_edg,
				  const uint32_t *edg, nbgpu_matrix_t **out);
/* number of colours of the device-built element colouring the COLOR schedule
 * uses (built on first use); colors[N_elems] gets each element's colour if not NULL */
int nbgpu_mesh_coloring(nbgpu_mesh_t *mesh, uint32_t *n_colors, uint8_t *colors);

/* element tables as struct nb_fem_elem_s holds them (element_struct.h:8-16),
 * indexed [node * N_gp + gp] (element.c:144-160) */
typedef struct {
	uint32_t N_nodes;      /* 3 or 4 */
	uint32_t N_gp;         /* 1 or 4 */
	double gp_weight[4];
	double Ni[16];
	double dNi_dpsi[16];
	double dNi_deta[16];
} nbgpu_elem_tables_t;

/* the reference's own tables (element.c:54-121), 12-digit quad literals */
int nbgpu_elem_tables_default(uint32_t nodes_per_elem, nbgpu_elem_tables_t *t);
/* nb_pde_get_constitutive_matrix (common_solid_mechanics/formulas.c:32-63),
 * including its switch fall-through: plane-stress D for every analysis id */
int nbgpu_constitutive_matrix(double E, double poisson, int analysis2D,
			      double D[4]);

#define NBGPU_ASSEMBLY_GATHER  0  /* row-parallel gather, bit-exact order   */
#define NBGPU_ASSEMBLY_ATOMIC  1  /* element-parallel, atomicAdd(double)    */
#define NBGPU_ASSEMBLY_COLOR   2  /* element-parallel, colour-scheduled     */

typedef struct {
	double D[4];           /* enabled elements (a12)                       */
	double density;
	double D_void[4];      /* disabled elements: 1e-6 x4 (pipeline.c:93)   */
	double density_void;   /* 1e-6 (pipeline.c:94)                         */
	double thickness;      /* params2D->thickness                          */
	int32_t self_weight;
	double gravity[2];
	int32_t mode;          /* NBGPU_ASSEMBLY_*                             */
} nbgpu_assembly_params_t;

/* pipeline_assemble_system (solid_mechanics/pipeline.c:42-73 and the helpers
 * down to :264; utils.c:9-60): resets K, zeroes F, integrates every element
 * (B'DB detJ t w per Gauss point) into K and the self-weight into F.
 * enabled: host array, one byte per element, NULL = all enabled.
 * elem_scale: optional host array of per-element stiffness factors applied to
 * D of enabled elements (SIMP-style loops); NULL = 1.
 * Returns 0, or 1 if any element has detJ < 0 (first_bad gets the lowest such
 * element id when not NULL). */
int nbgpu_assemble_elasticity2d(nbgpu_matrix_t *K, const nbgpu_mesh_t *mesh,
				const nbgpu_elem_tables_t *tables,
				const nbgpu_assembly_params_t *params,
				const uint8_t *enabled,
				const double *elem_scale, double *d_F,
				uint32_t *first_bad);

/* The assembly loop of the reference's damage driver, DMG_pipeline_assemble_system
 * (solid_mechanics/static_damage2D.c:474-569): as nbgpu_assemble_elasticity2d,
 * with the constitutive matrix of Gauss point j of element k multiplied by
 * (1 - gp_damage[k * N_gp + j]) (:530-536; the void material of a disabled element
 * is scaled too).  gp_damage: host array, N_elems * N_gp doubles.  GATHER schedule on
 * a 2-dof blocked matrix only (NBGPU_ERR_ARG otherwise); bit-exact.  The reference
 * leaves its loop silently at a distorted element; here that is status 1. */
int nbgpu_assemble_elasticity2d_damage(nbgpu_matrix_t *K, const nbgpu_mesh_t *mesh,
				       const nbgpu_elem_tables_t *tables,
				       const nbgpu_assembly_params_t *params,
				       const uint8_t *enabled,
				       const double *gp_damage, double *d_F,
				       uint32_t *first_bad);

/* The lumped mass vector pipeline_assemble_system fills when its M argument is
 * not NULL (solid_mechanics/pipeline.c:56-57 zeroed, :216-222 Me[2i] +=
 * Ni Nj density detJ thickness w over j and the Gauss points, :256-259
 * M[2v+a] += Me[2i+a] in element order).  d_M: 2 N_nod doubles on the device.
 * density_void: density of disabled elements (the reference uses 1e-6,
 * pipeline.c:93-94).  enabled / first_bad / return value as above. */
int nbgpu_assemble_lumped_mass(const nbgpu_mesh_t *mesh,
			       const nbgpu_elem_tables_t *tables,
			       double density, double density_void,
			       double thickness, const uint8_t *enabled,
			       double *d_M, uint32_t *first_bad);

/* F[dof[k]] += add[k], k in order (the Neumann part of nb_fem_set_bconditions,
 * solid_mechanics/set_bconditions.c:63-188, flattened by the caller) */
int nbgpu_vector_add_entries(double *d_F, uint32_t n, const uint32_t *dof,
			     const double *add);

/* nb_sparse_set_Dirichlet_condition (sparse.c:416-430) for a list of
 * constrained dofs applied in list order (set_bconditions.c:190-262) */
int nbgpu_apply_dirichlet(nbgpu_matrix_t *K, double *d_F, uint32_t n,
			  const uint32_t *dof, const double *value);

/* The same two steps with the lists kept on the device, for loops that re-assemble
 * and re-apply unchanged boundary conditions (no allocation, no sync per step). */
typedef struct nbgpu_dirichlet_s nbgpu_dirichlet_t;
int nbgpu_dirichlet_create(uint32_t N, uint32_t n, const uint32_t *dof,
			   const double *value, nbgpu_dirichlet_t **out);
int nbgpu_dirichlet_apply(nbgpu_matrix_t *K, double *d_F, const nbgpu_dirichlet_t *bc);
int nbgpu_dirichlet_destroy(nbgpu_dirichlet_t *bc);
int nbgpu_vector_add_entries_dev(double *d_F, uint32_t n, const uint32_t *d_dof,
				 const double *d_add);

/* pipeline_compute_strain (pipeline.c:266-319): strain[3*N_gp*N_elems] */
int nbgpu_compute_strain(const nbgpu_mesh_t *mesh,
			 const nbgpu_elem_tables_t *tables,
			 const double *d_disp, double *d_strain);
/* nb_fem_compute_stress_from_strain (static_elasticity2D.c:99-127) */
int nbgpu_stress_from_strain(uint32_t N_elems, uint32_t N_gp,
			     const double D[4], const double D_void[4],
			     const uint8_t *enabled, const double *d_strain,
			     double *d_stress);

/* nb_fem_interpolate_from_gpoints_to_nodes (finite_element/gaussp_to_nodes.c:50-75):
 * lumped-mass projection of N_comp values per Gauss point onto the nodes,
 * d_gp_values[(elem * N_gp + gp) * N_comp + c] -> d_nodal_values[node * N_comp + c].
 * Returns 1 when an element is distorted (detJ < 0); the reference leaves the
 * output untouched in that case, here its contents are then unspecified. */
int nbgpu_gp_to_nodes(const nbgpu_mesh_t *mesh,
		      const nbgpu_elem_tables_t *tables, uint32_t N_comp,
		      const double *d_gp_values, double *d_nodal_values);
/* nb_pde_get_vm_stress / nb_pde_get_main_stress (common_solid_mechanics/formulas.c:65-77)
 * over n_points stress triplets [sxx, syy, sxy]; d_main gets 2 values per point */
int nbgpu_von_mises(uint64_t n_points, const double *d_stress, double *d_vm);
int nbgpu_main_stress(uint64_t n_points, const double *d_stress, double *d_main);

/* ------------------------------------------- boundary-condition lists -- */

/* value callback of a function-valued condition, as nb_bcond_push_function
 * takes it (headers/nb/pde_bot/boundary_conditions/bcond.h) */
typedef void (*nbgpu_bc_fn)(const double *x, double t, double *out);

/* one nb_bcond_push / nb_bcond_push_function call (bcond.c:153-165) */
typedef struct {
	int32_t kind;          /* 0 Dirichlet, 1 Neumann (nb_bcond_id)           */
	int32_t where;         /* 0 input vertex, 1 input segment (nb_bcond_where) */
	uint32_t id;           /* input vertex / segment id                       */
	int32_t mask[2];       /* dof mask                                        */
	double val[2];         /* constant value (ignored when fval != NULL)      */
	nbgpu_bc_fn fval;      /* NULL = constant                                 */
} nbgpu_bcond_t;

/* Host side of nb_fem_set_bconditions (solid_mechanics/set_bconditions.c:52-262):
 * turns the condition queues into two ORDERED dof lists -- Neumann adds
 * (segments first, then vertices) and Dirichlet (dof, value) pairs (segments
 * first, then vertices) -- with exactly the reference's arithmetic for the
 * nodal shares.  Call with the output arrays NULL to get the counts. */
int nbgpu_bcond_flatten(const double *nod, const uint32_t *vtx,
			uint32_t N_sgm, const uint32_t *sgm_sizes,
			const uint32_t *sgm_nodes, uint32_t N_bc,
			const nbgpu_bcond_t *bc, double factor,
			uint32_t *n_neumann, uint32_t *neumann_dof,
			double *neumann_add, uint32_t *n_dirichlet,
			uint32_t *dirichlet_dof, double *dirichlet_val);

/* ------------------------------------------------------- FEM driver -- */

/* flat view of a struct nb_mshquad_s / nb_msh3trg_s mesh */
typedef struct {
	uint32_t N_nod;
	const double *nod;          /* [2 N_nod]                                */
	uint32_t N_elems;
	uint32_t nodes_per_elem;    /* 3 or 4                                   */
	const uint32_t *adj;        /* [nodes_per_elem N_elems]                 */
	uint32_t N_edg;
	const uint32_t *edg;        /* [2 N_edg] or NULL                        */
	uint32_t N_vtx;
	const uint32_t *vtx;        /* mesh node of every input vertex          */
	uint32_t N_sgm;
	const uint32_t *sgm_sizes;  /* nodes per input segment                  */
	const uint32_t *sgm_nodes;  /* concatenated node ids along the segments */
} nbgpu_mesh_desc_t;

typedef struct {
	uint32_t N;                 /* dofs                                     */
	uint64_t nnz;
	uint32_t solver_iters;
	int32_t solver_status;      /* 0 converged, 1 max_iter                  */
	double solver_residual;     /* tolerance_reached                        */
	double ms_pattern, ms_upload, ms_assembly, ms_bcond, ms_solve, ms_post;
} nbgpu_fem_report_t;

/* nb_fem_compute_2D_Solid_Mechanics
 * (solid_mechanics/static_elasticity2D.c:31-97, header
 * headers/nb/pde_bot/finite_element/solid_mechanics/static_elasticity2D.h:13-24):
 * pattern -> assembly -> boundary conditions -> Jacobi-PCG (x0 = 0, abs tol
 * 1e-8, max_iter = N; status 1 accepted like :92) -> strain.  K, F and the
 * iterates never leave the device.  Returns 0, or 1 when assembly met a
 * distorted element (then displacement/strain are untouched), or >= 10.
 * tables == NULL uses nbgpu_elem_tables_default; solver_tol <= 0 means the
 * reference's 1e-8; report may be NULL. */
/* Same pipeline with the boundary conditions already flattened into ordered
 * dof lists (what nbgpu_bcond_flatten produces) and the constitutive matrix
 * given directly; this is the form the reference-named shim uses after
 * walking the reference's own nb_bcond_t and calling its own
 * nb_pde_get_constitutive_matrix. */
int nbgpu_fem_static_elasticity2d_lists(const nbgpu_mesh_desc_t *mesh,
					const nbgpu_elem_tables_t *tables,
					const double D[4], double density,
					uint32_t n_neumann,
					const uint32_t *neumann_dof,
					const double *neumann_add,
					uint32_t n_dirichlet,
					const uint32_t *dirichlet_dof,
					const double *dirichlet_val,
					int self_weight, const double gravity[2],
					int analysis2D, double thickness,
					const uint8_t *enabled, int assembly_mode,
					double solver_tol, double *displacement,
					double *strain, nbgpu_fem_report_t *report);

/* Device-resident session for repeated assembly + solve on one mesh (the call
 * pattern of the reference's damage loop, static_damage2D.c:300-460, and of a
 * SIMP topology-optimisation loop): pattern, matrix and mesh are built once;
 * each step re-assembles with an optional enabled mask (pipeline.c:93-98) and
 * optional per-element stiffness factors, re-applies the boundary conditions
 * and runs Jacobi-PCG, warm-started from the previous displacement if asked.
 * max_iter 0 = N, solver_tol <= 0 = 1e-8 (static_elasticity2D.c:88-90). */
typedef struct nbgpu_fem_session_s nbgpu_fem_session_t;
int nbgpu_fem_session_create(const nbgpu_mesh_desc_t *mesh,
			     const nbgpu_elem_tables_t *tables, const double D[4],
			     double density, uint32_t n_neumann,
			     const uint32_t *neumann_dof, const double *neumann_add,
			     uint32_t n_dirichlet, const uint32_t *dirichlet_dof,
			     const double *dirichlet_val, int self_weight,
			     const double gravity[2], double thickness,
			     int assembly_mode, nbgpu_fem_session_t **out);
int nbgpu_fem_session_step(nbgpu_fem_session_t *session, const uint8_t *enabled,
			   const double *elem_scale, int warm_start,
			   uint32_t max_iter, double solver_tol,
			   nbgpu_fem_report_t *report);
int nbgpu_fem_session_results(nbgpu_fem_session_t *session, double *displacement,
			      double *strain);
int nbgpu_fem_session_destroy(nbgpu_fem_session_t *session);

int nbgpu_fem_static_elasticity2d(const nbgpu_mesh_desc_t *mesh,
				  const nbgpu_elem_tables_t *tables,
				  double E, double poisson, double density,
				  uint32_t N_bc, const nbgpu_bcond_t *bc,
				  int self_weight, const double gravity[2],
				  int analysis2D, double thickness,
				  const uint8_t *enabled, int assembly_mode,
				  double solver_tol, double *displacement,
				  double *strain, nbgpu_fem_report_t *report);

/* -------------------------------------------------------- multi-GPU -- */
/* Row-partitioned solves over several GPUs of one node, one process per GPU
 * (SURVEY.md §8e).  The reference has no distributed mode; these entry points
 * mirror the single-GPU ones with a partition plan added.  Halo values and
 * dot-product partials travel as NVLink peer stores into CUDA-IPC windows; no
 * collective library is called inside the iteration. */

typedef struct nbgpu_dist_s nbgpu_dist_t;             /* this rank's window + peers */
typedef struct nbgpu_dist_plan_s nbgpu_dist_plan_t;   /* halo / send lists of one matrix */

#define NBGPU_IPC_HANDLE_BYTES 64

/* Column space of a rank-local block: "lower halo | owned | upper halo" -- the global
 * order with the remote ranges squeezed out, each part starting on its own
 * 128-byte line: lower halo at 0, owned columns at off_own, upper halo at off_up,
 * ext_len entries in all. */
int nbgpu_dist_ext_layout(uint32_t n_lo, uint32_t N_loc, uint32_t n_hi,
			  uint32_t *off_own, uint32_t *off_up, uint32_t *ext_len);

/* a rank-local block of the matrix: N_rows owned rows over a column space of
 * N_cols (= ext_len) entries in which row r is column r + col_shift (= off_own);
 * cols_local from nbgpu_dist_plan_local_cols (ascending, like the global ids) */
int nbgpu_matrix_create_local(uint32_t N_rows, uint32_t N_cols, uint32_t col_shift,
			      const uint32_t *rows_size,
			      const uint32_t *cols_local, const double *vals,
			      nbgpu_matrix_t **out);

/* Host logic (no device needed): rank `rank` of `world` owns global rows
 * [row_starts[rank], row_starts[rank+1]); rows_size / cols_global describe those
 * rows in CSR form with GLOBAL column ids. */
int nbgpu_dist_plan_create(int rank, int world, const uint32_t *row_starts,
			   const uint32_t *rows_size, const uint32_t *cols_global,
			   nbgpu_dist_plan_t **out);
int nbgpu_dist_plan_destroy(nbgpu_dist_plan_t *plan);
/* recv_counts[world]: how many halo values come from each rank */
int nbgpu_dist_plan_info(const nbgpu_dist_plan_t *plan, uint32_t *N_loc,
			 uint32_t *n_halo, uint64_t *nnz, uint32_t *recv_counts);
/* n_lo halo columns lie below the owned range; layout as nbgpu_dist_ext_layout */
int nbgpu_dist_plan_layout(const nbgpu_dist_plan_t *plan, uint32_t *n_lo,
			   uint32_t *off_own, uint32_t *off_up, uint32_t *ext_len);
/* The order in which the SpMV kernels visit this rank's 32-row slices: visit index v
 * -> slice (v + visit_shift) mod n_slices; the slices that read halo columns are the
 * visits [late_from, late_to).  total_warps = 0: the plan's own order (they come last;
 * init kernel, nbgpu_dist_spmv); total_warps = streaming warps of the solver's SpMV
 * kernel: they end with the last full round of slices when the final partial round
 * leaves enough warps without a slice (hides their wait for the neighbours). */
int nbgpu_dist_plan_visit_order(const nbgpu_dist_plan_t *plan, uint32_t total_warps,
				uint32_t *visit_shift, uint32_t *late_from, uint32_t *late_to);
/* the halo columns (global row ids, ascending => grouped by owner) */
int nbgpu_dist_plan_halo_ids(const nbgpu_dist_plan_t *plan, uint32_t *halo_global);
int nbgpu_dist_plan_local_cols(const nbgpu_dist_plan_t *plan, uint32_t *cols_local);
/* what the other ranks need from me (the transpose of their halo lists, which the
 * processes exchange by any means): send_global grouped by destination in the
 * order of the destination's halo list; dst_offsets[d] = where my block starts in
 * rank d's COLUMN SPACE (its position in d's halo list mapped through d's
 * nbgpu_dist_ext_layout) */
int nbgpu_dist_plan_set_sends(nbgpu_dist_plan_t *plan, const uint32_t *send_counts,
			      const uint32_t *send_global, const uint32_t *dst_offsets);

/* ext_len from nbgpu_dist_plan_layout.  Writes this rank's 64-byte CUDA IPC handle
 * (ipc_handle_out may be NULL for ranks that connect with nbgpu_dist_connect_local). */
int nbgpu_dist_create(int rank, int world, size_t ext_len, void *ipc_handle_out,
		      nbgpu_dist_t **out);
/* all_handles: world x 64 bytes in rank order; all_ext_len: every rank's ext_len */
int nbgpu_dist_connect(nbgpu_dist_t *dist, const void *all_handles,
		       const uint64_t *all_ext_len);
/* ranks living in ONE process (one host thread per GPU, nbgpu_thread_bind_device):
 * all[r] = rank r's object, device_of[r] = its GPU; peer access instead of IPC */
int nbgpu_dist_connect_local(nbgpu_dist_t *dist, nbgpu_dist_t *const *all,
			     const int *device_of);
int nbgpu_dist_destroy(nbgpu_dist_t *dist);
int nbgpu_dist_error(nbgpu_dist_t *dist);

/* Collective over all ranks (every rank calls with its own block).  b, x: this
 * rank's rows; tolerance is the absolute bound on the GLOBAL residual norm, as in
 * the single-GPU call; every rank returns the same status / iteration count. */
int nbgpu_dist_pcg_jacobi(nbgpu_dist_t *dist, nbgpu_dist_plan_t *plan,
			  const nbgpu_matrix_t *A_local, const double *d_b,
			  double *d_x, uint32_t max_iter, double tolerance,
			  uint32_t *niter_performed, double *tolerance_reached);
int nbgpu_dist_cg(nbgpu_dist_t *dist, nbgpu_dist_plan_t *plan,
		  const nbgpu_matrix_t *A_local, const double *d_b, double *d_x,
		  uint32_t max_iter, double tolerance, uint32_t *niter_performed,
		  double *tolerance_reached);
int nbgpu_dist_spmv(nbgpu_dist_t *dist, nbgpu_dist_plan_t *plan,
		    const nbgpu_matrix_t *A_local, const double *d_in, double *d_out);
/* device pointer to the owned part of the window's input vector: an SpMV caller
 * that writes x there and passes it as d_in saves the copy into the window */
double *nbgpu_dist_input_vector(nbgpu_dist_t *dist, const nbgpu_dist_plan_t *plan);

/* ---------------------------------------- multi-GPU FEM (row slabs) -- */
/* nb_fem_compute_2D_Solid_Mechanics (static_elasticity2D.c:31-97) on a mesh
 * partitioned into contiguous NODE ranges, one rank per GPU: every rank gets the
 * whole mesh description, integrates the sub-mesh of all elements that touch its
 * nodes straight into its rank-local block ON THE DEVICE (elements on a cut are
 * integrated on both sides: no communication, K never visits the host), applies
 * the boundary conditions restricted to its rows, and the ranks solve together.
 * The rank-local rows are bit-identical to the rows of the single-GPU matrix.
 * All halo / send lists are derived locally from the mesh: the ranks exchange
 * only the 64-byte IPC handles (nbgpu_dist_fem_connect), or nothing when they
 * live in one process (nbgpu_dist_fem_connect_local). */
typedef struct nbgpu_dist_fem_s nbgpu_dist_fem_t;

/* balanced contiguous node ranges, node_starts[world + 1]; boundaries fall on
 * multiples of `align` nodes (nodes per grid line => slabs of whole lines) */
int nbgpu_partition_nodes(uint32_t N_nod, int world, uint32_t align,
			  uint32_t *node_starts);
/* boundary-condition lists: GLOBAL dof ids, ordered, as nbgpu_bcond_flatten
 * produces them.  ipc_handle_out (64 bytes) may be NULL for in-process ranks. */
int nbgpu_dist_fem_create(const nbgpu_mesh_desc_t *mesh, int rank, int world,
			  const uint32_t *node_starts,
			  const nbgpu_elem_tables_t *tables, const double D[4],
			  double density, uint32_t n_neumann,
			  const uint32_t *neumann_dof, const double *neumann_add,
			  uint32_t n_dirichlet, const uint32_t *dirichlet_dof,
			  const double *dirichlet_val, int self_weight,
			  const double gravity[2], double thickness,
			  void *ipc_handle_out, nbgpu_dist_fem_t **out);
int nbgpu_dist_fem_destroy(nbgpu_dist_fem_t *fem);
/* Host logic only (no device): the partition plan nbgpu_dist_fem_create derives
 * from the mesh by itself -- halo list, column-space layout, local column ids of
 * the owned rows AND the send lists (no exchange of lists between the ranks).
 * rows_size may be NULL or [2 * owned nodes]; read the rest back with
 * nbgpu_dist_plan_info / _layout / _halo_ids / _local_cols / _sends. */
int nbgpu_dist_plan_from_mesh(const nbgpu_mesh_desc_t *mesh, int rank, int world,
			      const uint32_t *node_starts, uint32_t *rows_size,
			      nbgpu_dist_plan_t **out);
/* send side of a plan: counts per destination, the global rows sent (grouped by
 * destination; NULL to get the counts first), where each block lands in the
 * destination's column space */
int nbgpu_dist_plan_sends(const nbgpu_dist_plan_t *plan, uint32_t *send_counts,
			  uint32_t *send_global, uint32_t *dst_offsets);
int nbgpu_dist_fem_info(const nbgpu_dist_fem_t *fem, uint32_t *N_loc,
			uint64_t *nnz_loc, uint32_t *n_halo, uint32_t *n_elems_loc,
			uint64_t *ext_len, double *ms_setup);
int nbgpu_dist_fem_connect(nbgpu_dist_fem_t *fem, const void *all_handles,
			   const uint64_t *all_ext_len);
int nbgpu_dist_fem_connect_local(nbgpu_dist_fem_t *fem,
				 nbgpu_dist_fem_t *const *all,
				 const int *device_of);
/* pipeline_assemble_system (pipeline.c:42-73) + nb_fem_set_bconditions
 * (set_bconditions.c:52-61) for this rank's rows; enabled / elem_scale are the
 * GLOBAL per-element arrays or NULL.  0, or 1 = distorted element (lowest
 * global id of this rank's sub-mesh in *first_bad). */
int nbgpu_dist_fem_assemble(nbgpu_dist_fem_t *fem, const uint8_t *enabled,
			    const double *elem_scale, uint32_t *first_bad);
/* collective Jacobi-PCG; max_iter 0 = global N, tolerance <= 0 = 1e-8 */
int nbgpu_dist_fem_solve(nbgpu_dist_fem_t *fem, int warm_start,
			 uint32_t max_iter, double tolerance,
			 uint32_t *niter_performed, double *tolerance_reached);
/* displacements of this rank's nodes: 2 * (node_starts[rank+1] - node_starts[rank]) */
int nbgpu_dist_fem_results(nbgpu_dist_fem_t *fem, double *displacement_owned);
/* the objects behind it (owned by fem) */
nbgpu_matrix_t *nbgpu_dist_fem_matrix(nbgpu_dist_fem_t *fem);
nbgpu_dist_plan_t *nbgpu_dist_fem_plan(nbgpu_dist_fem_t *fem);
nbgpu_dist_t *nbgpu_dist_fem_dist(nbgpu_dist_fem_t *fem);
double *nbgpu_dist_fem_rhs(nbgpu_dist_fem_t *fem);        /* device, N_loc */
double *nbgpu_dist_fem_solution(nbgpu_dist_fem_t *fem);   /* device, N_loc */

/* ------------------------- several GPUs from one (single-threaded) caller -- */
/* The reference's callers are ordinary single-process programs.  These entry
 * points keep that shape: the call blocks, inside it one worker thread per GPU
 * runs the rank code above with peer access between the windows.  The
 * reference-named shims use them when NBGPU_DEVICES=N (N > 1) is set. */
int nbgpu_devices_from_env(void);
/* nb_fem_compute_2D_Solid_Mechanics (static_elasticity2D.c:31-97) on n_devices
 * GPUs; arguments as nbgpu_fem_static_elasticity2d_lists (row-parallel assembly) */
int nbgpu_fem_static_elasticity2d_lists_multi(int n_devices,
					      const nbgpu_mesh_desc_t *mesh,
					      const nbgpu_elem_tables_t *tables,
					      const double D[4], double density,
					      uint32_t n_neumann,
					      const uint32_t *neumann_dof,
					      const double *neumann_add,
					      uint32_t n_dirichlet,
					      const uint32_t *dirichlet_dof,
					      const double *dirichlet_val,
					      int self_weight, const double gravity[2],
					      double thickness, const uint8_t *enabled,
					      double solver_tol, double *displacement,
					      double *strain, nbgpu_fem_report_t *report);
/* nb_sparse_solve_CG_precond_Jacobi (jacobi != 0) / nb_sparse_solve_conjugate_gradient
 * on the three arrays of a host nb_sparse_t, contiguous row blocks over n_devices GPUs */
int nbgpu_solve_rows_multi(int n_devices, int jacobi, uint32_t N,
			   const uint32_t *rows_size, uint32_t *const *rows_index,
			   double *const *rows_values, const double *b, double *x,
			   uint32_t max_iter, double tolerance,
			   uint32_t *niter_performed, double *tolerance_reached);

#ifdef __cplusplus
}
#endif
#endif /* NBGPU_H */
