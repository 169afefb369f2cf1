/*
 * oracle/port/nbo_fem.c -- TEST INFRASTRUCTURE (see nbo.h).
 * 2-D linear-elastic FEM assembly, boundary conditions and strain/stress
 * recovery restated from the reference's PDE bot over flat arrays.
 * Compiled with -ffp-contract=off: the reference binary (x86-64, no -mfma)
 * evaluates every product and sum separately, and so does this port, so that
 * assembled values can be compared bit for bit.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "nbo.h"

uint64_t nbo_find_entry(const uint64_t *row_ptr, const uint32_t *cols,
			uint32_t i, uint32_t col);

/* Reference: sources/nb/pde_bot/finite_element/element.c:54-121.
 * Triangle: 1 Gauss point, weight 0.5, N = 1/3, constant gradients.
 * Quad: 2x2 Gauss points ordered (-,-),(+,-),(+,+),(-,+), weight 1; the
 * reference tabulates N and dN at the points with 12 significant digits and
 * parity requires those literals, not the exact (1 +- 1/sqrt3) values. */
void nbo_elem_tables(int elem_type, uint32_t *N_nodes, uint32_t *N_gp,
		     double *w, double *Ni, double *dpsi, double *deta)
{
	if (!elem_type) {
		*N_nodes = 3;
		*N_gp = 1;
		w[0] = 0.5;
		for (int i = 0; i < 3; i++)
			Ni[i] = 0.33333333333333333333333333333;
		dpsi[0] = -1.0; dpsi[1] = 1.0; dpsi[2] = 0.0;
		deta[0] = -1.0; deta[1] = 0.0; deta[2] = 1.0;
		return;
	}
	*N_nodes = 4;
	*N_gp = 4;
	/* values of (1-+g)(1-+g)/4 and (1-+g)/4 at g = 1/sqrt(3), 12 digits */
	const double Nbig = 0.622008467928, Nmid = 0.166666666667,
		     Nsml = 0.044658198739;
	const double dbig = 0.394337567297, dsml = 0.105662432703;
	/* corner c sits at (sx[c], sy[c]); Gauss point g at (sx[g], sy[g])/sqrt3 */
	static const int sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
	for (int g = 0; g < 4; g++)
		w[g] = 1.0;
	for (int c = 0; c < 4; c++) {
		for (int g = 0; g < 4; g++) {
			int nearx = sx[c] == sx[g], neary = sy[c] == sy[g];
			double N = (nearx && neary) ? Nbig :
				   (nearx || neary) ? Nmid : Nsml;
			/* dN/dpsi = sx_c (1 + sy_c eta_g)/4, dN/deta likewise */
			double dp = sx[c] * (neary ? dbig : dsml);
			double de = sy[c] * (nearx ? dbig : dsml);
			Ni[c * 4 + g] = N;
			dpsi[c * 4 + g] = dp;
			deta[c * 4 + g] = de;
		}
	}
}

/* Reference: sources/nb/pde_bot/common_solid_mechanics/formulas.c:32-63.
 * The switch in nb_pde_get_constitutive_matrix has no breaks (:38-45), so
 * every analysis type ends in set_plane_stress: D is plane stress ALWAYS. */
void nbo_constitutive(double E, double nu, int analysis, double D[4])
{
	(void)analysis;
	D[0] = E / (1.0 - nu * nu);
	D[1] = nu * D[0];
	D[2] = D[0];
	D[3] = E / (2.0 * (1.0 + nu));
}

typedef struct {
	uint32_t n, ngp;
	double w[4], Ni[16], dpsi[16], deta[16];
} elem_t;

/* Reference: finite_element/utils.c:9-41 (Jacobian at one Gauss point) and
 * :49-60 (Cartesian shape-function gradients). */
static double jacobian_and_gradients(const elem_t *el, const double *nod,
				     const uint32_t *v, uint32_t gp,
				     double *dNdx, double *dNdy)
{
	double x_psi = 0.0, y_psi = 0.0, x_eta = 0.0, y_eta = 0.0;
	for (uint32_t i = 0; i < el->n; i++) {
		double xi = nod[2 * v[i]], yi = nod[2 * v[i] + 1];
		double dp = el->dpsi[i * el->ngp + gp];
		double de = el->deta[i * el->ngp + gp];
		x_psi += dp * xi;
		x_eta += de * xi;
		y_psi += dp * yi;
		y_eta += de * yi;
	}
	double detJ = x_psi * y_eta - y_psi * x_eta;
	double Jinv[4] = { y_eta / detJ, -y_psi / detJ,
			  -x_eta / detJ,  x_psi / detJ};
	for (uint32_t i = 0; i < el->n; i++) {
		double dp = el->dpsi[i * el->ngp + gp];
		double de = el->deta[i * el->ngp + gp];
		dNdx[i] = Jinv[0] * dp + Jinv[1] * de;
		dNdy[i] = Jinv[2] * dp + Jinv[3] * de;
	}
	return detJ;
}

/* Reference: solid_mechanics/pipeline.c:42-73 (element loop), :83-122
 * (material choice: disabled elements get D = {1e-6 x4}, density 1e-6),
 * :124-172 (Gauss loop, stop at detJ < 0), :174-230 (B'DB accumulation,
 * expression order kept), :232-264 (scatter by node pairs). */
static int assemble_impl(uint32_t N_nod, const double *nod, uint32_t N_elems,
			 int elem_type, const uint32_t *adj, double E, double nu,
			 double density, int self_weight, double gx, double gy,
			 int analysis, double thickness, const uint8_t *enabled,
			 const double *elem_scale, const double *gp_damage,
			 const uint64_t *row_ptr, const uint32_t *cols, double *vals,
			 double *F);

int nbo_assemble(uint32_t N_nod, const double *nod, uint32_t N_elems,
		 int elem_type, const uint32_t *adj, double E, double nu,
		 double density, int self_weight, double gx, double gy,
		 int analysis, double thickness, const uint8_t *enabled,
		 const uint64_t *row_ptr, const uint32_t *cols, double *vals,
		 double *F)
{
	return assemble_impl(N_nod, nod, N_elems, elem_type, adj, E, nu, density,
			     self_weight, gx, gy, analysis, thickness, enabled,
			     NULL, NULL, row_ptr, cols, vals, F);
}

/* The damage driver's assembly loop, static_damage2D.c:474-569: as above, with
 * the constitutive matrix of Gauss point j of element k multiplied by
 * (1 - gp_damage[k * N_gp + j]) (:530-536) -- applied to the void material of
 * a disabled element as well.  (The reference leaves its loop silently at a
 * distorted element, :543-544; here that is status 1.) */
int nbo_assemble_damage(uint32_t N_nod, const double *nod, uint32_t N_elems,
			int elem_type, const uint32_t *adj, double E, double nu,
			double density, int self_weight, double gx, double gy,
			int analysis, double thickness, const uint8_t *enabled,
			const double *gp_damage,
			const uint64_t *row_ptr, const uint32_t *cols, double *vals,
			double *F)
{
	return assemble_impl(N_nod, nod, N_elems, elem_type, adj, E, nu, density,
			     self_weight, gx, gy, analysis, thickness, enabled,
			     NULL, gp_damage, row_ptr, cols, vals, F);
}

/* Extension used only to check the product's SIMP-style hook (the reference has
 * no such parameter): the constitutive matrix of every ENABLED element is
 * multiplied by elem_scale[e] before integration; everything else as above. */
int nbo_assemble_scaled(uint32_t N_nod, const double *nod, uint32_t N_elems,
			int elem_type, const uint32_t *adj, double E, double nu,
			double density, int self_weight, double gx, double gy,
			int analysis, double thickness, const uint8_t *enabled,
			const double *elem_scale,
			const uint64_t *row_ptr, const uint32_t *cols, double *vals,
			double *F)
{
	return assemble_impl(N_nod, nod, N_elems, elem_type, adj, E, nu, density,
			     self_weight, gx, gy, analysis, thickness, enabled,
			     elem_scale, NULL, row_ptr, cols, vals, F);
}

static int assemble_impl(uint32_t N_nod, const double *nod, uint32_t N_elems,
			 int elem_type, const uint32_t *adj, double E, double nu,
			 double density, int self_weight, double gx, double gy,
			 int analysis, double thickness, const uint8_t *enabled,
			 const double *elem_scale, const double *gp_damage,
			 const uint64_t *row_ptr, const uint32_t *cols, double *vals,
			 double *F)
{
	elem_t el;
	nbo_elem_tables(elem_type, &el.n, &el.ngp, el.w, el.Ni, el.dpsi,
			el.deta);
	uint32_t n = el.n, N = 2 * N_nod;
	memset(vals, 0, row_ptr[N] * sizeof(double));
	memset(F, 0, (size_t)N * sizeof(double));
	for (uint32_t e = 0; e < N_elems; e++) {
		const uint32_t *v = adj + (size_t)n * e;
		double D[4] = {1e-6, 1e-6, 1e-6, 1e-6};
		double rho = 1e-6;
		if (!enabled || enabled[e]) {
			nbo_constitutive(E, nu, analysis, D);
			rho = density;
			if (elem_scale)
				for (int k = 0; k < 4; k++)
					D[k] = D[k] * elem_scale[e];
		}
		double fx = 0.0, fy = 0.0;
		if (self_weight) {
			fx = gx * rho;
			fy = gy * rho;
		}
		double Ke[64], Fe[8], dx[4], dy[4];
		memset(Ke, 0, sizeof(Ke));
		memset(Fe, 0, sizeof(Fe));
		for (uint32_t gp = 0; gp < el.ngp; gp++) {
			double detJ = jacobian_and_gradients(&el, nod, v, gp,
							     dx, dy);
			if (detJ < 0)
				return 1;
			double wp = el.w[gp];
			double Dr[4] = {D[0], D[1], D[2], D[3]};
			if (gp_damage)   /* static_damage2D.c:530-536 */
				for (int k = 0; k < 4; k++)
					Dr[k] *= (1.0 - gp_damage[(size_t)e * el.ngp + gp]);
			for (uint32_t i = 0; i < n; i++) {
				for (uint32_t j = 0; j < n; j++) {
					Ke[(2 * i) * (2 * n) + 2 * j] +=
						(dx[i] * dx[j] * Dr[0] +
						 dy[i] * dy[j] * Dr[3]) *
						detJ * thickness * wp;
					Ke[(2 * i) * (2 * n) + 2 * j + 1] +=
						(dx[i] * dy[j] * Dr[1] +
						 dy[i] * dx[j] * Dr[3]) *
						detJ * thickness * wp;
					Ke[(2 * i + 1) * (2 * n) + 2 * j] +=
						(dy[i] * dx[j] * Dr[1] +
						 dx[i] * dy[j] * Dr[3]) *
						detJ * thickness * wp;
					Ke[(2 * i + 1) * (2 * n) + 2 * j + 1] +=
						(dy[i] * dy[j] * Dr[2] +
						 dx[i] * dx[j] * Dr[3]) *
						detJ * thickness * wp;
				}
				double integral = el.Ni[i * el.ngp + gp] *
					detJ * thickness * wp;
				Fe[2 * i] += integral * fx;
				Fe[2 * i + 1] += integral * fy;
			}
		}
		for (uint32_t i = 0; i < n; i++) {
			for (uint32_t j = 0; j < n; j++) {
				for (uint32_t a = 0; a < 2; a++) {
					for (uint32_t c = 0; c < 2; c++) {
						uint32_t r = 2 * v[i] + a;
						uint32_t col = 2 * v[j] + c;
						uint64_t m = nbo_find_entry(
							row_ptr, cols, r, col);
						if (m == UINT64_MAX) {
							/* sparse.c:213-217 */
							fprintf(stderr,
								"nbo_assemble: entry"
								" (%u,%u) missing\n",
								r, col);
							exit(1);
						}
						vals[m] += Ke[(2 * i + a) *
							      (2 * n) +
							      2 * j + c];
					}
				}
			}
			F[2 * v[i]] += Fe[2 * i];
			F[2 * v[i] + 1] += Fe[2 * i + 1];
		}
	}
	return 0;
}

/* Lumped mass vector of pipeline_assemble_system(K, M != NULL, ...).
 * Reference: solid_mechanics/pipeline.c:56-57 (M zeroed), :93-98 (density of a
 * disabled element 1e-6), :216-222 (Me[2i] += Ni*Nj*density*detJ*thickness*wp
 * for every j, inside the Gauss loop; both components get the same value),
 * :256-259 (M[2 v_i + a] += Me[2 i + a], elements in ascending id), :139-161
 * (stop at the first element with detJ < 0: status 1, M holds the elements
 * before it). */
int nbo_lumped_mass(uint32_t N_nod, const double *nod, uint32_t N_elems,
		    int elem_type, const uint32_t *adj, double density,
		    double thickness, const uint8_t *enabled, double *M)
{
	elem_t el;
	nbo_elem_tables(elem_type, &el.n, &el.ngp, el.w, el.Ni, el.dpsi,
			el.deta);
	uint32_t n = el.n;
	memset(M, 0, 2 * (size_t)N_nod * sizeof(double));
	for (uint32_t e = 0; e < N_elems; e++) {
		const uint32_t *v = adj + (size_t)n * e;
		double rho = (!enabled || enabled[e]) ? density : 1e-6;
		double Me[8], dx[4], dy[4];
		memset(Me, 0, sizeof(Me));
		for (uint32_t gp = 0; gp < el.ngp; gp++) {
			double detJ = jacobian_and_gradients(&el, nod, v, gp,
							     dx, dy);
			if (detJ < 0)
				return 1;
			double wp = el.w[gp];
			for (uint32_t i = 0; i < n; i++) {
				double Ni = el.Ni[i * el.ngp + gp];
				for (uint32_t j = 0; j < n; j++) {
					double Nj = el.Ni[j * el.ngp + gp];
					double integral = Ni * Nj * rho * detJ *
						thickness * wp;
					Me[2 * i] += integral;
					Me[2 * i + 1] += integral;
				}
			}
		}
		for (uint32_t i = 0; i < n; i++) {
			M[2 * v[i]] += Me[2 * i];
			M[2 * v[i] + 1] += Me[2 * i + 1];
		}
	}
	return 0;
}

/* Kirsch solution (infinite plate, hole radius 0.5, far-field sxx = 1e4),
 * same expression as oracle/ref_harness.c so both checkers evaluate the
 * function-valued conditions identically. */
void nbo_kirsch_stress(double x, double y, double s[3])
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

static void bc_value(const nbo_bc_t *bc, const double *xy, double val[2])
{
	if (!bc->fn) {
		val[0] = bc->val[0];
		val[1] = bc->val[1];
		return;
	}
	double s[3];
	nbo_kirsch_stress(xy[0], xy[1], s);
	if (bc->fn == 1) {
		val[0] = s[0];
		val[1] = s[2];
	} else {
		val[0] = s[2];
		val[1] = s[1];
	}
}

static double node_dist(const double *nod, uint32_t a, uint32_t b)
{
	/* mesh2D.c:650-674 */
	double dx = nod[2 * a] - nod[2 * b], dy = nod[2 * a + 1] - nod[2 * b + 1];
	return sqrt(dx * dx + dy * dy);
}

/* Reference: solid_mechanics/set_bconditions.c:52-61 and the helpers below
 * it.  Neumann on a segment with constant value is "integrated": the value is
 * the TOTAL load of the segment, each sub-segment takes the share
 * len_sub/len_sgm and gives half to each end (:133-154, :156-170); a
 * function-valued one is the trapezoid rule per sub-segment (:87-131).
 * Dirichlet conditions eliminate rows/columns one dof at a time (:218-238). */
void nbo_set_bconditions(const double *nod, const uint32_t *vtx,
			 const uint32_t *sgm_sizes, const uint32_t *sgm_nodes,
			 uint32_t N_bc, const nbo_bc_t *bc, double factor,
			 const uint64_t *row_ptr, const uint32_t *cols,
			 double *vals, double *F)
{
	static const int order[4][2] = {{1, 1}, {1, 0}, {0, 1}, {0, 0}};
	for (int pass = 0; pass < 4; pass++) {
		for (uint32_t b = 0; b < N_bc; b++) {
			const nbo_bc_t *c = bc + b;
			if (c->kind != order[pass][0] ||
			    c->where != order[pass][1])
				continue;
			const uint32_t *sn = NULL;
			uint32_t ns = 0;
			if (c->where) {
				uint64_t off = 0;
				for (uint32_t s = 0; s < c->id; s++)
					off += sgm_sizes[s];
				sn = sgm_nodes + off;
				ns = sgm_sizes[c->id];
			}
			if (c->kind == 1 && c->where == 1 && c->fn) {
				uint32_t v1 = sn[0];
				double val1[2], val2[2];
				bc_value(c, nod + 2 * v1, val1);
				for (uint32_t i = 0; i + 1 < ns; i++) {
					uint32_t v2 = sn[i + 1];
					double len = node_dist(nod, sn[i], v2);
					bc_value(c, nod + 2 * v2, val2);
					for (int j = 0; j < 2; j++) {
						if (!c->mask[j])
							continue;
						double val = 0.5 * (val1[j] +
							val2[j]) * len;
						F[2 * v1 + j] += factor * val * 0.5;
						F[2 * v2 + j] += factor * val * 0.5;
					}
					v1 = v2;
					val1[0] = val2[0];
					val1[1] = val2[1];
				}
			} else if (c->kind == 1 && c->where == 1) {
				double total = node_dist(nod, sn[0], sn[ns - 1]);
				for (uint32_t i = 0; i + 1 < ns; i++) {
					double len = node_dist(nod, sn[i],
							       sn[i + 1]);
					double share = len / total;
					double f = factor * share * 0.5;
					for (int e = 0; e < 2; e++) {
						uint32_t v = sn[i + e];
						for (int j = 0; j < 2; j++)
							if (c->mask[j])
								F[2 * v + j] +=
									f * c->val[j];
					}
				}
			} else if (c->kind == 1) {
				uint32_t v = vtx[c->id];
				for (int j = 0; j < 2; j++)
					if (c->mask[j])
						F[2 * v + j] += factor * c->val[j];
			} else {
				uint32_t cnt = c->where ? ns : 1;
				for (uint32_t i = 0; i < cnt; i++) {
					uint32_t v = c->where ? sn[i]
							      : vtx[c->id];
					double val[2];
					bc_value(c, nod + 2 * v, val);
					for (int j = 0; j < 2; j++)
						if (c->mask[j])
							nbo_dirichlet(row_ptr,
								cols, vals, F,
								2 * v + j,
								factor * val[j]);
				}
			}
		}
	}
}

/* Reference: solid_mechanics/pipeline.c:266-319.  Strain per Gauss point,
 * [exx, eyy, gxy] with gxy = du/dy + dv/dx; an element stops at its first
 * Gauss point with detJ < 0 and the remaining entries stay zero. */
int nbo_compute_strain(const double *nod, uint32_t N_elems, int elem_type,
		       const uint32_t *adj, const double *disp, double *strain)
{
	elem_t el;
	nbo_elem_tables(elem_type, &el.n, &el.ngp, el.w, el.Ni, el.dpsi,
			el.deta);
	int bad = 0;
	memset(strain, 0, (size_t)3 * el.ngp * N_elems * sizeof(double));
	for (uint32_t e = 0; e < N_elems; e++) {
		const uint32_t *v = adj + (size_t)el.n * e;
		double dx[4], dy[4];
		for (uint32_t gp = 0; gp < el.ngp; gp++) {
			double detJ = jacobian_and_gradients(&el, nod, v, gp,
							     dx, dy);
			if (detJ < 0) {
				bad = 1;
				break;
			}
			double *s = strain + 3 * ((size_t)e * el.ngp + gp);
			for (uint32_t i = 0; i < el.n; i++) {
				double ux = disp[2 * v[i]], uy = disp[2 * v[i] + 1];
				s[0] += dx[i] * ux;
				s[1] += dy[i] * uy;
				s[2] += (dy[i] * ux + dx[i] * uy);
			}
		}
	}
	return bad;
}

/* Reference: solid_mechanics/static_elasticity2D.c:99-127 */
void nbo_stress_from_strain(uint32_t N_elems, int elem_type, double E,
			    double nu, int analysis, const double *strain,
			    const uint8_t *enabled, double *stress)
{
	uint32_t ngp = elem_type ? 4 : 1;
	for (uint32_t e = 0; e < N_elems; e++) {
		double D[4] = {1e-6, 1e-6, 1e-6, 1e-6};
		if (!enabled || enabled[e])
			nbo_constitutive(E, nu, analysis, D);
		for (uint32_t gp = 0; gp < ngp; gp++) {
			const double *s = strain + 3 * ((size_t)e * ngp + gp);
			double *t = stress + 3 * ((size_t)e * ngp + gp);
			t[0] = s[0] * D[0] + s[1] * D[1];
			t[1] = s[0] * D[1] + s[1] * D[2];
			t[2] = s[2] * D[3];
		}
	}
}

/* Reference: finite_element/gaussp_to_nodes.c:50-218
 * (nb_fem_interpolate_from_gpoints_to_nodes).  Lumped-mass L2 projection of
 * Gauss-point values onto the nodes: M[v] = sum_e sum_gp sum_j Ni Nj detJ w,
 * b[v][c] = sum_e sum_gp val(e,gp,c) Ni detJ w, nodal = b / M.  Returns 1 at
 * the first element with detJ < 0 (nodal_values then stay untouched, :69-70). */
int nbo_gp_to_nodes(uint32_t N_nod, const double *nod, uint32_t N_elems,
		    int elem_type, const uint32_t *adj, uint32_t N_comp,
		    const double *gp_values, double *nodal_values)
{
	elem_t el;
	nbo_elem_tables(elem_type, &el.n, &el.ngp, el.w, el.Ni, el.dpsi,
			el.deta);
	double *M = calloc(N_nod ? N_nod : 1, sizeof(double));
	double *b = calloc((size_t)N_nod * N_comp + 1, sizeof(double));
	double *be = malloc(((size_t)4 * N_comp + 1) * sizeof(double));
	int status = 0;
	for (uint32_t e = 0; e < N_elems && !status; e++) {
		const uint32_t *v = adj + (size_t)el.n * e;
		double Me[4] = {0, 0, 0, 0};
		memset(be, 0, (size_t)4 * N_comp * sizeof(double));
		for (uint32_t gp = 0; gp < el.ngp; gp++) {
			double dx[4], dy[4];
			double detJ = jacobian_and_gradients(&el, nod, v, gp,
							     dx, dy);
			if (detJ < 0) {
				status = 1;
				break;
			}
			double wp = el.w[gp];
			for (uint32_t i = 0; i < el.n; i++) {          /* :160-176 */
				double Ni = el.Ni[i * el.ngp + gp];
				for (uint32_t j = 0; j < el.n; j++) {
					double Nj = el.Ni[j * el.ngp + gp];
					double integral = Ni * Nj * detJ * wp;
					Me[i] += integral;
				}
				size_t ggp = (size_t)e * el.ngp + gp;
				double integral = Ni * detJ * wp;
				for (uint32_t c = 0; c < N_comp; c++)
					be[i * N_comp + c] +=
						gp_values[ggp * N_comp + c] * integral;
			}
		}
		if (status)
			break;
		for (uint32_t i = 0; i < el.n; i++) {                  /* :186-196 */
			M[v[i]] += Me[i];
			for (uint32_t c = 0; c < N_comp; c++)
				b[(size_t)v[i] * N_comp + c] += be[i * N_comp + c];
		}
	}
	if (!status)
		for (size_t i = 0; i < (size_t)N_nod; i++)             /* :199-209 */
			for (uint32_t c = 0; c < N_comp; c++)
				nodal_values[i * N_comp + c] =
					b[i * N_comp + c] / M[i];
	free(M);
	free(b);
	free(be);
	return status;
}

/* Reference: common_solid_mechanics/formulas.c:65-68 */
double nbo_vm_stress(double sxx, double syy, double sxy)
{
	return sqrt(sxx * sxx + syy * syy - sxx * syy + 3.0 * (sxy * sxy));
}

/* Reference: formulas.c:70-77 -- as written there (the radius uses the MEAN
 * stress, not half the difference; restated, not corrected). */
void nbo_main_stress(double sxx, double syy, double sxy, double main_stress[2])
{
	double avg = (sxx + syy) / 2.0;
	double R = sqrt(avg * avg + sxy * sxy);
	main_stress[0] = avg + R;
	main_stress[1] = avg - R;
}
