// assembly.cu -- 2-D linear-elastic FEM stiffness assembly straight into the
// SELL-32 matrix, boundary conditions, strain/stress recovery.
//
// Reference: pipeline_assemble_system and helpers
//   (sources/nb/pde_bot/finite_element/solid_mechanics/pipeline.c:42-264),
//   nb_fem_get_jacobian / nb_fem_get_derivatives (finite_element/utils.c:9-60),
//   element tables (finite_element/element.c:54-121),
//   nb_sparse_set_Dirichlet_condition (solver_bot/sparse/sparse.c:416-430),
//   pipeline_compute_strain (pipeline.c:266-319),
//   nb_fem_compute_stress_from_strain (static_elasticity2D.c:99-127).
//
// This file is compiled with -fmad=false: every product and sum is rounded
// separately, in the reference's expression order, so element matrices come
// out bit-identical to the reference's (x86-64 without FMA).
//
// Three assembly schedules write the same values:
//   GATHER  one thread per matrix ROW.  The thread walks the elements around
//           its node in ascending element id and adds its own row of each
//           element matrix into its own SELL row: no atomics, coalesced
//           stores, and every K entry receives its contributions in the same
//           order as the reference's serial element loop -> bit-exact K and F.
//   ATOMIC  one thread per ELEMENT, full Ke in registers, scattered with
//           atomicAdd(double) (RED.ADD.F64); order is nondeterministic, values
//           agree to rounding.
//   COLOR   one thread per element, one launch per colour of a greedy
//           element colouring (no two elements of a colour share a node), plain
//           read-modify-write; deterministic, colour order instead of id order.
// The reference has no colouring code; the greedy colouring here
// is this library's own.
#include <algorithm>
#include <cstring>

#include "matrix.cuh"
#include "mesh.cuh"

using namespace nbgpu;

namespace {

struct ElemTables {
	double w[4], Ni[16], dpsi[16], deta[16];
};

struct AsmParams {
	double D[4], D_void[4];
	double density, density_void, thickness;
	double gx, gy;
	int self_weight;
};

__constant__ ElemTables c_tab;

// Jacobian and Cartesian gradients at Gauss point gp (utils.c:9-41, :49-60).
template <int NPE, int NGP>
__device__ __forceinline__ double jacobian_gradients(const double (&xs)[NPE], const double (&ys)[NPE],
						      int gp, double (&dNdx)[NPE], double (&dNdy)[NPE])
{
	double x_psi = 0.0, y_psi = 0.0, x_eta = 0.0, y_eta = 0.0;
#pragma unroll
	for (int i = 0; i < NPE; i++) {
		const double dp = c_tab.dpsi[i * NGP + gp], de = c_tab.deta[i * NGP + gp];
		x_psi += dp * xs[i];
		x_eta += de * xs[i];
		y_psi += dp * ys[i];
		y_eta += de * ys[i];
	}
	const double detJ = x_psi * y_eta - y_psi * x_eta;
	const double j0 = y_eta / detJ, j1 = -y_psi / detJ, j2 = -x_eta / detJ, j3 = x_psi / detJ;
#pragma unroll
	for (int i = 0; i < NPE; i++) {
		const double dp = c_tab.dpsi[i * NGP + gp], de = c_tab.deta[i * NGP + gp];
		dNdx[i] = j0 * dp + j1 * de;
		dNdy[i] = j2 * dp + j3 * de;
	}
	return detJ;
}

__device__ __forceinline__ void element_material(const AsmParams &P, const uint8_t *enabled,
						 const double *scale, uint32_t e, double (&D)[4],
						 double &rho)
{
	if (!enabled || enabled[e]) {
		// pipeline.c:95-98; the optional factor scales the enabled material (SIMP-style)
		const double s = scale ? scale[e] : 1.0;
#pragma unroll
		for (int k = 0; k < 4; k++)
			D[k] = scale ? P.D[k] * s : P.D[k];
		rho = P.density;
	} else {
		// pipeline.c:93-94: "void" material, including the off-diagonal 1e-6
#pragma unroll
		for (int k = 0; k < 4; k++)
			D[k] = P.D_void[k];
		rho = P.density_void;
	}
}

// position of column `c` in SELL row `row` (entries j < width), or SIZE_MAX
__device__ __forceinline__ size_t find_in_row(const uint32_t *__restrict__ col, uint32_t off,
					      uint32_t width, uint32_t lane, uint32_t c)
{
	const uint32_t *p = col + (size_t)off * kSliceRows + lane;
	for (uint32_t j = 0; j < width; j++) {
		const uint32_t cj = p[(size_t)j * kSliceRows];
		if (cj == c)
			return ((size_t)off + j) * kSliceRows + lane;
		if (cj > c)   // ascending columns; padding is 0xFFFFFFFF
			break;
	}
	return (size_t)-1;
}

// ---- GATHER: one thread per matrix row -------------------------------------
template <int NPE, int NGP>
__global__ void __launch_bounds__(kBlock)
assemble_gather_kernel(uint32_t n_rows, uint32_t col_shift, const double *__restrict__ nod,
		       const uint32_t *__restrict__ adj,
		       const uint32_t *__restrict__ n2e_ptr, const uint32_t *__restrict__ n2e,
		       const uint8_t *__restrict__ enabled, const double *__restrict__ scale, AsmParams P,
		       const uint32_t *__restrict__ slice_off, const uint32_t *__restrict__ perm,
		       const uint32_t *__restrict__ col, double *__restrict__ val, double *__restrict__ F,
		       unsigned int *first_bad, int *pattern_miss)
{
	// thread = storage position (slice * 32 + lane), so that the value stores stay coalesced
	// whatever the row order of the layout is
	// Rank-local block (dist_fem.cu): the mesh is the sub-mesh numbered like the block's column
	// space, row r belongs to node (r + col_shift) / 2 of it; otherwise col_shift = 0.
	const uint32_t spos = blockIdx.x * blockDim.x + threadIdx.x;
	if (spos >= ((n_rows + kSliceRows - 1) / kSliceRows) * kSliceRows)
		return;
	const uint32_t row = perm ? perm[spos] : spos;
	if (row >= n_rows)
		return;
	const uint32_t node = (row + col_shift) >> 1, a = row & 1, lane = spos & 31;
	const uint32_t off = slice_off[spos >> 5], width = slice_off[(spos >> 5) + 1] - off;
	double f_acc = 0.0;
	for (uint32_t t = n2e_ptr[node]; t < n2e_ptr[node + 1]; t++) {
		const uint32_t e = n2e[t];
		uint32_t v[NPE];
		double xs[NPE], ys[NPE];
		int li = 0;
#pragma unroll
		for (int i = 0; i < NPE; i++) {
			v[i] = adj[(size_t)e * NPE + i];
			xs[i] = nod[2 * (size_t)v[i]];
			ys[i] = nod[2 * (size_t)v[i] + 1];
		}
#pragma unroll
		for (int i = NPE - 1; i >= 0; i--)
			if (v[i] == node)
				li = i;   // first local index of this node
		double D[4], rho;
		element_material(P, enabled, scale, e, D, rho);
		const double fa = P.self_weight ? (a ? P.gy : P.gx) * rho : 0.0;
		double kv[2 * NPE];   // row (2 li + a) of Ke
#pragma unroll
		for (int c = 0; c < 2 * NPE; c++)
			kv[c] = 0.0;
		double fe = 0.0;
		bool bad = false;
#pragma unroll
		for (int gp = 0; gp < NGP; gp++) {
			double dx[NPE], dy[NPE];
			const double detJ = jacobian_gradients<NPE, NGP>(xs, ys, gp, dx, dy);
			if (detJ < 0)
				bad = true;   // utils.c:44-47
			const double wp = c_tab.w[gp];
			double dxi = dx[0], dyi = dy[0], Ni = c_tab.Ni[gp];
#pragma unroll
			for (int i = 1; i < NPE; i++)
				if (li == i) {
					dxi = dx[i];
					dyi = dy[i];
					Ni = c_tab.Ni[i * NGP + gp];
				}
#pragma unroll
			for (int j = 0; j < NPE; j++) {
				// pipeline.c:196-214, row a of the 2x2 block (li, j)
				if (a == 0) {
					kv[2 * j] += (dxi * dx[j] * D[0] + dyi * dy[j] * D[3]) * detJ *
						     P.thickness * wp;
					kv[2 * j + 1] += (dxi * dy[j] * D[1] + dyi * dx[j] * D[3]) * detJ *
							 P.thickness * wp;
				} else {
					kv[2 * j] += (dyi * dx[j] * D[1] + dxi * dy[j] * D[3]) * detJ *
						     P.thickness * wp;
					kv[2 * j + 1] += (dyi * dy[j] * D[2] + dxi * dx[j] * D[3]) * detJ *
							 P.thickness * wp;
				}
			}
			const double integral = Ni * detJ * P.thickness * wp;   // pipeline.c:225-228
			fe += integral * fa;
		}
		if (bad) {
			atomicMin(first_bad, e);
			continue;
		}
#pragma unroll
		for (int j = 0; j < NPE; j++) {
#pragma unroll
			for (int b = 0; b < 2; b++) {
				const size_t pos = find_in_row(col, off, width, lane, 2 * v[j] + b);
				if (pos == (size_t)-1) {
					*pattern_miss = 1;   // sparse.c:213-217
					continue;
				}
				val[pos] += kv[2 * j + b];
			}
		}
		f_acc += fe;
	}
	F[row] = f_acc;
}

// ---- GATHER, node-parallel: one thread per NODE (both of its rows) ------------------------
// For matrices in the 2x2-blocked layout (every mesh-built FEM matrix).  Against the row-parallel
// kernel above: the Jacobians and gradients of an element are shared by the node's two rows (4x
// instead of 8x redundant integration for quads), the node's block ids (one per 2x2 block, read once,
// coalesced, from bcol) replace 32 linear searches through the per-entry column array per element,
// and the two rows are accumulated in shared memory and written ONCE -- no memset of K, no
// read-modify-write per contributing element.  Every entry still receives its contributions from 0
// upwards in ascending element id, each of them summed over the Gauss points in order, so K and F
// keep the reference's bits (pipeline.c:174-264).
constexpr int kNodeThreads = 128;   // 8 slices per CTA
constexpr int kNodeMaxBlocks = 24;  // 2x2 blocks per node row held in (dynamic) shared memory: 36 bytes per block and thread

// DAMAGE: the damage driver's loop (static_damage2D.c:474-569) -- the constitutive matrix of Gauss point gp of
// element e is multiplied by (1 - gp_damage[e NGP + gp]) (:530-536), for the void material of a disabled
// element as well.  A separate instantiation: the plain kernel's code is unchanged.
template <int NPE, int NGP, bool DAMAGE = false>
__global__ void __launch_bounds__(kNodeThreads)
assemble_node_kernel(uint32_t n_rows, uint32_t col_shift, uint32_t nb_max, const double *__restrict__ nod,
		     const uint32_t *__restrict__ adj, const uint32_t *__restrict__ n2e_ptr,
		     const uint32_t *__restrict__ n2e, const uint8_t *__restrict__ enabled,
		     const double *__restrict__ scale, const double *__restrict__ gp_damage, AsmParams P,
		     const uint32_t *__restrict__ slice_off,
		     const uint32_t *__restrict__ perm, const uint32_t *__restrict__ bcol, double *__restrict__ val,
		     double *__restrict__ F, unsigned int *first_bad, int *pattern_miss)
{
	// dynamic shared memory: double acc[4 nb_max][threads] | uint32 id[nb_max][threads]
	extern __shared__ __align__(16) unsigned char node_smem[];
	double(*s_acc)[kNodeThreads] = reinterpret_cast<double(*)[kNodeThreads]>(node_smem);
	uint32_t(*s_id)[kNodeThreads] =
		reinterpret_cast<uint32_t(*)[kNodeThreads]>(node_smem + (size_t)4 * nb_max * kNodeThreads * sizeof(double));
	const uint32_t tid = threadIdx.x;
	const uint32_t pair = blockIdx.x * kNodeThreads + tid;   // storage pair: slice * 16 + nl
	const uint32_t slice = pair >> 4, nl = pair & 15;
	const uint32_t spos = slice * kSliceRows + 2 * nl;
	if (spos >= ((n_rows + kSliceRows - 1) / kSliceRows) * kSliceRows)
		return;
	const uint32_t row0 = perm ? perm[spos] : spos;
	if (row0 >= n_rows)
		return;
	const uint32_t node = (row0 + col_shift) >> 1;
	const uint32_t off = slice_off[slice], nb = (slice_off[slice + 1] - off) >> 1;
	for (uint32_t jb = 0; jb < nb; jb++) {
		s_id[jb][tid] = bcol[((size_t)(off >> 1) + jb) * 16u + nl];
#pragma unroll
		for (int q = 0; q < 4; q++)
			s_acc[jb * 4 + q][tid] = 0.0;
	}
	double f0 = 0.0, f1 = 0.0;
	for (uint32_t t = n2e_ptr[node]; t < n2e_ptr[node + 1]; t++) {
		const uint32_t e = n2e[t];
		uint32_t v[NPE];
		double xs[NPE], ys[NPE];
		int li = 0;
#pragma unroll
		for (int i = 0; i < NPE; i++) {
			v[i] = adj[(size_t)e * NPE + i];
			xs[i] = nod[2 * (size_t)v[i]];
			ys[i] = nod[2 * (size_t)v[i] + 1];
		}
#pragma unroll
		for (int i = NPE - 1; i >= 0; i--)
			if (v[i] == node)
				li = i;   // first local index of this node
		double D0[4], rho;
		element_material(P, enabled, scale, e, D0, rho);
		const double fx = P.self_weight ? P.gx * rho : 0.0, fy = P.self_weight ? P.gy * rho : 0.0;
		double k0[2 * NPE], k1[2 * NPE];   // rows (2 li) and (2 li + 1) of Ke
#pragma unroll
		for (int c = 0; c < 2 * NPE; c++)
			k0[c] = k1[c] = 0.0;
		double fe0 = 0.0, fe1 = 0.0;
		bool bad = false;
#pragma unroll
		for (int gp = 0; gp < NGP; gp++) {
			double dx[NPE], dy[NPE];
			const double detJ = jacobian_gradients<NPE, NGP>(xs, ys, gp, dx, dy);
			if (detJ < 0)
				bad = true;   // utils.c:44-47
			const double wp = c_tab.w[gp];
			double D[4];
			const double undamaged = DAMAGE ? 1.0 - gp_damage[(size_t)e * NGP + gp] : 1.0;
#pragma unroll
			for (int k = 0; k < 4; k++)
				D[k] = DAMAGE ? D0[k] * undamaged : D0[k];
			double dxi = dx[0], dyi = dy[0], Ni = c_tab.Ni[gp];
#pragma unroll
			for (int i = 1; i < NPE; i++)
				if (li == i) {
					dxi = dx[i];
					dyi = dy[i];
					Ni = c_tab.Ni[i * NGP + gp];
				}
#pragma unroll
			for (int j = 0; j < NPE; j++) {
				// pipeline.c:196-214, the 2x2 block (li, j)
				k0[2 * j] += (dxi * dx[j] * D[0] + dyi * dy[j] * D[3]) * detJ * P.thickness * wp;
				k0[2 * j + 1] += (dxi * dy[j] * D[1] + dyi * dx[j] * D[3]) * detJ * P.thickness * wp;
				k1[2 * j] += (dyi * dx[j] * D[1] + dxi * dy[j] * D[3]) * detJ * P.thickness * wp;
				k1[2 * j + 1] += (dyi * dy[j] * D[2] + dxi * dx[j] * D[3]) * detJ * P.thickness * wp;
			}
			const double integral = Ni * detJ * P.thickness * wp;   // pipeline.c:225-228
			fe0 += integral * fx;
			fe1 += integral * fy;
		}
		if (bad) {
			atomicMin(first_bad, e);
			continue;
		}
#pragma unroll
		for (int j = 0; j < NPE; j++) {
			uint32_t jb = 0;
			while (jb < nb && s_id[jb][tid] != v[j])
				jb++;
			if (jb == nb) {
				*pattern_miss = 1;   // sparse.c:213-217
				continue;
			}
			s_acc[jb * 4 + 0][tid] += k0[2 * j];
			s_acc[jb * 4 + 1][tid] += k0[2 * j + 1];
			s_acc[jb * 4 + 2][tid] += k1[2 * j];
			s_acc[jb * 4 + 3][tid] += k1[2 * j + 1];
		}
		f0 += fe0;
		f1 += fe1;
	}
	const uint32_t lane = spos & 31;
	for (uint32_t jb = 0; jb < nb; jb++) {
		double *p = val + ((size_t)off + 2 * jb) * kSliceRows + lane;
		p[0] = s_acc[jb * 4 + 0][tid];
		p[kSliceRows] = s_acc[jb * 4 + 1][tid];
		p[1] = s_acc[jb * 4 + 2][tid];
		p[kSliceRows + 1] = s_acc[jb * 4 + 3][tid];
	}
	F[row0] = f0;
	F[row0 + 1] = f1;
}

// ---- lumped mass vector: one thread per node ---------------------------------
// What pipeline_assemble_system fills when M != NULL (pipeline.c:56-57, :216-222, :256-259): for every element
// of the node in ascending id, Me[2i] accumulates Ni*Nj*density*detJ*thickness*wp over j inside the Gauss loop
// (that expression order), and M[2v+a] += Me[2i+a].  Same gather schedule as the stiffness rows, so bit-exact.
template <int NPE, int NGP>
__global__ void __launch_bounds__(kBlock)
lumped_mass_kernel(uint32_t N_nod, const double *__restrict__ nod, const uint32_t *__restrict__ adj,
		   const uint32_t *__restrict__ n2e_ptr, const uint32_t *__restrict__ n2e,
		   const uint8_t *__restrict__ enabled, double density, double density_void, double thickness,
		   double *__restrict__ M, unsigned int *first_bad)
{
	const uint32_t node = blockIdx.x * blockDim.x + threadIdx.x;
	if (node >= N_nod)
		return;
	double m = 0.0;
	for (uint32_t t = n2e_ptr[node]; t < n2e_ptr[node + 1]; t++) {
		const uint32_t e = n2e[t];
		uint32_t v[NPE];
		double xs[NPE], ys[NPE];
		int li = 0;
#pragma unroll
		for (int i = 0; i < NPE; i++) {
			v[i] = adj[(size_t)e * NPE + i];
			xs[i] = nod[2 * (size_t)v[i]];
			ys[i] = nod[2 * (size_t)v[i] + 1];
		}
#pragma unroll
		for (int i = NPE - 1; i >= 0; i--)
			if (v[i] == node)
				li = i;   // first local index of this node
		const double rho = (!enabled || enabled[e]) ? density : density_void;   // pipeline.c:93-98
		double me = 0.0;
		bool bad = false;
#pragma unroll
		for (int gp = 0; gp < NGP; gp++) {
			double dx[NPE], dy[NPE];
			const double detJ = jacobian_gradients<NPE, NGP>(xs, ys, gp, dx, dy);
			if (detJ < 0)
				bad = true;   // utils.c:44-47
			const double wp = c_tab.w[gp];
			double Ni = c_tab.Ni[gp];
#pragma unroll
			for (int i = 1; i < NPE; i++)
				if (li == i)
					Ni = c_tab.Ni[i * NGP + gp];
#pragma unroll
			for (int j = 0; j < NPE; j++) {
				const double Nj = c_tab.Ni[j * NGP + gp];
				me += Ni * Nj * rho * detJ * thickness * wp;   // pipeline.c:216-221
			}
		}
		if (bad) {
			atomicMin(first_bad, e);
			continue;
		}
		m += me;
	}
	M[2 * (size_t)node] = m;
	M[2 * (size_t)node + 1] = m;
}

// ---- ATOMIC / COLOR: one thread per element --------------------------------
template <int NPE, int NGP, bool ATOMIC>
__global__ void __launch_bounds__(128)
assemble_element_kernel(uint32_t n_work, const uint32_t *__restrict__ work /* null = identity */,
			const double *__restrict__ nod, const uint32_t *__restrict__ adj,
			const uint8_t *__restrict__ enabled, const double *__restrict__ scale, AsmParams P,
			const uint32_t *__restrict__ slice_off, const uint32_t *__restrict__ inv_perm,
			const uint32_t *__restrict__ col, double *__restrict__ val, double *__restrict__ F,
			unsigned int *first_bad, int *pattern_miss)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_work)
		return;
	const uint32_t e = work ? work[t] : t;
	uint32_t v[NPE];
	double xs[NPE], ys[NPE];
#pragma unroll
	for (int i = 0; i < NPE; i++) {
		v[i] = adj[(size_t)e * NPE + i];
		xs[i] = nod[2 * (size_t)v[i]];
		ys[i] = nod[2 * (size_t)v[i] + 1];
	}
	double D[4], rho;
	element_material(P, enabled, scale, e, D, rho);
	const double fx = P.self_weight ? P.gx * rho : 0.0, fy = P.self_weight ? P.gy * rho : 0.0;
	double Ke[4 * NPE * NPE], Fe[2 * NPE];
#pragma unroll
	for (int k = 0; k < 4 * NPE * NPE; k++)
		Ke[k] = 0.0;
#pragma unroll
	for (int k = 0; k < 2 * NPE; k++)
		Fe[k] = 0.0;
	bool bad = false;
#pragma unroll
	for (int gp = 0; gp < NGP; gp++) {
		double dx[NPE], dy[NPE];
		const double detJ = jacobian_gradients<NPE, NGP>(xs, ys, gp, dx, dy);
		if (detJ < 0)
			bad = true;
		const double wp = c_tab.w[gp];
#pragma unroll
		for (int i = 0; i < NPE; i++) {
#pragma unroll
			for (int j = 0; j < NPE; j++) {
				Ke[(2 * i) * (2 * NPE) + 2 * j] +=
					(dx[i] * dx[j] * D[0] + dy[i] * dy[j] * D[3]) * detJ * P.thickness * wp;
				Ke[(2 * i) * (2 * NPE) + 2 * j + 1] +=
					(dx[i] * dy[j] * D[1] + dy[i] * dx[j] * D[3]) * detJ * P.thickness * wp;
				Ke[(2 * i + 1) * (2 * NPE) + 2 * j] +=
					(dy[i] * dx[j] * D[1] + dx[i] * dy[j] * D[3]) * detJ * P.thickness * wp;
				Ke[(2 * i + 1) * (2 * NPE) + 2 * j + 1] +=
					(dy[i] * dy[j] * D[2] + dx[i] * dx[j] * D[3]) * detJ * P.thickness * wp;
			}
			const double integral = c_tab.Ni[i * NGP + gp] * detJ * P.thickness * wp;
			Fe[2 * i] += integral * fx;
			Fe[2 * i + 1] += integral * fy;
		}
	}
	if (bad) {
		atomicMin(first_bad, e);
		return;
	}
#pragma unroll
	for (int i = 0; i < NPE; i++) {
#pragma unroll
		for (int a = 0; a < 2; a++) {
			const uint32_t row = 2 * v[i] + a;
			const uint32_t spos = inv_perm ? inv_perm[row] : row;
			const uint32_t off = slice_off[spos >> 5], width = slice_off[(spos >> 5) + 1] - off;
#pragma unroll
			for (int j = 0; j < NPE; j++) {
#pragma unroll
				for (int b = 0; b < 2; b++) {
					const size_t pos = find_in_row(col, off, width, spos & 31, 2 * v[j] + b);
					if (pos == (size_t)-1) {
						*pattern_miss = 1;
						continue;
					}
					const double k = Ke[(2 * i + a) * (2 * NPE) + 2 * j + b];
					if (ATOMIC)
						atomicAdd(val + pos, k);
					else
						val[pos] += k;
				}
			}
			if (P.self_weight) {
				if (ATOMIC)
					atomicAdd(F + row, Fe[2 * i + a]);
				else
					F[row] += Fe[2 * i + a];
			}
		}
	}
}

// ---- boundary conditions ---------------------------------------------------

// F[dof[k]] += add[k] in list order (set_bconditions.c:156-170).  Entries of different dofs commute; the
// entries of ONE dof are added by one thread in list order: thread k owns dof[k] if no earlier entry names
// the same dof, and then walks the rest of the list.  The lists are boundary-sized (thousands of entries),
// so the quadratic scan is a few microseconds -- the single-thread loop this replaces took 215 us for the
// 4000 entries of the Q1 workload, as long as the assembly itself.
__global__ void __launch_bounds__(256)
vector_add_entries_kernel(double *F, uint32_t n, const uint32_t *__restrict__ dof, const double *__restrict__ add)
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n)
		return;
	const uint32_t d = dof[k];
	for (uint32_t j = 0; j < k; j++)
		if (dof[j] == d)
			return;
	double acc = F[d];
	for (uint32_t j = k; j < n; j++)
		if (dof[j] == d)
			acc += add[j];
	F[d] = acc;
}

// nb_sparse_set_Dirichlet_condition applied for a whole ordered list at once.
// order[i] = 1 + position of dof i's FIRST occurrence in the list (0 = free),
// v_first / v_last = its first / last prescribed value.  Sequential semantics
// reproduced per row (sparse.c:416-430):
//   constrained row i : identity row, F[i] = last value written for i;
//   free row i        : every entry (i,c) with c constrained is zeroed and
//                       F[i] -= A_ic * v_first[c], subtracted in the order the
//                       constraints were applied (ascending order[c]).
__global__ void __launch_bounds__(kBlock)
dirichlet_kernel(uint32_t N, uint32_t col_shift, uint32_t n_pos, const uint32_t *__restrict__ slice_off,
		 const uint32_t *__restrict__ perm, const uint32_t *__restrict__ col,
		 double *__restrict__ val, double *__restrict__ F, const uint32_t *__restrict__ order,
		 const double *__restrict__ v_first, const double *__restrict__ v_last)
{
	const uint32_t spos = blockIdx.x * blockDim.x + threadIdx.x;   // storage position
	if (spos >= n_pos)
		return;
	const uint32_t row = perm ? perm[spos] : spos;
	if (row >= N)
		return;
	const uint32_t lane = spos & 31;
	const uint32_t off = slice_off[spos >> 5], width = slice_off[(spos >> 5) + 1] - off;
	const uint32_t *cp = col + (size_t)off * kSliceRows + lane;
	double *vp = val + (size_t)off * kSliceRows + lane;
	// `order` is indexed by COLUMN (rank-local block: row r is column r + col_shift)
	const uint32_t my = order[row + col_shift];
	if (my) {
		for (uint32_t j = 0; j < width; j++) {
			const uint32_t c = cp[(size_t)j * kSliceRows];
			if (c == kPadCol)
				break;
			vp[(size_t)j * kSliceRows] = (c == row + col_shift) ? 1.0 : 0.0;
		}
		F[row] = v_last[my - 1];
		return;
	}
	uint32_t last = 0;
	double f = F[row];
	bool touched = false;
	for (;;) {
		uint32_t best = 0xFFFFFFFFu, best_j = 0;
		for (uint32_t j = 0; j < width; j++) {
			const uint32_t c = cp[(size_t)j * kSliceRows];
			if (c == kPadCol)
				break;
			const uint32_t o = order[c];
			if (o > last && o < best) {
				best = o;
				best_j = j;
			}
		}
		if (best == 0xFFFFFFFFu)
			break;
		const double a = vp[(size_t)best_j * kSliceRows];
		vp[(size_t)best_j * kSliceRows] = 0.0;
		f -= a * v_first[best - 1];
		touched = true;
		last = best;
	}
	if (touched)
		F[row] = f;
}

// ---- strain / stress -------------------------------------------------------
template <int NPE, int NGP>
__global__ void __launch_bounds__(kBlock)
strain_kernel(uint32_t N_elems, const double *__restrict__ nod, const uint32_t *__restrict__ adj,
	      const double *__restrict__ disp, double *__restrict__ strain)
{
	const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= N_elems)
		return;
	uint32_t v[NPE];
	double xs[NPE], ys[NPE], ux[NPE], uy[NPE];
#pragma unroll
	for (int i = 0; i < NPE; i++) {
		v[i] = adj[(size_t)e * NPE + i];
		xs[i] = nod[2 * (size_t)v[i]];
		ys[i] = nod[2 * (size_t)v[i] + 1];
		ux[i] = disp[2 * (size_t)v[i]];
		uy[i] = disp[2 * (size_t)v[i] + 1];
	}
	bool stop = false;
#pragma unroll
	for (int gp = 0; gp < NGP; gp++) {
		double dx[NPE], dy[NPE];
		const double detJ = jacobian_gradients<NPE, NGP>(xs, ys, gp, dx, dy);
		if (detJ < 0)
			stop = true;   // pipeline.c:302-303: the element stops at its first bad point
		double s0 = 0.0, s1 = 0.0, s2 = 0.0;
		if (!stop) {
#pragma unroll
			for (int i = 0; i < NPE; i++) {
				s0 += dx[i] * ux[i];
				s1 += dy[i] * uy[i];
				s2 += (dy[i] * ux[i] + dx[i] * uy[i]);
			}
		}
		double *s = strain + 3 * ((size_t)e * NGP + gp);
		s[0] = s0;
		s[1] = s1;
		s[2] = s2;
	}
}

__global__ void __launch_bounds__(kBlock)
stress_kernel(uint32_t N_elems, uint32_t NGP, const uint8_t *__restrict__ enabled, double D0, double D1,
	      double D2, double D3, double V0, double V1, double V2, double V3,
	      const double *__restrict__ strain, double *__restrict__ stress)
{
	const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (id >= (size_t)N_elems * NGP)
		return;
	const uint32_t e = (uint32_t)(id / NGP);
	const bool en = !enabled || enabled[e];
	const double d0 = en ? D0 : V0, d1 = en ? D1 : V1, d2 = en ? D2 : V2, d3 = en ? D3 : V3;
	const double *s = strain + 3 * id;
	double *t = stress + 3 * id;
	const double s0 = s[0], s1 = s[1], s2 = s[2];
	t[0] = s0 * d0 + s1 * d1;
	t[1] = s0 * d1 + s1 * d2;
	t[2] = s2 * d3;
}

// ---- Gauss points -> nodes (gaussp_to_nodes.c:50-218) ------------------------------
// Lumped-mass L2 projection.  One thread per (node, component): it walks the elements around its
// node in ascending element id -- the order in which the reference's serial element loop adds into
// M[v] and b[v][c] (:186-196) -- so the quotient b / M comes out bit-identical.
template <int NPE, int NGP>
__global__ void __launch_bounds__(kBlock)
gp_to_nodes_kernel(uint32_t N_nod, uint32_t N_comp, const double *__restrict__ nod,
		   const uint32_t *__restrict__ adj, const uint32_t *__restrict__ n2e_ptr,
		   const uint32_t *__restrict__ n2e, const double *__restrict__ gp_values,
		   double *__restrict__ nodal, unsigned int *first_bad)
{
	const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= (size_t)N_nod * N_comp)
		return;
	const uint32_t node = (uint32_t)(t / N_comp), c = (uint32_t)(t % N_comp);
	double M = 0.0, b = 0.0;
	for (uint32_t k = n2e_ptr[node]; k < n2e_ptr[node + 1]; k++) {
		const uint32_t e = n2e[k];
		uint32_t v[NPE];
		double xs[NPE], ys[NPE];
#pragma unroll
		for (int i = 0; i < NPE; i++) {
			v[i] = adj[(size_t)e * NPE + i];
			xs[i] = nod[2 * (size_t)v[i]];
			ys[i] = nod[2 * (size_t)v[i] + 1];
		}
		// a node listed twice by an element receives both local rows, as in the reference
#pragma unroll
		for (int i = 0; i < NPE; i++) {
			if (v[i] != node)
				continue;
			double Me = 0.0, be = 0.0;
			bool bad = false;
#pragma unroll
			for (int gp = 0; gp < NGP; gp++) {
				double dx[NPE], dy[NPE];
				const double detJ = jacobian_gradients<NPE, NGP>(xs, ys, gp, dx, dy);
				if (detJ < 0)
					bad = true;
				const double wp = c_tab.w[gp];
				const double Ni = c_tab.Ni[i * NGP + gp];
#pragma unroll
				for (int j = 0; j < NPE; j++)
					Me += Ni * c_tab.Ni[j * NGP + gp] * detJ * wp;      // :163-167
				const double integral = Ni * detJ * wp;                     // :171
				be += gp_values[((size_t)e * NGP + gp) * N_comp + c] * integral;
			}
			if (bad) {
				atomicMin(first_bad, e);
				continue;
			}
			M += Me;
			b += be;
		}
	}
	nodal[t] = b / M;
}

__global__ void __launch_bounds__(kBlock)
von_mises_kernel(size_t n, const double *__restrict__ stress, double *__restrict__ vm)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const double sxx = stress[3 * i], syy = stress[3 * i + 1], sxy = stress[3 * i + 2];
	vm[i] = sqrt(sxx * sxx + syy * syy - sxx * syy + 3.0 * (sxy * sxy));   // formulas.c:65-68
}

__global__ void __launch_bounds__(kBlock)
main_stress_kernel(size_t n, const double *__restrict__ stress, double *__restrict__ out)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const double sxx = stress[3 * i], syy = stress[3 * i + 1], sxy = stress[3 * i + 2];
	// formulas.c:70-77 exactly as the reference writes it (radius from the MEAN stress)
	const double avg = (sxx + syy) / 2.0;
	const double R = sqrt(avg * avg + sxy * sxy);
	out[2 * i] = avg + R;
	out[2 * i + 1] = avg - R;
}

// ---- host helpers ------------------------------------------------------------

int upload_tables(const nbgpu_elem_tables_t *t, uint32_t npe)
{
	NB_ARG(t != nullptr && t->N_nodes == npe);
	NB_ARG((npe == 3 && t->N_gp == 1) || (npe == 4 && t->N_gp == 4));
	ElemTables h;
	memcpy(h.w, t->gp_weight, sizeof(h.w));
	memcpy(h.Ni, t->Ni, sizeof(h.Ni));
	memcpy(h.dpsi, t->dNi_dpsi, sizeof(h.dpsi));
	memcpy(h.deta, t->dNi_deta, sizeof(h.deta));
	NB_CUDA(cudaMemcpyToSymbolAsync(c_tab, &h, sizeof(h), 0, cudaMemcpyHostToDevice, ctx().stream));
	NB_CUDA(cudaStreamSynchronize(ctx().stream));   // `h` is a stack object
	return NBGPU_OK;
}

// ---- element colouring on the device (the reference has no colouring code) ----------
// Jones-Plassmann rounds: an uncoloured element whose priority (a hash of its id, ties by id) is the largest
// among its uncoloured neighbours -- the elements it shares a node with, found through the node -> element
// lists -- takes the lowest colour none of its coloured neighbours uses.  Two neighbours can never both be
// local maxima, so no two elements of a colour share a node.  A dozen rounds at a few microseconds each;
// the host only reads the count of elements still uncoloured.
__device__ __forceinline__ uint32_t color_priority(uint32_t e)
{
	uint32_t h = e * 2654435761u;
	h ^= h >> 15;
	h *= 2246822519u;
	h ^= h >> 13;
	return h;
}

__global__ void __launch_bounds__(256)
color_round_kernel(uint32_t N_elems, uint32_t npe, const uint32_t *__restrict__ adj, const uint32_t *__restrict__ n2e_ptr,
		   const uint32_t *__restrict__ n2e, const uint8_t *__restrict__ color, uint8_t *__restrict__ color_out,
		   unsigned int *remaining, int *overflow)
{
	// decisions are taken on the colours of the PREVIOUS round only (color), so the result does not depend
	// on the order in which the threads of a round run
	const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= N_elems)
		return;
	color_out[e] = color[e];
	if (color[e] != 0xFF)
		return;
	const uint32_t pe = color_priority(e);
	unsigned long long used = 0;
	bool is_max = true;
	for (uint32_t i = 0; i < npe && is_max; i++) {
		const uint32_t v = adj[(size_t)e * npe + i];
		for (uint32_t t = n2e_ptr[v]; t < n2e_ptr[v + 1]; t++) {
			const uint32_t f = n2e[t];
			if (f == e)
				continue;
			const uint8_t cf = color[f];
			if (cf == 0xFF) {
				const uint32_t pf = color_priority(f);
				if (pf > pe || (pf == pe && f > e)) {
					is_max = false;
					break;
				}
			} else {
				used |= 1ull << cf;
			}
		}
	}
	if (!is_max) {
		atomicAdd(remaining, 1u);
		return;
	}
	if (~used == 0) {
		*overflow = 1;
		return;
	}
	color_out[e] = (uint8_t)__ffsll((long long)~used) - 1;
}

__global__ void __launch_bounds__(256)
color_count_kernel(uint32_t N_elems, const uint8_t *__restrict__ color, unsigned int *counts)
{
	const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < N_elems)
		atomicAdd(counts + color[e], 1u);
}

// elements of one colour never touch the same entry, so their order inside the colour's list does not matter
__global__ void __launch_bounds__(256)
color_scatter_kernel(uint32_t N_elems, const uint8_t *__restrict__ color, unsigned int *next, uint32_t *__restrict__ elems)
{
	const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < N_elems)
		elems[atomicAdd(next + color[e], 1u)] = e;
}

int build_coloring(nbgpu_mesh_t *m)
{
	if (m->n_colors)
		return NBGPU_OK;
	Context &c = ctx();
	uint8_t *d_color = nullptr, *d_color2 = nullptr;
	unsigned int *d_cnt = nullptr;   // [0] remaining, [1] overflow, [2..66) per-colour counters
	NB_CUDA(nbgpu::dmalloc(&d_color, 2 * std::max<size_t>(1, m->N_elems)));
	d_color2 = d_color + std::max<size_t>(1, m->N_elems);
	cudaError_t e = nbgpu::dmalloc(&d_cnt, 66 * sizeof(unsigned int));
	if (e != cudaSuccess) {
		nbgpu::dfree(d_color);
		NB_CUDA(e);
	}
	uint8_t *const d_color_block = d_color;
	auto done = [&](int st) {
		if (st != NBGPU_OK)
			nbgpu::dfree(d_color_block);
		nbgpu::dfree(d_cnt);
		return st;
	};
	cudaMemsetAsync(d_color, 0xFF, m->N_elems, c.stream);
	const int grid = (int)((m->N_elems + 255) / 256);
	unsigned int h_cnt[66];
	for (int round = 0;; round++) {
		cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned int), c.stream);
		color_round_kernel<<<grid, 256, 0, c.stream>>>(m->N_elems, m->npe, m->d_adj, m->d_n2e_ptr, m->d_n2e, d_color, d_color2,
							       d_cnt, (int *)(d_cnt + 1));
		c.launches++;
		std::swap(d_color, d_color2);
		e = cudaMemcpyAsync(h_cnt, d_cnt, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c.stream);
		if (e == cudaSuccess)
			e = cudaStreamSynchronize(c.stream);
		if (e != cudaSuccess) {
			set_error("colouring: %s", cudaGetErrorString(e));
			return done(NBGPU_ERR_CUDA);
		}
		if (h_cnt[1] || round > 4096) {
			set_error("element colouring needs more than 64 colours");
			return done(NBGPU_ERR_ARG);
		}
		if (h_cnt[0] == 0)
			break;
	}
	cudaMemsetAsync(d_cnt, 0, 66 * sizeof(unsigned int), c.stream);
	color_count_kernel<<<grid, 256, 0, c.stream>>>(m->N_elems, d_color, d_cnt + 2);
	c.launches++;
	e = cudaMemcpyAsync(h_cnt, d_cnt, 66 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c.stream);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(c.stream);
	if (e != cudaSuccess) {
		set_error("colouring: %s", cudaGetErrorString(e));
		return done(NBGPU_ERR_CUDA);
	}
	uint32_t n_colors = 0;
	for (uint32_t k = 0; k < 64; k++)
		if (h_cnt[2 + k])
			n_colors = k + 1;
	m->color_ptr.assign(n_colors + 1, 0);
	for (uint32_t k = 0; k < n_colors; k++)
		m->color_ptr[k + 1] = m->color_ptr[k] + h_cnt[2 + k];
	unsigned int h_next[64] = {0};
	for (uint32_t k = 0; k < n_colors; k++)
		h_next[k] = m->color_ptr[k];
	e = nbgpu::dmalloc(&m->d_color_elems, std::max<size_t>(1, m->N_elems) * sizeof(uint32_t));
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(d_cnt + 2, h_next, 64 * sizeof(unsigned int), cudaMemcpyHostToDevice, c.stream);
	if (e == cudaSuccess) {
		color_scatter_kernel<<<grid, 256, 0, c.stream>>>(m->N_elems, d_color, d_cnt + 2, m->d_color_elems);
		c.launches++;
		e = cudaStreamSynchronize(c.stream);
	}
	if (e != cudaSuccess) {
		set_error("colouring: %s", cudaGetErrorString(e));
		return done(NBGPU_ERR_CUDA);
	}
	m->n_colors = n_colors;
	m->d_color_block = d_color_block;
	m->d_color = d_color;
	return done(NBGPU_OK);
}

template <int NPE, int NGP>
int launch_assembly(nbgpu_matrix_t *K, nbgpu_mesh_t *m, const AsmParams &P, int mode, const uint8_t *d_en,
		    const double *d_scale, const double *d_damage, double *d_F, unsigned int *d_bad, int *d_miss)
{
	Context &c = ctx();
	const bool node_form = mode == NBGPU_ASSEMBLY_GATHER && K->blocked && K->max_width <= 2 * kNodeMaxBlocks &&
			       (K->N & 1u) == 0;
	if (d_damage) {
		// the damage driver's loop exists in the node-parallel schedule only
		if (!node_form) {
			set_error("per-Gauss-point damage needs the GATHER schedule on a 2-dof blocked matrix");
			return NBGPU_ERR_ARG;
		}
		const uint32_t pairs = K->n_slices * 16u;
		const uint32_t nb_max = (K->max_width + 1) / 2;
		const size_t smem = (size_t)nb_max * kNodeThreads * (4 * sizeof(double) + sizeof(uint32_t));
		if (smem > 48 * 1024)
			NB_CUDA(cudaFuncSetAttribute(assemble_node_kernel<NPE, NGP, true>,
						     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		assemble_node_kernel<NPE, NGP, true><<<(pairs + kNodeThreads - 1) / kNodeThreads, kNodeThreads, smem, c.stream>>>(
			K->N, K->col_shift, nb_max, m->d_nod, m->d_adj, m->d_n2e_ptr, m->d_n2e, d_en, d_scale, d_damage, P,
			K->d_slice_off, K->d_perm, K->d_bcol, K->d_val, d_F, d_bad, d_miss);
		NB_LAUNCHED();
		return NBGPU_OK;
	}
	if (node_form && !getenv("NBGPU_ASSEMBLY_ROWS")) {
		// node-parallel form: overwrites every entry of the rows, no reset needed
		const uint32_t pairs = K->n_slices * 16u;
		const uint32_t nb_max = (K->max_width + 1) / 2;
		const size_t smem = (size_t)nb_max * kNodeThreads * (4 * sizeof(double) + sizeof(uint32_t));
		if (smem > 48 * 1024)
			NB_CUDA(cudaFuncSetAttribute(assemble_node_kernel<NPE, NGP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
						     (int)smem));
		assemble_node_kernel<NPE, NGP><<<(pairs + kNodeThreads - 1) / kNodeThreads, kNodeThreads, smem, c.stream>>>(
			K->N, K->col_shift, nb_max, m->d_nod, m->d_adj, m->d_n2e_ptr, m->d_n2e, d_en, d_scale, nullptr, P,
			K->d_slice_off, K->d_perm, K->d_bcol, K->d_val, d_F, d_bad, d_miss);
		NB_LAUNCHED();
	} else if (mode == NBGPU_ASSEMBLY_GATHER) {
		NB_CUDA(cudaMemsetAsync(K->d_val, 0, K->stored * sizeof(double), c.stream));   // nb_sparse_reset (pipeline.c:54)
		const uint32_t rows = K->n_slices * kSliceRows;
		assemble_gather_kernel<NPE, NGP><<<(rows + kBlock - 1) / kBlock, kBlock, 0, c.stream>>>(
			K->N, K->col_shift, m->d_nod, m->d_adj, m->d_n2e_ptr, m->d_n2e, d_en, d_scale, P, K->d_slice_off,
			K->d_perm, K->d_col, K->d_val, d_F, d_bad, d_miss);
		NB_LAUNCHED();
	} else if (mode == NBGPU_ASSEMBLY_ATOMIC) {
		NB_CUDA(cudaMemsetAsync(K->d_val, 0, K->stored * sizeof(double), c.stream));
		NB_CUDA(cudaMemsetAsync(d_F, 0, 2 * (size_t)m->N_nod * sizeof(double), c.stream));
		assemble_element_kernel<NPE, NGP, true><<<(m->N_elems + 127) / 128, 128, 0, c.stream>>>(
			m->N_elems, nullptr, m->d_nod, m->d_adj, d_en, d_scale, P, K->d_slice_off, K->d_inv_perm,
			K->d_col, K->d_val, d_F, d_bad, d_miss);
		NB_LAUNCHED();
	} else {
		NB_TRY(build_coloring(m));
		NB_CUDA(cudaMemsetAsync(K->d_val, 0, K->stored * sizeof(double), c.stream));
		NB_CUDA(cudaMemsetAsync(d_F, 0, 2 * (size_t)m->N_nod * sizeof(double), c.stream));
		for (uint32_t col = 0; col < m->n_colors; col++) {
			const uint32_t n = m->color_ptr[col + 1] - m->color_ptr[col];
			if (!n)
				continue;
			assemble_element_kernel<NPE, NGP, false><<<(n + 127) / 128, 128, 0, c.stream>>>(
				n, m->d_color_elems + m->color_ptr[col], m->d_nod, m->d_adj, d_en, d_scale, P,
				K->d_slice_off, K->d_inv_perm, K->d_col, K->d_val, d_F, d_bad, d_miss);
			NB_LAUNCHED();
		}
	}
	return NBGPU_OK;
}

}  // namespace

extern "C" {

int nbgpu_elem_tables_default(uint32_t nodes_per_elem, nbgpu_elem_tables_t *t)
{
	NB_ARG(t != nullptr && (nodes_per_elem == 3 || nodes_per_elem == 4));
	memset(t, 0, sizeof(*t));
	t->N_nodes = nodes_per_elem;
	if (nodes_per_elem == 3) {
		// element.c:54-73
		t->N_gp = 1;
		t->gp_weight[0] = 0.5;
		for (int i = 0; i < 3; i++)
			t->Ni[i] = 0.33333333333333333333333333333;
		t->dNi_dpsi[0] = -1.0; t->dNi_dpsi[1] = 1.0; t->dNi_dpsi[2] = 0.0;
		t->dNi_deta[0] = -1.0; t->dNi_deta[1] = 0.0; t->dNi_deta[2] = 1.0;
		return NBGPU_OK;
	}
	// element.c:75-121.  Bilinear shape functions at the 2x2 Gauss points
	// (+-1/sqrt3), corners and points both ordered (-,-),(+,-),(+,+),(-,+).
	// The reference tabulates them with 12 significant digits; those rounded
	// literals (not the exact values) are what parity requires.
	t->N_gp = 4;
	const double n_near = 0.622008467928, n_side = 0.166666666667, n_far = 0.044658198739;
	const double d_near = 0.394337567297, d_far = 0.105662432703;
	const int sx[4] = {-1, 1, 1, -1}, sy[4] = {-1, -1, 1, 1};
	for (int g = 0; g < 4; g++)
		t->gp_weight[g] = 1.0;
	for (int c = 0; c < 4; c++)
		for (int g = 0; g < 4; g++) {
			const bool same_x = sx[c] == sx[g], same_y = sy[c] == sy[g];
			t->Ni[c * 4 + g] = (same_x && same_y) ? n_near : (same_x || same_y) ? n_side : n_far;
			t->dNi_dpsi[c * 4 + g] = sx[c] * (same_y ? d_near : d_far);
			t->dNi_deta[c * 4 + g] = sy[c] * (same_x ? d_near : d_far);
		}
	return NBGPU_OK;
}

int nbgpu_constitutive_matrix(double E, double poisson, int analysis2D, double D[4])
{
	NB_ARG(D != nullptr);
	// formulas.c:38-45: the switch has no `break`, so NB_PLANE_STRESS,
	// NB_PLANE_STRAIN and NB_SOLID_OF_REVOLUTION all end in set_plane_stress
	// (formulas.c:48-54).  Kept on purpose: parity with the reference.
	(void)analysis2D;
	D[0] = E / (1.0 - poisson * poisson);
	D[1] = poisson * D[0];
	D[2] = D[0];
	D[3] = E / (2.0 * (1.0 + poisson));
	return NBGPU_OK;
}

int nbgpu_mesh_create(uint32_t N_nod, const double *nod, uint32_t N_elems, uint32_t nodes_per_elem,
		      const uint32_t *adj, nbgpu_mesh_t **out)
{
	NB_INIT();
	NB_ARG(out != nullptr && nod != nullptr && adj != nullptr);
	NB_ARG(nodes_per_elem == 3 || nodes_per_elem == 4);
	const uint32_t npe = nodes_per_elem;
	for (size_t k = 0; k < (size_t)npe * N_elems; k++)
		NB_ARG(adj[k] < N_nod);
	nbgpu_mesh_t *m = new nbgpu_mesh_t();
	m->N_nod = N_nod;
	m->N_elems = N_elems;
	m->npe = npe;
	// elements around each node, ascending element id (counting sort)
	std::vector<uint32_t> ptr((size_t)N_nod + 1, 0), n2e((size_t)npe * N_elems);
	for (size_t k = 0; k < (size_t)npe * N_elems; k++)
		ptr[adj[k] + 1]++;
	for (uint32_t i = 0; i < N_nod; i++)
		ptr[i + 1] += ptr[i];
	{
		std::vector<uint32_t> next(ptr.begin(), ptr.end() - 1);
		for (uint32_t e = 0; e < N_elems; e++)
			for (uint32_t i = 0; i < npe; i++) {
				const uint32_t v = adj[(size_t)e * npe + i];
				// a node listed twice in one element is visited once
				if (next[v] > ptr[v] && n2e[next[v] - 1] == e)
					continue;
				n2e[next[v]++] = e;
			}
		// compact (only needed if some element repeated a node)
		bool compact = false;
		for (uint32_t i = 0; i < N_nod && !compact; i++)
			compact = next[i] != ptr[i + 1];
		if (compact) {
			std::vector<uint32_t> ptr2((size_t)N_nod + 1, 0), n2e2;
			for (uint32_t i = 0; i < N_nod; i++) {
				n2e2.insert(n2e2.end(), n2e.begin() + ptr[i], n2e.begin() + next[i]);
				ptr2[i + 1] = (uint32_t)n2e2.size();
			}
			n2e2.resize((size_t)npe * N_elems);
			ptr.swap(ptr2);
			n2e.swap(n2e2);
		}
	}
	cudaError_t e = nbgpu::dmalloc(&m->d_nod, std::max<size_t>(1, 2 * (size_t)N_nod) * sizeof(double));
	if (e == cudaSuccess)
		e = nbgpu::dmalloc(&m->d_adj, std::max<size_t>(1, (size_t)npe * N_elems) * sizeof(uint32_t));
	if (e == cudaSuccess)
		e = nbgpu::dmalloc(&m->d_n2e_ptr, ptr.size() * sizeof(uint32_t));
	if (e == cudaSuccess)
		e = nbgpu::dmalloc(&m->d_n2e, std::max<size_t>(1, n2e.size()) * sizeof(uint32_t));
	if (e == cudaSuccess)
		e = nbgpu::dmalloc(&m->d_enabled, std::max<size_t>(1, N_elems));
	if (e == cudaSuccess)
		e = nbgpu::dmalloc(&m->d_scale, std::max<size_t>(1, N_elems) * sizeof(double));
	if (e == cudaSuccess)
		e = cudaMemcpy(m->d_nod, nod, 2 * (size_t)N_nod * sizeof(double), cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
		e = cudaMemcpy(m->d_adj, adj, (size_t)npe * N_elems * sizeof(uint32_t), cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
		e = cudaMemcpy(m->d_n2e_ptr, ptr.data(), ptr.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
		e = cudaMemcpy(m->d_n2e, n2e.data(), n2e.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
	if (e != cudaSuccess) {
		set_error("mesh upload: %s", cudaGetErrorString(e));
		cudaGetLastError();
		nbgpu_mesh_destroy(m);
		return NBGPU_ERR_CUDA;
	}
	*out = m;
	return NBGPU_OK;
}

int nbgpu_mesh_destroy(nbgpu_mesh_t *m)
{
	if (!m)
		return NBGPU_OK;
	if (ctx().ready) {
		cudaSetDevice(ctx().device);
		cudaStreamSynchronize(ctx().stream);
		nbgpu::dfree(m->d_nod);
		nbgpu::dfree(m->d_adj);
		nbgpu::dfree(m->d_n2e_ptr);
		nbgpu::dfree(m->d_n2e);
		nbgpu::dfree(m->d_enabled);
		nbgpu::dfree(m->d_scale);
		nbgpu::dfree(m->d_color_elems);
		nbgpu::dfree(m->d_color_block);
	}
	delete m;
	return NBGPU_OK;
}

int nbgpu_mesh_coloring(nbgpu_mesh_t *m, uint32_t *n_colors, uint8_t *colors)
{
	NB_INIT();
	NB_ARG(m != nullptr);
	NB_TRY(build_coloring(m));
	if (n_colors)
		*n_colors = m->n_colors;
	if (colors && m->N_elems)
		NB_CUDA(cudaMemcpy(colors, m->d_color, m->N_elems, cudaMemcpyDeviceToHost));
	return NBGPU_OK;
}

static int assemble_elasticity2d(nbgpu_matrix_t *K, const nbgpu_mesh_t *mesh_c, const nbgpu_elem_tables_t *tables,
				 const nbgpu_assembly_params_t *params, const uint8_t *enabled,
				 const double *elem_scale, const double *gp_damage, double *d_F, uint32_t *first_bad);

int nbgpu_assemble_elasticity2d(nbgpu_matrix_t *K, const nbgpu_mesh_t *mesh_c,
				const nbgpu_elem_tables_t *tables, const nbgpu_assembly_params_t *params,
				const uint8_t *enabled, const double *elem_scale, double *d_F,
				uint32_t *first_bad)
{
	return assemble_elasticity2d(K, mesh_c, tables, params, enabled, elem_scale, nullptr, d_F, first_bad);
}

int nbgpu_assemble_elasticity2d_damage(nbgpu_matrix_t *K, const nbgpu_mesh_t *mesh_c,
				       const nbgpu_elem_tables_t *tables, const nbgpu_assembly_params_t *params,
				       const uint8_t *enabled, const double *gp_damage, double *d_F,
				       uint32_t *first_bad)
{
	if (!gp_damage) {
		set_error("gp_damage is NULL");
		return NBGPU_ERR_ARG;
	}
	return assemble_elasticity2d(K, mesh_c, tables, params, enabled, nullptr, gp_damage, d_F, first_bad);
}

static int assemble_elasticity2d(nbgpu_matrix_t *K, const nbgpu_mesh_t *mesh_c, const nbgpu_elem_tables_t *tables,
				 const nbgpu_assembly_params_t *params, const uint8_t *enabled,
				 const double *elem_scale, const double *gp_damage, double *d_F, uint32_t *first_bad)
{
	NB_INIT();
	nbgpu_mesh_t *m = const_cast<nbgpu_mesh_t *>(mesh_c);
	NB_ARG(K != nullptr && m != nullptr && params != nullptr && d_F != nullptr);
	NB_ARG(K->local_block ? K->n_cols == 2 * m->N_nod : K->N == 2 * m->N_nod);
	NB_ARG(params->mode >= NBGPU_ASSEMBLY_GATHER && params->mode <= NBGPU_ASSEMBLY_COLOR);
	/* rank-local blocks: the row-parallel schedule only (element-parallel scatter would hit ghost rows) */
	NB_ARG(!K->local_block || params->mode == NBGPU_ASSEMBLY_GATHER);
	Context &c = ctx();
	NB_TRY(upload_tables(tables, m->npe));
	AsmParams P;
	memcpy(P.D, params->D, sizeof(P.D));
	memcpy(P.D_void, params->D_void, sizeof(P.D_void));
	P.density = params->density;
	P.density_void = params->density_void;
	P.thickness = params->thickness;
	P.self_weight = params->self_weight != 0;
	P.gx = P.self_weight ? params->gravity[0] : 0.0;
	P.gy = P.self_weight ? params->gravity[1] : 0.0;
	const uint8_t *d_en = nullptr;
	const double *d_scale = nullptr;
	if (enabled) {
		NB_CUDA(cudaMemcpyAsync(m->d_enabled, enabled, m->N_elems, cudaMemcpyHostToDevice, c.stream));
		d_en = m->d_enabled;
	}
	if (elem_scale) {
		NB_CUDA(cudaMemcpyAsync(m->d_scale, elem_scale, (size_t)m->N_elems * sizeof(double),
					cudaMemcpyHostToDevice, c.stream));
		d_scale = m->d_scale;
	}
	double *d_damage = nullptr;
	if (gp_damage) {
		const size_t n_gp = (size_t)m->N_elems * (m->npe == 3 ? 1 : 4);
		NB_CUDA(nbgpu::dmalloc(&d_damage, n_gp * sizeof(double)));
		cudaError_t ce = cudaMemcpyAsync(d_damage, gp_damage, n_gp * sizeof(double), cudaMemcpyHostToDevice, c.stream);
		if (ce != cudaSuccess) {
			nbgpu::dfree(d_damage);
			NB_CUDA(ce);
		}
	}
	// flags: [0] lowest distorted element id, [1] pattern miss
	unsigned int *d_flags = nullptr;
	{
		cudaError_t ce = nbgpu::dmalloc(&d_flags, 2 * sizeof(unsigned int));
		if (ce != cudaSuccess) {
			nbgpu::dfree(d_damage);
			NB_CUDA(ce);
		}
	}
	const unsigned int init_flags[2] = {0xFFFFFFFFu, 0u};
	NB_CUDA(cudaMemcpyAsync(d_flags, init_flags, sizeof(init_flags), cudaMemcpyHostToDevice, c.stream));
	// nb_sparse_reset (pipeline.c:54) and the zeroing of F (pipeline.c:57) are done by the schedules themselves
	int st;
	if (m->npe == 3)
		st = launch_assembly<3, 1>(K, m, P, params->mode, d_en, d_scale, d_damage, d_F, d_flags, (int *)(d_flags + 1));
	else
		st = launch_assembly<4, 4>(K, m, P, params->mode, d_en, d_scale, d_damage, d_F, d_flags, (int *)(d_flags + 1));
	unsigned int h_flags[2] = {0xFFFFFFFFu, 0u};
	cudaError_t e = cudaMemcpyAsync(h_flags, d_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, c.stream);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(c.stream);
	nbgpu::dfree(d_flags);
	nbgpu::dfree(d_damage);
	if (st != NBGPU_OK)
		return st;
	if (e != cudaSuccess) {
		set_error("assembly: %s", cudaGetErrorString(e));
		return NBGPU_ERR_CUDA;
	}
	if (h_flags[1]) {
		set_error("assembly: an element entry is not in the sparsity pattern "
			  "(the reference exits here, sparse.c:213-217)");
		return NBGPU_ERR_PATTERN;
	}
	if (first_bad)
		*first_bad = h_flags[0];
	return h_flags[0] != 0xFFFFFFFFu ? NBGPU_DISTORTED_ELEMENT : NBGPU_OK;
}

int nbgpu_assemble_lumped_mass(const nbgpu_mesh_t *mesh_c, const nbgpu_elem_tables_t *tables, double density,
			       double density_void, double thickness, const uint8_t *enabled, double *d_M,
			       uint32_t *first_bad)
{
	NB_INIT();
	nbgpu_mesh_t *m = const_cast<nbgpu_mesh_t *>(mesh_c);
	NB_ARG(m != nullptr && d_M != nullptr);
	Context &c = ctx();
	NB_TRY(upload_tables(tables, m->npe));
	const uint8_t *d_en = nullptr;
	if (enabled) {
		NB_CUDA(cudaMemcpyAsync(m->d_enabled, enabled, m->N_elems, cudaMemcpyHostToDevice, c.stream));
		d_en = m->d_enabled;
	}
	unsigned int *d_flag = nullptr;
	NB_CUDA(nbgpu::dmalloc(&d_flag, sizeof(unsigned int)));
	const unsigned int init_flag = 0xFFFFFFFFu;
	NB_CUDA(cudaMemcpyAsync(d_flag, &init_flag, sizeof(init_flag), cudaMemcpyHostToDevice, c.stream));
	const int grid = (int)((m->N_nod + kBlock - 1) / kBlock);
	if (m->npe == 3)
		lumped_mass_kernel<3, 1><<<std::max(grid, 1), kBlock, 0, c.stream>>>(
			m->N_nod, m->d_nod, m->d_adj, m->d_n2e_ptr, m->d_n2e, d_en, density, density_void, thickness, d_M,
			d_flag);
	else
		lumped_mass_kernel<4, 4><<<std::max(grid, 1), kBlock, 0, c.stream>>>(
			m->N_nod, m->d_nod, m->d_adj, m->d_n2e_ptr, m->d_n2e, d_en, density, density_void, thickness, d_M,
			d_flag);
	NB_LAUNCHED();
	unsigned int h_flag = 0xFFFFFFFFu;
	cudaError_t e = cudaMemcpyAsync(&h_flag, d_flag, sizeof(h_flag), cudaMemcpyDeviceToHost, c.stream);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(c.stream);
	nbgpu::dfree(d_flag);
	if (e != cudaSuccess) {
		set_error("lumped mass: %s", cudaGetErrorString(e));
		return NBGPU_ERR_CUDA;
	}
	if (first_bad)
		*first_bad = h_flag;
	return h_flag != 0xFFFFFFFFu ? NBGPU_DISTORTED_ELEMENT : NBGPU_OK;
}

int nbgpu_vector_add_entries(double *d_F, uint32_t n, const uint32_t *dof, const double *add)
{
	NB_INIT();
	if (n == 0)
		return NBGPU_OK;
	NB_ARG(d_F != nullptr && dof != nullptr && add != nullptr);
	Context &c = ctx();
	void *buf = nullptr;
	NB_CUDA(nbgpu::dmalloc(&buf, (size_t)n * (sizeof(uint32_t) + sizeof(double))));
	double *d_add = (double *)buf;
	uint32_t *d_dof = (uint32_t *)(d_add + n);
	cudaError_t e = cudaMemcpyAsync(d_add, add, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c.stream);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(d_dof, dof, (size_t)n * sizeof(uint32_t), cudaMemcpyHostToDevice, c.stream);
	if (e == cudaSuccess) {
		vector_add_entries_kernel<<<(n + 255) / 256, 256, 0, c.stream>>>(d_F, n, d_dof, d_add);
		ctx().launches++;
		e = cudaGetLastError();
	}
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(c.stream);
	nbgpu::dfree(buf);
	if (e != cudaSuccess) {
		set_error("vector_add_entries: %s", cudaGetErrorString(e));
		return NBGPU_ERR_CUDA;
	}
	return NBGPU_OK;
}

/* prepared (device-resident) list of prescribed dofs: built once, applied after every re-assembly */
struct nbgpu_dirichlet_s {
	uint32_t N = 0;
	size_t m = 0;
	void *buf = nullptr;          // v_first[m] | v_last[m] | order[N]
	const double *d_first = nullptr, *d_last = nullptr;
	const uint32_t *d_order = nullptr;
};

int nbgpu_dirichlet_destroy(nbgpu_dirichlet_t *bc)
{
	if (!bc)
		return NBGPU_OK;
	if (ctx().ready && bc->buf) {
		cudaStreamSynchronize(ctx().stream);
		nbgpu::dfree(bc->buf);
	}
	delete bc;
	return NBGPU_OK;
}

int nbgpu_dirichlet_create(uint32_t N, uint32_t n, const uint32_t *dof, const double *value,
			   nbgpu_dirichlet_t **out)
{
	NB_INIT();
	NB_ARG(out != nullptr && (n == 0 || (dof != nullptr && value != nullptr)));
	// flatten the ordered list: first occurrence decides the elimination order
	// and the value moved to the right-hand side, the last occurrence decides
	// F[dof] (see dirichlet_kernel)
	std::vector<uint32_t> order(N, 0);
	std::vector<double> v_first, v_last;
	v_first.reserve(n);
	v_last.reserve(n);
	for (uint32_t k = 0; k < n; k++) {
		NB_ARG(dof[k] < N);
		if (!order[dof[k]]) {
			v_first.push_back(value[k]);
			v_last.push_back(value[k]);
			order[dof[k]] = (uint32_t)v_first.size();
		} else {
			v_last[order[dof[k]] - 1] = value[k];
		}
	}
	nbgpu_dirichlet_t *bc = new nbgpu_dirichlet_t();
	bc->N = N;
	bc->m = v_first.size();
	const size_t m = bc->m;
	cudaError_t e = nbgpu::dmalloc(&bc->buf, (size_t)N * sizeof(uint32_t) + 2 * m * sizeof(double) + 16);
	double *d_first = (double *)bc->buf, *d_last = d_first + m;
	uint32_t *d_order = (uint32_t *)(d_last + m);
	if (e == cudaSuccess)
		e = cudaMemcpy(d_first, v_first.data(), m * sizeof(double), cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
		e = cudaMemcpy(d_last, v_last.data(), m * sizeof(double), cudaMemcpyHostToDevice);
	if (e == cudaSuccess)
		e = cudaMemcpy(d_order, order.data(), (size_t)N * sizeof(uint32_t), cudaMemcpyHostToDevice);
	if (e != cudaSuccess) {
		set_error("dirichlet_create: %s", cudaGetErrorString(e));
		cudaGetLastError();
		bc->buf = e == cudaErrorMemoryAllocation ? nullptr : bc->buf;
		nbgpu_dirichlet_destroy(bc);
		return NBGPU_ERR_CUDA;
	}
	bc->d_first = d_first;
	bc->d_last = d_last;
	bc->d_order = d_order;
	*out = bc;
	return NBGPU_OK;
}

/* stream-ordered; does not synchronise */
int nbgpu_dirichlet_apply(nbgpu_matrix_t *K, double *d_F, const nbgpu_dirichlet_t *bc)
{
	NB_INIT();
	NB_ARG(K != nullptr && d_F != nullptr && bc != nullptr && bc->N == K->n_cols);
	if (bc->m == 0)
		return NBGPU_OK;
	const uint32_t n_pos = K->n_slices * kSliceRows;
	dirichlet_kernel<<<(n_pos + kBlock - 1) / kBlock, kBlock, 0, ctx().stream>>>(
		K->N, K->col_shift, n_pos, K->d_slice_off, K->d_perm, K->d_col, K->d_val, d_F, bc->d_order, bc->d_first, bc->d_last);
	NB_LAUNCHED();
	return NBGPU_OK;
}

int nbgpu_apply_dirichlet(nbgpu_matrix_t *K, double *d_F, uint32_t n, const uint32_t *dof,
			  const double *value)
{
	NB_INIT();
	if (n == 0)
		return NBGPU_OK;
	NB_ARG(K != nullptr && d_F != nullptr && dof != nullptr && value != nullptr);
	nbgpu_dirichlet_t *bc = nullptr;
	NB_TRY(nbgpu_dirichlet_create(K->n_cols, n, dof, value, &bc));
	int st = nbgpu_dirichlet_apply(K, d_F, bc);
	if (st == NBGPU_OK && cudaStreamSynchronize(ctx().stream) != cudaSuccess) {
		set_error("apply_dirichlet: %s", cudaGetErrorString(cudaGetLastError()));
		st = NBGPU_ERR_CUDA;
	}
	nbgpu_dirichlet_destroy(bc);
	return st;
}

/* F[d_dof[k]] += d_add[k] with the lists already on the device; stream-ordered */
int nbgpu_vector_add_entries_dev(double *d_F, uint32_t n, const uint32_t *d_dof, const double *d_add)
{
	NB_INIT();
	if (n == 0)
		return NBGPU_OK;
	NB_ARG(d_F != nullptr && d_dof != nullptr && d_add != nullptr);
	vector_add_entries_kernel<<<(n + 255) / 256, 256, 0, ctx().stream>>>(d_F, n, d_dof, d_add);
	NB_LAUNCHED();
	return NBGPU_OK;
}

int nbgpu_compute_strain(const nbgpu_mesh_t *m, const nbgpu_elem_tables_t *tables, const double *d_disp,
			 double *d_strain)
{
	NB_INIT();
	NB_ARG(m != nullptr && d_disp != nullptr && d_strain != nullptr);
	NB_TRY(upload_tables(tables, m->npe));
	if (m->N_elems == 0)
		return NBGPU_OK;
	const int grid = (m->N_elems + kBlock - 1) / kBlock;
	if (m->npe == 3)
		strain_kernel<3, 1><<<grid, kBlock, 0, ctx().stream>>>(m->N_elems, m->d_nod, m->d_adj, d_disp,
								       d_strain);
	else
		strain_kernel<4, 4><<<grid, kBlock, 0, ctx().stream>>>(m->N_elems, m->d_nod, m->d_adj, d_disp,
								       d_strain);
	NB_LAUNCHED();
	return NBGPU_OK;
}

int nbgpu_gp_to_nodes(const nbgpu_mesh_t *m, const nbgpu_elem_tables_t *tables, uint32_t N_comp,
		      const double *d_gp_values, double *d_nodal_values)
{
	NB_INIT();
	NB_ARG(m != nullptr && d_gp_values != nullptr && d_nodal_values != nullptr && N_comp > 0);
	NB_TRY(upload_tables(tables, m->npe));
	if (m->N_nod == 0)
		return NBGPU_OK;
	Context &c = ctx();
	unsigned int *d_bad = nullptr;
	NB_CUDA(nbgpu::dmalloc(&d_bad, sizeof(unsigned int)));
	cudaMemsetAsync(d_bad, 0xFF, sizeof(unsigned int), c.stream);
	const size_t n = (size_t)m->N_nod * N_comp;
	const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
	if (m->npe == 3)
		gp_to_nodes_kernel<3, 1><<<grid, kBlock, 0, c.stream>>>(m->N_nod, N_comp, m->d_nod, m->d_adj, m->d_n2e_ptr,
								       m->d_n2e, d_gp_values, d_nodal_values, d_bad);
	else
		gp_to_nodes_kernel<4, 4><<<grid, kBlock, 0, c.stream>>>(m->N_nod, N_comp, m->d_nod, m->d_adj, m->d_n2e_ptr,
								       m->d_n2e, d_gp_values, d_nodal_values, d_bad);
	c.launches++;
	unsigned int bad = 0xFFFFFFFFu;
	cudaError_t e = cudaGetLastError();
	if (e == cudaSuccess)
		e = cudaMemcpyAsync(&bad, d_bad, sizeof(bad), cudaMemcpyDeviceToHost, c.stream);
	if (e == cudaSuccess)
		e = cudaStreamSynchronize(c.stream);
	nbgpu::dfree(d_bad);
	if (e != cudaSuccess) {
		set_error("gp_to_nodes: %s", cudaGetErrorString(e));
		return NBGPU_ERR_CUDA;
	}
	return bad == 0xFFFFFFFFu ? NBGPU_OK : NBGPU_DISTORTED_ELEMENT;
}

int nbgpu_von_mises(uint64_t n_points, const double *d_stress, double *d_vm)
{
	NB_INIT();
	NB_ARG(d_stress != nullptr && d_vm != nullptr);
	if (n_points == 0)
		return NBGPU_OK;
	von_mises_kernel<<<(unsigned)((n_points + kBlock - 1) / kBlock), kBlock, 0, ctx().stream>>>(n_points, d_stress, d_vm);
	NB_LAUNCHED();
	return NBGPU_OK;
}

int nbgpu_main_stress(uint64_t n_points, const double *d_stress, double *d_main)
{
	NB_INIT();
	NB_ARG(d_stress != nullptr && d_main != nullptr);
	if (n_points == 0)
		return NBGPU_OK;
	main_stress_kernel<<<(unsigned)((n_points + kBlock - 1) / kBlock), kBlock, 0, ctx().stream>>>(n_points, d_stress, d_main);
	NB_LAUNCHED();
	return NBGPU_OK;
}

int nbgpu_stress_from_strain(uint32_t N_elems, uint32_t N_gp, const double D[4], const double D_void[4],
			     const uint8_t *enabled, const double *d_strain, double *d_stress)
{
	NB_INIT();
	NB_ARG(D != nullptr && D_void != nullptr && d_strain != nullptr && d_stress != nullptr);
	NB_ARG(N_gp == 1 || N_gp == 4);
	if (N_elems == 0)
		return NBGPU_OK;
	Context &c = ctx();
	uint8_t *d_en = nullptr;
	if (enabled) {
		NB_CUDA(nbgpu::dmalloc(&d_en, N_elems));
		NB_CUDA(cudaMemcpyAsync(d_en, enabled, N_elems, cudaMemcpyHostToDevice, c.stream));
	}
	const size_t n = (size_t)N_elems * N_gp;
	stress_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, c.stream>>>(
		N_elems, N_gp, d_en, D[0], D[1], D[2], D[3], D_void[0], D_void[1], D_void[2], D_void[3],
		d_strain, d_stress);
	ctx().launches++;
	cudaError_t e = cudaGetLastError();
	if (e == cudaSuccess && d_en)
		e = cudaStreamSynchronize(c.stream);
	if (d_en)
		nbgpu::dfree(d_en);
	if (e != cudaSuccess) {
		set_error("stress_from_strain: %s", cudaGetErrorString(e));
		return NBGPU_ERR_CUDA;
	}
	return NBGPU_OK;
}

}  // extern "C"
