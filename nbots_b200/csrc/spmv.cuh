// spmv.cuh -- the SELL-32 row kernel shared by SpMV and the Krylov solvers.
//
// One thread owns one matrix row; a warp owns one slice.  For entry j the
// warp's 32 value loads are one 256-byte request and its 32 column-id loads
// one 128-byte request (streamed with evict-first: the matrix is read once per
// SpMV and must not displace the vectors from L1/L2); x is gathered through
// the read-only path.  UNROLL entries are loaded before the first is used so
// that each thread keeps 2*UNROLL independent streaming loads plus UNROLL
// gathers in flight.
//
// Arithmetic: products and sums are rounded separately (__dmul_rn/__dadd_rn,
// no FMA contraction) and accumulated in ascending column order from 0, which
// is exactly what the reference's row loop computes
// (sources/nb/solver_bot/sparse/sparse.c:405-414 compiled for x86-64 without
// FMA), so y = A x is reproduced bit for bit.  Padding entries are skipped by
// selection, never added.
#pragma once

#include "matrix.cuh"

namespace nbgpu {

// Streaming (evict-first) loads of the matrix arrays and read-only gathers of
// the vector, as volatile asm: ptxas keeps volatile asm statements in source
// order, which pins the "issue every load of the batch, then gather, then
// accumulate" schedule (left alone it interleaves load and use to save
// registers and serialises the memory latency).
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p)
{
	uint32_t r;
	asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(r) : "l"(p));
	return r;
}
__device__ __forceinline__ double ld_stream_f64(const double *p)
{
	double r;
	asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(r) : "l"(p));
	return r;
}
__device__ __forceinline__ double ld_gather_f64(const double *p)
{
	double r;
	asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(r) : "l"(p));
	return r;
}

template <int UNROLL, bool WANT_DIAG>
__device__ __forceinline__ double sell_row_times(const double *__restrict__ val,
						  const uint32_t *__restrict__ col, uint32_t off,
						  uint32_t width, uint32_t lane, uint32_t row,
						  uint32_t safe_col, const double *__restrict__ x,
						  double *diag)
{
	const double *v = val + (size_t)off * kSliceRows + lane;
	const uint32_t *c = col + (size_t)off * kSliceRows + lane;
	double acc = 0.0;
	double d = 0.0;
	uint32_t j = 0;
	for (; j + UNROLL <= width; j += UNROLL) {
		uint32_t cj[UNROLL];
		double vj[UNROLL], xj[UNROLL];
#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
			cj[u] = ld_stream_u32(c + (size_t)(j + u) * kSliceRows);
			vj[u] = ld_stream_f64(v + (size_t)(j + u) * kSliceRows);
		}
#pragma unroll
		for (int u = 0; u < UNROLL; u++)
			xj[u] = ld_gather_f64(x + (cj[u] == kPadCol ? safe_col : cj[u]));
#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
			const double t = __dmul_rn(vj[u], xj[u]);
			acc = (cj[u] == kPadCol) ? acc : __dadd_rn(acc, t);
			if (WANT_DIAG && cj[u] == row)
				d = vj[u];
		}
	}
	for (; j < width; j++) {
		const uint32_t cj = ld_stream_u32(c + (size_t)j * kSliceRows);
		const double vj = ld_stream_f64(v + (size_t)j * kSliceRows);
		const double xj = ld_gather_f64(x + (cj == kPadCol ? safe_col : cj));
		const double t = __dmul_rn(vj, xj);
		acc = (cj == kPadCol) ? acc : __dadd_rn(acc, t);
		if (WANT_DIAG && cj == row)
			d = vj;
	}
	if (WANT_DIAG)
		*diag = d;
	return acc;
}

// persistent grid for the slice-parallel kernels
int spmv_grid(uint32_t n_slices);

}  // namespace nbgpu
