/*
 * oracle/port/nbo_cg.c -- TEST INFRASTRUCTURE (see nbo.h).
 * Jacobi-preconditioned CG and plain CG restated from the reference.
 */
#include <stdlib.h>
#include <math.h>
#include "nbo.h"

uint64_t nbo_find_entry(const uint64_t *row_ptr, const uint32_t *cols,
			uint32_t i, uint32_t col);

/* Reference: sources/nb/solver_bot/sparse/solvers/cg_precond_jacobi.c:13-90.
 *   init (:33-43)   g = A x - b, q = g / diag(A), p = -q, gg = g.g
 *   loop (:45)      while gg > tol^2 and k < max_iter
 *     pass 1 (:50-58)  w = A p ; pw = p.w ; gg = g.g ; gq = g.q
 *                      alpha = gq / pw
 *     pass 2 (:62-68)  x += alpha p ; g += alpha w ; q = g / diag ; gq' = g.q
 *                      beta = gq' / gq
 *     pass 3 (:72-74)  p = -q + beta p
 *   The loop test therefore sees the residual of the iterate BEFORE the last
 *   update, and tol_reached (:84) reports that same stale value. */
int nbo_pcg_jacobi(uint32_t N, const uint64_t *row_ptr, const uint32_t *cols,
		   const double *vals, const double *b, double *x,
		   uint32_t max_iter, double tol, uint32_t *iters,
		   double *tol_reached, uint32_t threads)
{
	double *g = calloc((size_t)5 * (N ? N : 1), sizeof(double));
	double *p = g + N, *q = p + N, *w = q + N, *d = w + N;
	double gg = 0;
#pragma omp parallel for reduction(+:gg) num_threads(threads) schedule(guided)
	for (uint32_t i = 0; i < N; i++) {
		double acc = 0;
		for (uint64_t k = row_ptr[i]; k < row_ptr[i + 1]; k++)
			acc += vals[k] * x[cols[k]];
		g[i] = acc - b[i];
		uint64_t m = nbo_find_entry(row_ptr, cols, i, i);
		d[i] = (m == UINT64_MAX) ? 0.0 : vals[m];  /* nb_sparse_get */
		q[i] = g[i] / d[i];
		p[i] = -q[i];
		gg += g[i] * g[i];
	}
	uint32_t k = 0;
	while (gg > tol * tol && k < max_iter) {
		double pw = 0, gq = 0;
		gg = 0;
#pragma omp parallel for reduction(+:pw, gg, gq) num_threads(threads)
		for (uint32_t i = 0; i < N; i++) {
			double acc = 0;
			for (uint64_t e = row_ptr[i]; e < row_ptr[i + 1]; e++)
				acc += vals[e] * p[cols[e]];
			w[i] = acc;
			pw += p[i] * w[i];
			gg += g[i] * g[i];
			gq += g[i] * q[i];
		}
		double alpha = gq / pw;
		double gq_new = 0;
#pragma omp parallel for reduction(+:gq_new) num_threads(threads)
		for (uint32_t i = 0; i < N; i++) {
			x[i] += alpha * p[i];
			g[i] += alpha * w[i];
			q[i] = g[i] / d[i];
			gq_new += g[i] * q[i];
		}
		double beta = gq_new / gq;
#pragma omp parallel for num_threads(threads)
		for (uint32_t i = 0; i < N; i++)
			p[i] = -q[i] + beta * p[i];
		k++;
	}
	free(g);
	if (iters)
		*iters = k;
	if (tol_reached)
		*tol_reached = sqrt(gg);
	return gg > tol * tol ? 1 : 0;
}

/* Reference: sources/nb/solver_bot/sparse/solvers/conjugate_gradient.c:13-77
 * -- the same recurrence with q == g (alpha = gg/pw, beta = g'g'/gg). */
int nbo_cg(uint32_t N, const uint64_t *row_ptr, const uint32_t *cols,
	   const double *vals, const double *b, double *x,
	   uint32_t max_iter, double tol, uint32_t *iters,
	   double *tol_reached, uint32_t threads)
{
	double *g = calloc((size_t)3 * (N ? N : 1), sizeof(double));
	double *p = g + N, *w = p + N;
	double gg = 0;
#pragma omp parallel for reduction(+:gg) num_threads(threads) schedule(guided)
	for (uint32_t i = 0; i < N; i++) {
		double acc = 0;
		for (uint64_t k = row_ptr[i]; k < row_ptr[i + 1]; k++)
			acc += vals[k] * x[cols[k]];
		g[i] = acc - b[i];
		p[i] = -g[i];
		gg += g[i] * g[i];
	}
	uint32_t k = 0;
	while (gg > tol * tol && k < max_iter) {
		double pw = 0;
		gg = 0;
#pragma omp parallel for reduction(+:pw, gg) num_threads(threads) schedule(guided)
		for (uint32_t i = 0; i < N; i++) {
			double acc = 0;
			for (uint64_t e = row_ptr[i]; e < row_ptr[i + 1]; e++)
				acc += vals[e] * p[cols[e]];
			w[i] = acc;
			pw += p[i] * w[i];
			gg += g[i] * g[i];
		}
		double alpha = gg / pw;
		double gg_new = 0;
#pragma omp parallel for reduction(+:gg_new) num_threads(threads) schedule(guided)
		for (uint32_t i = 0; i < N; i++) {
			x[i] += alpha * p[i];
			g[i] += alpha * w[i];
			gg_new += g[i] * g[i];
		}
		double beta = gg_new / gg;
#pragma omp parallel for num_threads(threads)
		for (uint32_t i = 0; i < N; i++)
			p[i] = -g[i] + beta * p[i];
		k++;
	}
	free(g);
	if (iters)
		*iters = k;
	if (tol_reached)
		*tol_reached = sqrt(gg);
	return gg > tol * tol ? 1 : 0;
}
