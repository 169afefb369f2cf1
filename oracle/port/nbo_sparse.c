/*
 * oracle/port/nbo_sparse.c -- TEST INFRASTRUCTURE (see nbo.h).
 * Node graph, sparsity pattern, SpMV, Dirichlet elimination; restated from
 * the reference over flat CSR arrays.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "nbo.h"

/* Reference: sources/nb/geometric_bot/mesh/mesh2D/load_graph.c:230-328.
 * Adjacency of node a = for every mesh edge (a,b) the other end, in edge
 * order (:296-303); then, for elements with more than 3 nodes, the nodes of
 * the element that are not edge-neighbours, k = (i + j + 2) % npe for
 * j < npe-3 (:309-328) -- for a quad: the opposite corner. */
uint64_t nbo_graph_nodes_by_elems(uint32_t N_nod, uint32_t N_edg,
				  const uint32_t *edg, uint32_t N_elems,
				  uint32_t npe, const uint32_t *adj,
				  uint32_t *N_adj, uint32_t *adj_flat)
{
	uint64_t *start = NULL;
	if (adj_flat) {
		/* N_adj holds the counts of the first pass */
		start = malloc(((size_t)N_nod + 1) * sizeof(*start));
		start[0] = 0;
		for (uint32_t i = 0; i < N_nod; i++)
			start[i + 1] = start[i] + N_adj[i];
	}
	memset(N_adj, 0, (size_t)N_nod * sizeof(*N_adj));
	for (uint32_t e = 0; e < N_edg; e++) {
		uint32_t a = edg[2 * e], b = edg[2 * e + 1];
		if (adj_flat) {
			adj_flat[start[a] + N_adj[a]] = b;
			adj_flat[start[b] + N_adj[b]] = a;
		}
		N_adj[a]++;
		N_adj[b]++;
	}
	if (npe > 3) {
		for (uint32_t el = 0; el < N_elems; el++) {
			const uint32_t *v = adj + (size_t)npe * el;
			for (uint32_t i = 0; i < npe; i++) {
				for (uint32_t j = 0; j + 3 < npe; j++) {
					uint32_t k = (i + j + 2) % npe;
					if (adj_flat)
						adj_flat[start[v[i]] +
							 N_adj[v[i]]] = v[k];
					N_adj[v[i]]++;
				}
			}
		}
	}
	uint64_t tot = 0;
	for (uint32_t i = 0; i < N_nod; i++)
		tot += N_adj[i];
	free(start);
	return tot;
}

static int cmp_u32(const void *a, const void *b)
{
	uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
	return (x > y) - (x < y);
}

/* Reference: sources/nb/solver_bot/sparse/sparse.c:20-60 (perm == NULL).
 * Node i with neighbours adj(i) gives `vars` rows i*vars+k1, each holding the
 * columns {j*vars+k2 : j in adj(i)} and {i*vars+k2}, sorted ascending
 * (nb_qsort with nb_compare_uint32, :55). */
uint64_t nbo_sparse_pattern(uint32_t N_graph, const uint32_t *N_adj,
			    const uint32_t *adj_flat, uint32_t vars,
			    uint32_t *rows_size, uint32_t *cols)
{
	uint64_t nnz = 0, goff = 0;
	for (uint32_t i = 0; i < N_graph; i++) {
		uint32_t len = (N_adj[i] + 1) * vars;
		for (uint32_t k1 = 0; k1 < vars; k1++) {
			rows_size[i * vars + k1] = len;
			if (cols) {
				uint32_t *row = cols + nnz;
				for (uint32_t j = 0; j < N_adj[i]; j++)
					for (uint32_t k2 = 0; k2 < vars; k2++)
						row[j * vars + k2] =
							adj_flat[goff + j] * vars + k2;
				for (uint32_t k2 = 0; k2 < vars; k2++)
					row[N_adj[i] * vars + k2] = i * vars + k2;
				qsort(row, len, sizeof(uint32_t), cmp_u32);
			}
			nnz += len;
		}
		goff += N_adj[i];
	}
	return nnz;
}

void nbo_row_ptr(uint32_t N, const uint32_t *rows_size, uint64_t *row_ptr)
{
	row_ptr[0] = 0;
	for (uint32_t i = 0; i < N; i++)
		row_ptr[i + 1] = row_ptr[i] + rows_size[i];
}

/* Reference: sparse.c:405-414.  out[i] accumulates in ascending column order
 * starting from 0. */
void nbo_spmv(uint32_t N, const uint64_t *row_ptr, const uint32_t *cols,
	      const double *vals, const double *in, double *out,
	      uint32_t threads)
{
#pragma omp parallel for num_threads(threads) schedule(guided)
	for (uint32_t i = 0; i < N; i++) {
		double acc = 0;
		for (uint64_t k = row_ptr[i]; k < row_ptr[i + 1]; k++)
			acc += vals[k] * in[cols[k]];
		out[i] = acc;
	}
}

/* position of column `col` in row i, or UINT64_MAX (binary search like
 * sparse_struct.c:8-21; the columns are ascending and unique) */
static uint64_t find_entry(const uint64_t *row_ptr, const uint32_t *cols,
			   uint32_t i, uint32_t col)
{
	int64_t lo = (int64_t)row_ptr[i], hi = (int64_t)row_ptr[i + 1] - 1;
	while (lo <= hi) {
		int64_t mid = (lo + hi) / 2;
		if (cols[mid] == col)
			return (uint64_t)mid;
		if (cols[mid] < col)
			lo = mid + 1;
		else
			hi = mid - 1;
	}
	return UINT64_MAX;
}

uint64_t nbo_find_entry(const uint64_t *row_ptr, const uint32_t *cols,
			uint32_t i, uint32_t col)
{
	return find_entry(row_ptr, cols, i, col);
}

/* Reference: sparse.c:416-430.  Row idx becomes the identity row and
 * rhs[idx] = value; every other stored (idx,j) is zeroed together with its
 * mirror (j,idx), whose old value moves to the right-hand side. */
void nbo_dirichlet(const uint64_t *row_ptr, const uint32_t *cols,
		   double *vals, double *rhs, uint32_t idx, double value)
{
	for (uint64_t k = row_ptr[idx]; k < row_ptr[idx + 1]; k++) {
		uint32_t j = cols[k];
		if (j == idx) {
			vals[k] = 1.0;
			rhs[idx] = value;
		} else {
			vals[k] = 0.0;
			uint64_t m = find_entry(row_ptr, cols, j, idx);
			if (m == UINT64_MAX) {
				/* sparse.c:189-205: the reference exits here */
				fprintf(stderr, "nbo_dirichlet: (%u,%u) not "
					"in the pattern\n", j, idx);
				exit(1);
			}
			double a = vals[m];
			vals[m] = 0.0;
			rhs[j] -= a * value;
		}
	}
}
