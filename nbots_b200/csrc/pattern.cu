// pattern.cu -- host-side construction of the stiffness sparsity pattern,
// bit-exact with the reference's two steps:
//   nb_mesh2D_load_graph(mesh, graph, NB_NODES_LINKED_BY_ELEMS)
//       (sources/nb/geometric_bot/mesh/mesh2D/load_graph.c:230-328)
//   nb_sparse_create(graph, NULL, vars_per_node)
//       (sources/nb/solver_bot/sparse/sparse.c:20-60)
// The reference makes two heap blocks per row and qsorts each row; here the
// node lists are built once per node in flat arrays (OpenMP over nodes) and
// every dof row of a node is written from the same sorted list.
#include <algorithm>
#include <cstring>

#include "common.cuh"

using namespace nbgpu;

extern "C" int nbgpu_pattern_from_mesh(uint32_t N_nod, uint32_t N_elems, uint32_t nodes_per_elem,
				       const uint32_t *adj, uint32_t N_edg, const uint32_t *edg,
				       uint32_t vars_per_node, uint32_t *rows_size, uint32_t *cols,
				       uint64_t *nnz_out)
{
	NB_ARG(adj != nullptr || N_elems == 0);
	NB_ARG(rows_size != nullptr && vars_per_node >= 1);
	NB_ARG(nodes_per_elem >= 3 && nodes_per_elem <= 4);
	const uint32_t npe = nodes_per_elem, vars = vars_per_node;
	const bool from_edges = edg != nullptr;

	// neighbour-slot count per node (an upper bound when deduplicating)
	std::vector<uint64_t> ptr((size_t)N_nod + 1, 0);
	if (from_edges) {
		// load_graph.c:281-303: both ends of every edge ...
		for (uint32_t e = 0; e < N_edg; e++) {
			NB_ARG(edg[2 * e] < N_nod && edg[2 * e + 1] < N_nod);
			ptr[edg[2 * e] + 1]++;
			ptr[edg[2 * e + 1] + 1]++;
		}
		// ... plus, per element, the npe-3 nodes that are not edge neighbours (:309-328)
		if (npe > 3)
			for (size_t k = 0; k < (size_t)npe * N_elems; k++)
				ptr[adj[k] + 1] += npe - 3;
	} else {
		for (size_t k = 0; k < (size_t)npe * N_elems; k++) {
			NB_ARG(adj[k] < N_nod);
			ptr[adj[k] + 1] += npe - 1;
		}
	}
	for (uint32_t i = 0; i < N_nod; i++)
		ptr[i + 1] += ptr[i];
	std::vector<uint32_t> nbr(ptr[N_nod]);
	{
		std::vector<uint64_t> next(ptr.begin(), ptr.end() - 1);
		if (from_edges) {
			for (uint32_t e = 0; e < N_edg; e++) {
				const uint32_t a = edg[2 * e], b = edg[2 * e + 1];
				nbr[next[a]++] = b;
				nbr[next[b]++] = a;
			}
			if (npe > 3)
				for (uint32_t el = 0; el < N_elems; el++) {
					const uint32_t *v = adj + (size_t)npe * el;
					for (uint32_t i = 0; i < npe; i++)
						for (uint32_t j = 0; j + 3 < npe; j++)
							nbr[next[v[i]]++] = v[(i + j + 2) % npe];
				}
		} else {
			for (uint32_t el = 0; el < N_elems; el++) {
				const uint32_t *v = adj + (size_t)npe * el;
				for (uint32_t i = 0; i < npe; i++)
					for (uint32_t j = 0; j < npe; j++)
						if (j != i)
							nbr[next[v[i]]++] = v[j];
			}
		}
	}
	// sort each node's list, insert the node itself; without an edge list the
	// same neighbour is seen once per shared element and is deduplicated
	std::vector<uint32_t> count(N_nod);
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t)N_nod; i++) {
		uint32_t *b = nbr.data() + ptr[i], *e = nbr.data() + ptr[i + 1];
		std::sort(b, e);
		if (!from_edges)
			e = std::unique(b, e);
		count[i] = (uint32_t)(e - b);
	}
	uint64_t nnz = 0;
	std::vector<uint64_t> out_ptr((size_t)N_nod + 1);
	for (uint32_t i = 0; i < N_nod; i++) {
		out_ptr[i] = nnz;
		const uint32_t len = (count[i] + 1) * vars;
		for (uint32_t k = 0; k < vars; k++)
			rows_size[(size_t)i * vars + k] = len;
		nnz += (uint64_t)len * vars;
	}
	out_ptr[N_nod] = nnz;
	if (nnz_out)
		*nnz_out = nnz;
	if (!cols)
		return NBGPU_OK;
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t)N_nod; i++) {
		const uint32_t *b = nbr.data() + ptr[i];
		const uint32_t n = count[i], len = (n + 1) * vars;
		uint32_t *row0 = cols + out_ptr[i];
		// merge the node itself into the sorted neighbour list
		uint32_t w = 0;
		bool self_done = false;
		for (uint32_t k = 0; k <= n; k++) {
			uint32_t node;
			if (!self_done && (k == n || b[k - (self_done ? 1 : 0)] > (uint32_t)i)) {
				node = (uint32_t)i;
				self_done = true;
			} else {
				node = b[k - (self_done ? 1 : 0)];
			}
			for (uint32_t k2 = 0; k2 < vars; k2++)
				row0[w++] = node * vars + k2;
		}
		for (uint32_t k1 = 1; k1 < vars; k1++)
			memcpy(row0 + (size_t)k1 * len, row0, (size_t)len * sizeof(uint32_t));
	}
	return NBGPU_OK;
}
