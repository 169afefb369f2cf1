// pattern.cu -- host-side construction of the stiffness sparsity pattern,
// bit-exact with the reference's two steps:
//   nb_mesh2D_load_graph(mesh, graph, NB_NODES_LINKED_BY_ELEMS)
//       (sources/nb/geometric_bot/mesh/mesh2D/load_graph.c:230-328)
//   nb_sparse_create(graph, NULL, vars_per_node)
//       (sources/nb/solver_bot/sparse/sparse.c:20-60)
// The reference makes two heap blocks per row and qsorts each row; here the
// node lists are built once per node in flat arrays (OpenMP over nodes) and
// every dof row of a node is written from the same sorted list.
//
// The C ABI is the usual two-call form (sizes first, then the caller-allocated
// column array).  The sorted node lists of the sizing call are kept for the
// fill call that follows it (one entry, validated by sizes, addresses and a
// checksum of the mesh arrays), so the pair costs one construction, not two.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <mutex>

#include "common.cuh"

using namespace nbgpu;

namespace {

struct NodeLists {
	std::vector<uint64_t> ptr;     // [N_nod + 1] start of each node's neighbour list in nbr
	std::vector<uint32_t> nbr;     // sorted (and, without an edge list, deduplicated) neighbours
	std::vector<uint32_t> count;   // [N_nod] valid entries of each list
};

struct ListKey {
	uint32_t N_nod = 0, N_elems = 0, npe = 0, N_edg = 0;
	const uint32_t *adj = nullptr, *edg = nullptr;
	uint64_t checksum = 0;
	bool operator==(const ListKey &o) const
	{
		return N_nod == o.N_nod && N_elems == o.N_elems && npe == o.npe && N_edg == o.N_edg && adj == o.adj &&
		       edg == o.edg && checksum == o.checksum;
	}
};

std::mutex g_cache_lock;
bool g_cache_valid = false;
ListKey g_cache_key;
NodeLists g_cache_lists;

// order-dependent 64-bit checksum of an array (parallel over blocks, combined in block order)
uint64_t checksum_u32(const uint32_t *a, size_t n, uint64_t seed)
{
	if (!a || !n)
		return seed;
	const size_t block = size_t(1) << 16, n_blocks = (n + block - 1) / block;
	std::vector<uint64_t> part(n_blocks);
#pragma omp parallel for schedule(static)
	for (int64_t b = 0; b < (int64_t)n_blocks; b++) {
		uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)b;
		const size_t lo = (size_t)b * block, hi = std::min(n, lo + block);
		for (size_t i = lo; i < hi; i++)
			h = (h ^ a[i]) * 0x100000001B3ull;
		part[b] = h;
	}
	uint64_t h = seed;
	for (uint64_t p : part)
		h = (h ^ p) * 0x100000001B3ull + 0x632BE59BD9B4E019ull;
	return h;
}

// small lists (a node has 4-9 neighbours): insertion sort beats std::sort's dispatch
inline void sort_small(uint32_t *b, uint32_t *e)
{
	if (e - b > 24) {
		std::sort(b, e);
		return;
	}
	for (uint32_t *i = b + 1; i < e; i++) {
		const uint32_t v = *i;
		uint32_t *j = i;
		for (; j > b && j[-1] > v; j--)
			*j = j[-1];
		*j = v;
	}
}

int build_node_lists(uint32_t N_nod, uint32_t N_elems, uint32_t npe, const uint32_t *adj, uint32_t N_edg,
		     const uint32_t *edg, NodeLists &L)
{
	const bool from_edges = edg != nullptr;
	// neighbour-slot count per node (an upper bound when deduplicating)
	L.ptr.assign((size_t)N_nod + 1, 0);
	std::vector<uint64_t> &ptr = L.ptr;
	if (from_edges) {
		// load_graph.c:281-303: both ends of every edge ...
		for (uint32_t e = 0; e < N_edg; e++) {
			NB_ARG(edg[2 * e] < N_nod && edg[2 * e + 1] < N_nod);
			ptr[edg[2 * e] + 1]++;
			ptr[edg[2 * e + 1] + 1]++;
		}
		// ... plus, per element, the npe-3 nodes that are not edge neighbours (:309-328)
		for (size_t k = 0; k < (size_t)npe * N_elems; k++) {
			NB_ARG(adj[k] < N_nod);
			ptr[adj[k] + 1] += npe - 3;
		}
	} else {
		for (size_t k = 0; k < (size_t)npe * N_elems; k++) {
			NB_ARG(adj[k] < N_nod);
			ptr[adj[k] + 1] += npe - 1;
		}
	}
	for (uint32_t i = 0; i < N_nod; i++)
		ptr[i + 1] += ptr[i];
	L.nbr.resize(ptr[N_nod]);
	std::vector<uint32_t> &nbr = L.nbr;
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
	// sort each node's list; without an edge list the same neighbour is seen once per shared
	// element and is deduplicated
	L.count.resize(N_nod);
#pragma omp parallel for schedule(static)
	for (int64_t i = 0; i < (int64_t)N_nod; i++) {
		uint32_t *b = nbr.data() + ptr[i], *e = nbr.data() + ptr[i + 1];
		sort_small(b, e);
		if (!from_edges)
			e = std::unique(b, e);
		L.count[i] = (uint32_t)(e - b);
	}
	return NBGPU_OK;
}

}  // namespace

extern "C" int nbgpu_pattern_from_mesh(uint32_t N_nod, uint32_t N_elems, uint32_t nodes_per_elem,
				       const uint32_t *adj, uint32_t N_edg, const uint32_t *edg,
				       uint32_t vars_per_node, uint32_t *rows_size, uint32_t *cols,
				       uint64_t *nnz_out)
{
	NB_ARG(adj != nullptr || N_elems == 0);
	NB_ARG(rows_size != nullptr && vars_per_node >= 1);
	NB_ARG(nodes_per_elem >= 3 && nodes_per_elem <= 4);
	const uint32_t npe = nodes_per_elem, vars = vars_per_node;
	const bool trace = getenv("NBGPU_TRACE") != nullptr;
	auto t_last = std::chrono::steady_clock::now();
	auto lap = [&](const char *what) {
		if (!trace)
			return;
		auto t = std::chrono::steady_clock::now();
		fprintf(stderr, "[nbgpu pattern] %-10s %8.3f ms\n", what,
			std::chrono::duration<double, std::milli>(t - t_last).count());
		t_last = t;
	};

	ListKey key;
	key.N_nod = N_nod; key.N_elems = N_elems; key.npe = npe; key.N_edg = edg ? N_edg : 0;
	key.adj = adj; key.edg = edg;
	key.checksum = checksum_u32(edg, edg ? 2 * (size_t)N_edg : 0, checksum_u32(adj, (size_t)npe * N_elems, 1));
	NodeLists L;
	bool reused = false;
	{
		std::lock_guard<std::mutex> guard(g_cache_lock);
		if (cols && g_cache_valid && g_cache_key == key) {
			L = std::move(g_cache_lists);
			reused = true;
		}
		g_cache_valid = false;   // consumed, or superseded by what this call builds
	}
	lap(reused ? "reuse" : "checksum");
	if (!reused)
		NB_TRY(build_node_lists(N_nod, N_elems, npe, adj, edg ? N_edg : 0, edg, L));
	lap("lists");
	const std::vector<uint64_t> &ptr = L.ptr;
	const std::vector<uint32_t> &nbr = L.nbr, &count = L.count;

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
	lap("sizes");
	if (!cols) {
		// sizing call: the fill call that normally follows takes the lists from here
		std::lock_guard<std::mutex> guard(g_cache_lock);
		g_cache_lists = std::move(L);
		g_cache_key = key;
		g_cache_valid = true;
		return NBGPU_OK;
	}
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
	lap("rows");
	return NBGPU_OK;
}
