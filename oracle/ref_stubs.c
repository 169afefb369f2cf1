/*
 * oracle/ref_stubs.c -- TEST INFRASTRUCTURE.
 * The reference's vendored METIS/GKlib (sources/nb/graph_bot/imported_libs,
 * ~40 kLoC) is only reached from the direct solvers' nested-dissection
 * relabelling (sources/nb/graph_bot/labeling/nested_dissection.c:22,39), never
 * from the Jacobi-PCG / assembly path.  oracle/Makefile leaves it out and
 * satisfies the two symbols with traps so libnbots_ref.so can be dlopen'ed.
 */
#include <stdio.h>
#include <stdlib.h>

int METIS_SetDefaultOptions(void *options)
{
	(void)options;
	return 1;
}

int METIS_NodeND(void *a, void *b, void *c, void *d, void *e, void *f, void *g)
{
	(void)a; (void)b; (void)c; (void)d; (void)e; (void)f; (void)g;
	fprintf(stderr, "oracle/_ref: METIS is not part of the oracle build\n");
	abort();
}
