// mesh.cuh -- device-resident FEM mesh (nbgpu_mesh_create, assembly.cu)
#pragma once

#include <vector>

#include "common.cuh"

struct nbgpu_mesh_s {
	uint32_t N_nod = 0, N_elems = 0, npe = 0;
	double *d_nod = nullptr;          // [2 N_nod]
	uint32_t *d_adj = nullptr;        // [npe N_elems]
	uint32_t *d_n2e_ptr = nullptr;    // [N_nod + 1] elements around each node ...
	uint32_t *d_n2e = nullptr;        // [npe N_elems] ... ascending element id
	uint8_t *d_enabled = nullptr;     // [N_elems] scratch for the enabled mask
	double *d_scale = nullptr;        // [N_elems] scratch for per-element factors
	uint32_t n_colors = 0;
	std::vector<uint32_t> color_ptr;  // [n_colors + 1]
	uint32_t *d_color_elems = nullptr;    // elements grouped by colour
	uint8_t *d_color = nullptr;           // [N_elems] colour of every element (inside d_color_block)
	uint8_t *d_color_block = nullptr;
};

