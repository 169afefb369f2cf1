"""Runs every FEM-side kernel once at Q1 size (for the ncu pass of scripts/profile_r02.sh): assembly in the three
schedules, Dirichlet, strain, stress, Gauss points -> nodes, von Mises / main stresses."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nbots_b200 import api, capi, meshgen
from ab_pcg import BCS
from util import flatten_bcs
L = capi.lib(); capi.check(L.nbgpu_init(0))
m = meshgen.structured_mesh(1000, 500, 2.0, 1.0, kind=1)
rs, cols = api.pattern_from_mesh(m)
K = api.Matrix.from_csr(rs, cols); mesh = api.Mesh(m); d_F = api.DeviceBuffer.zeros(K.N)
neu_dof, neu_add, dir_dof, dir_val = flatten_bcs(m, BCS)
for rep in range(3):
    for mode in (capi.ASSEMBLY_GATHER, capi.ASSEMBLY_ATOMIC, capi.ASSEMBLY_COLOR):
        mesh.assemble(K, d_F, 1.0, 0.3, thickness=1.0, mode=mode)
    os.environ["NBGPU_ASSEMBLY_ROWS"] = "1"
    mesh.assemble(K, d_F, 1.0, 0.3, thickness=1.0)            # the row-parallel GATHER kept for unblocked matrices
    del os.environ["NBGPU_ASSEMBLY_ROWS"]
    mesh.assemble(K, d_F, 1.0, 0.3, thickness=1.0)
    api.vector_add_entries(d_F, neu_dof, neu_add)
    K.apply_dirichlet(d_F, dir_dof, dir_val)
    d_x = api.DeviceBuffer.from_host(meshgen.uniform_rhs(K.N))
    n_gp = 4 * m.n_elems
    d_strain = api.DeviceBuffer.zeros(3 * n_gp); d_stress = api.DeviceBuffer.zeros(3 * n_gp)
    mesh.compute_strain(d_x, d_strain)
    api.stress_from_strain(m.n_elems, 4, api.constitutive_matrix(1.0, 0.3), d_strain, d_stress)
    d_nod = api.DeviceBuffer.zeros(3 * m.n_nod)
    mesh.gp_to_nodes(3, d_stress, d_nod)
    d_vm = api.DeviceBuffer.zeros(n_gp); d_main = api.DeviceBuffer.zeros(2 * n_gp)
    api.von_mises(n_gp, d_stress, d_vm); api.main_stress(n_gp, d_stress, d_main)
api.sync()
print("fem kernels ok")
