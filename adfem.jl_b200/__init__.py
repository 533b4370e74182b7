"""adfem.jl_b200 — B200-native differentiable FEM assembly: host-side mirror of the AdFem.jl operator API
over libadfem_cuda.so (include/adfem_cuda.h).  Import as `adfem_jl_b200` (see the shim at the repo root)."""
from . import _lib, meshgen  # noqa: F401
from ._lib import AdfemError, build  # noqa: F401
from .mesh import (P1, P2, Mesh, Mesh3, bcedge, bcnode, fem_nodes, gauss_nodes, gauss_nodes_soa, gauss_weights, get_area,  # noqa: F401
                   get_edge_dof, get_ngauss, get_volume, read_mesh_file)
from .ops import (CSRTensor, SparseTensor, compute_fem_laplace_matrix1, compute_fem_mass_matrix1,  # noqa: F401,E402
                  compute_fem_source_term, compute_fem_source_term1, compute_fem_stiffness_matrix, compute_fem_stiffness_matrix1,
                  compute_space_varying_tangent_elasticity_matrix, coo_indices, impose_Dirichlet_boundary_conditions, fem_impose_Dirichlet_boundary_condition_experimental,
                  pcl_compute_fem_laplace_matrix1, pcl_impose_Dirichlet_boundary_conditions,
                  fem_to_gauss_points, dof_to_gauss_points, eval_grad_on_gauss_pts1, eval_strain_on_gauss_pts,
                  compute_strain_energy_term, compute_fem_laplace_term1, compute_plane_strain_matrix, compute_plane_stress_matrix,
                  compute_fem_stiffness_matrix_from_moduli, compute_fem_stiffness_matrix1_from_mu)
