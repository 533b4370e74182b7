// Launchers of the Gauss-point operator kernels (gauss_ops.cu); the C-ABI wrappers live in adfem_cuda.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "device_fem.cuh"
#include "grid_index.cuh"
#include "tet_grid.cuh"

namespace adfem {

// dof -> (element, local dof) adjacency of the symbolic phase, device pointers
struct DofAdjacency {
  const long long* ptr;
  const int* elem;
  const uint8_t* loc;
};

// Every launcher returns 0 or records an error (adfem_last_error) and returns 1.
// basis: GpBasis (gauss_ops.cuh); weighted: multiply by the Gauss weights w_k.
int launch_gp_gather(const DevMesh& dm, int degree, int basis, bool weighted, const double* in, double* out, cudaStream_t st);
int launch_gp_scatter(const DevMesh& dm, int degree, const DofAdjacency& adj, int basis, bool weighted, const double* in, double* out,
                      cudaStream_t st);
int launch_laplace_term(const DevMesh& dm, int degree, const DofAdjacency& adj, const double* nu, const double* u, double* out, cudaStream_t st);
int launch_laplace_term_grad_nu(const DevMesh& dm, int degree, const double* u, const double* grad_out, double* grad_nu, cudaStream_t st);
int launch_plane_matrix(int mode, long long n, const double* E, const double* nu, double* H, cudaStream_t st);
int launch_plane_matrix_grad(int mode, long long n, const double* E, const double* nu, const double* grad_H, double* grad_E, double* grad_nu,
                             cudaStream_t st);

// structured triangulation Mesh(m, n, h), P1 (grid_gauss.cuh): index-free versions of the scatter-type kernels
// xy: nullptr on rectilinear grids; the coordinate array ([node][2]) for structured connectivity on mapped / jittered node positions
int launch_grid_gp_scatter(const DevMesh& dm, const GridTri& gt, int basis, bool weighted, const double* in, double* out, cudaStream_t st, const double* xy = nullptr);
int launch_grid_laplace_term(const DevMesh& dm, const GridTri& gt, const double* nu, const double* u, double* out, cudaStream_t st, const double* xy = nullptr);

// structured tetrahedral grid Mesh3(n, n, l, h), P1 (tet_gauss.cuh)
int launch_tet_gp_scatter(const DevMesh& dm, const GridTet& gt, int basis, bool weighted, const double* in, double* out, cudaStream_t st);
int launch_tet_laplace_term(const DevMesh& dm, const GridTet& gt, const double* nu, const double* u, double* out, cudaStream_t st);

// option "coef_presum" (P1 elasticity): hbar[ne*ns2] = sum_k w_k coef[(e*g+k)*ns2 + c]; grad[(e*g+k)*ns2 + c] = w_k gbar[e*ns2 + c]
int launch_presum_coef(const DevMesh& dm, int ns2, const double* coef, double* hbar, cudaStream_t st);
int launch_expand_grad(const DevMesh& dm, int ns2, const double* gbar, double* grad, cudaStream_t st);
// fused constitutive pre-step (P1 triangles): hbar[ne*9] from E[G], nu[G]; (grad_E, grad_nu)[G] from gbar[ne*9]
int launch_presum_plane(const DevMesh& dm, int mode, const double* E, const double* nu, double* hbar, cudaStream_t st);
int launch_expand_plane_grad(const DevMesh& dm, int mode, const double* E, const double* nu, const double* gbar, double* grad_E, double* grad_nu,
                             cudaStream_t st);

}  // namespace adfem
