// Kernels of the Gauss-point operators (bodies in gauss_ops.cuh): one thread per element (gather), one thread per dof row
// (scatter / Laplace term), one thread per point (constitutive pre-step).  fp64, sm_100a; no atomics, outputs are written, never
// accumulated, so no caller-side zero fill is needed (the reference requires zero-filled outputs for most of these ops).
#include "gauss_ops.h"

#include <string>

#include "gauss_ops.cuh"
#include "grid_gauss.cuh"
#include "tet_gauss.cuh"
#include "internal.h"

namespace adfem {
namespace {

constexpr int GP_THREADS = 128;
inline unsigned gp_blocks(long long n) { return (unsigned)((n + GP_THREADS - 1) / GP_THREADS); }

template <int DIM, int DEG, int B, bool W>
__global__ void __launch_bounds__(GP_THREADS) k_gp_gather(DevMesh m, const double* __restrict__ in, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < m.ne) gp_gather_body<DIM, DEG, B, W>(m, e, in, out);
}

template <int DIM, int DEG, int B, bool W>
__global__ void __launch_bounds__(GP_THREADS) k_gp_scatter(DevMesh m, DofAdjacency adj, int nrows, const double* __restrict__ in,
                                                            double* __restrict__ out) {
  constexpr int NC = GpShape<DIM, DEG, B>::NC;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  double acc[NC];
  gp_scatter_row<DIM, DEG, B, W>(m, adj.ptr, adj.elem, adj.loc, r, in, acc);
#pragma unroll
  for (int c = 0; c < NC; c++) out[r + (size_t)c * nrows] = acc[c];
}

template <int DIM, int DEG>
__global__ void __launch_bounds__(GP_THREADS) k_laplace_term(DevMesh m, DofAdjacency adj, const double* __restrict__ nu,
                                                              const double* __restrict__ u, double* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < m.ndof) out[r] = laplace_term_row<DIM, DEG>(m, adj.ptr, adj.elem, adj.loc, r, nu, u);
}

template <int DIM, int DEG>
__global__ void __launch_bounds__(GP_THREADS) k_laplace_term_grad_nu(DevMesh m, const double* __restrict__ u, const double* __restrict__ go,
                                                                      double* __restrict__ grad_nu) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < m.ne) laplace_term_grad_nu_body<DIM, DEG>(m, e, u, go, grad_nu);
}

__global__ void k_plane_matrix(int mode, long long n, const double* __restrict__ E, const double* __restrict__ nu, double* __restrict__ H) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) plane_matrix_body(mode, E[i], nu[i], H + 9 * i);
}
__global__ void k_plane_matrix_grad(int mode, long long n, const double* __restrict__ E, const double* __restrict__ nu,
                                    const double* __restrict__ gH, double* __restrict__ gE, double* __restrict__ gnu) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) plane_matrix_grad_body(mode, E[i], nu[i], gH + 9 * i, gE + i, gnu + i);
}

// structured triangulation: one thread per node, node id = i*(m+1) + j
template <int B, bool W>
__global__ void __launch_bounds__(GP_THREADS) k_grid_gp_scatter(DevMesh m, GridTri gt, const double* __restrict__ xy, const double* __restrict__ in, double* __restrict__ out) {
  constexpr int NC = GpShape<2, 1, B>::NC;
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x, nn = (long long)(gt.m + 1) * (gt.n + 1);
  if (r >= nn) return;
  double acc[NC];
  grid_scatter_node<B, W>(gt, m.heron, m.rule, m.g, (int)(r / (gt.m + 1)), (int)(r % (gt.m + 1)), in, acc, xy);
#pragma unroll
  for (int c = 0; c < NC; c++) out[r + c * nn] = acc[c];
}
__global__ void __launch_bounds__(GP_THREADS) k_grid_laplace_term(DevMesh m, GridTri gt, const double* __restrict__ xy, const double* __restrict__ nu, const double* __restrict__ u,
                                                                   double* __restrict__ out) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x, nn = (long long)(gt.m + 1) * (gt.n + 1);
  if (r < nn) out[r] = grid_laplace_term_node(gt, m.heron, m.rule, m.g, (int)(r / (gt.m + 1)), (int)(r % (gt.m + 1)), nu, u, xy);
}

// structured tetrahedral grid: one thread per node, node id = (k*(n+1) + j)*(n+1) + i
template <int B, bool W>
__global__ void __launch_bounds__(GP_THREADS) k_tet_gp_scatter(DevMesh m, GridTet gt, const double* __restrict__ in, double* __restrict__ out) {
  constexpr int NC = GpShape<3, 1, B>::NC;
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x, n1 = gt.n + 1, nn = n1 * n1 * (gt.l + 1);
  if (r >= nn) return;
  double acc[NC];
  tg_scatter_node<B, W>(gt, m.rule, m.g, (int)(r % n1), (int)((r / n1) % n1), (int)(r / (n1 * n1)), in, acc);
#pragma unroll
  for (int c = 0; c < NC; c++) out[r + c * nn] = acc[c];
}
__global__ void __launch_bounds__(GP_THREADS) k_tet_laplace_term(DevMesh m, GridTet gt, const double* __restrict__ nu, const double* __restrict__ u,
                                                                  double* __restrict__ out) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x, n1 = gt.n + 1, nn = n1 * n1 * (gt.l + 1);
  if (r < nn) out[r] = tg_laplace_term_node(gt, m.rule, m.g, (int)(r % n1), (int)((r / n1) % n1), (int)(r / (n1 * n1)), nu, u);
}

__global__ void k_presum_coef(DevMesh m, int ns2, long long n, const double* __restrict__ coef, double* __restrict__ hbar) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) hbar[i] = presum_coef_body(m.rule, m.g, ns2, i, coef);
}
__global__ void k_expand_grad(DevMesh m, int ns2, long long n, const double* __restrict__ gbar, double* __restrict__ grad) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) grad[i] = expand_grad_body(m.rule, m.g, ns2, i, gbar);
}

__global__ void k_presum_plane(DevMesh m, int mode, long long n, const double* __restrict__ E, const double* __restrict__ nu, double* __restrict__ hbar) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) hbar[i] = presum_plane_body(m.rule, m.g, mode, i, E, nu);
}
__global__ void k_expand_plane_grad(DevMesh m, int mode, long long n, const double* __restrict__ E, const double* __restrict__ nu,
                                    const double* __restrict__ gbar, double* __restrict__ gE, double* __restrict__ gnu) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < n) expand_plane_grad_body(m.rule, m.g, mode, t, E, nu, gbar, gE, gnu);
}

int launched(const char* what) {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : fail(std::string(what) + ": " + cudaGetErrorString(e));
}

#define GP_DISPATCH_ELEM(dm, degree, CALL)                 \
  do {                                                      \
    if ((dm).dim == 2 && (degree) == 1) { CALL(2, 1); }     \
    else if ((dm).dim == 2) { CALL(2, 2); }                 \
    else if ((degree) == 1) { CALL(3, 1); }                 \
    else { CALL(3, 2); }                                    \
  } while (0)

}  // namespace

int launch_gp_gather(const DevMesh& dm, int degree, int basis, bool weighted, const double* in, double* out, cudaStream_t st) {
  if (dm.ne == 0) return 0;
#define CALL_G(DIM, DEG)                                                                                                   \
  switch (basis) {                                                                                                         \
    case GB_P1SHAPE: k_gp_gather<DIM, DEG, GB_P1SHAPE, false><<<gp_blocks(dm.ne), GP_THREADS, 0, st>>>(dm, in, out); break; \
    case GB_SHAPE: k_gp_gather<DIM, DEG, GB_SHAPE, false><<<gp_blocks(dm.ne), GP_THREADS, 0, st>>>(dm, in, out); break;     \
    case GB_GRAD: k_gp_gather<DIM, DEG, GB_GRAD, false><<<gp_blocks(dm.ne), GP_THREADS, 0, st>>>(dm, in, out); break;       \
    default:                                                                                                               \
      if (weighted) k_gp_gather<DIM, DEG, GB_STRAIN, true><<<gp_blocks(dm.ne), GP_THREADS, 0, st>>>(dm, in, out);           \
      else k_gp_gather<DIM, DEG, GB_STRAIN, false><<<gp_blocks(dm.ne), GP_THREADS, 0, st>>>(dm, in, out);                   \
  }
  if (basis < GB_P1SHAPE || basis > GB_STRAIN || (weighted && basis != GB_STRAIN)) return fail("gauss-point gather: unknown operator");
  GP_DISPATCH_ELEM(dm, degree, CALL_G);
#undef CALL_G
  return launched("gauss-point gather kernel");
}

int launch_gp_scatter(const DevMesh& dm, int degree, const DofAdjacency& adj, int basis, bool weighted, const double* in, double* out,
                      cudaStream_t st) {
  // P1SHAPE acts on the vertex rows only (FemToGaussPointsMfem reads u[node]); the vertex dofs are rows 0..nv-1 of the adjacency
  const int nrows = basis == GB_P1SHAPE ? dm.nv : dm.ndof;
  if (nrows == 0) return 0;
#define CALL_S(DIM, DEG)                                                                                                                 \
  switch (basis) {                                                                                                                       \
    case GB_P1SHAPE: k_gp_scatter<DIM, DEG, GB_P1SHAPE, false><<<gp_blocks(nrows), GP_THREADS, 0, st>>>(dm, adj, nrows, in, out); break;  \
    case GB_SHAPE: k_gp_scatter<DIM, DEG, GB_SHAPE, false><<<gp_blocks(nrows), GP_THREADS, 0, st>>>(dm, adj, nrows, in, out); break;      \
    case GB_GRAD: k_gp_scatter<DIM, DEG, GB_GRAD, false><<<gp_blocks(nrows), GP_THREADS, 0, st>>>(dm, adj, nrows, in, out); break;        \
    default:                                                                                                                             \
      if (weighted) k_gp_scatter<DIM, DEG, GB_STRAIN, true><<<gp_blocks(nrows), GP_THREADS, 0, st>>>(dm, adj, nrows, in, out);            \
      else k_gp_scatter<DIM, DEG, GB_STRAIN, false><<<gp_blocks(nrows), GP_THREADS, 0, st>>>(dm, adj, nrows, in, out);                    \
  }
  if (basis < GB_P1SHAPE || basis > GB_STRAIN || (weighted && basis != GB_STRAIN)) return fail("gauss-point scatter: unknown operator");
  GP_DISPATCH_ELEM(dm, degree, CALL_S);
#undef CALL_S
  return launched("gauss-point scatter kernel");
}

int launch_laplace_term(const DevMesh& dm, int degree, const DofAdjacency& adj, const double* nu, const double* u, double* out, cudaStream_t st) {
  if (dm.ndof == 0) return 0;
#define CALL_L(DIM, DEG) k_laplace_term<DIM, DEG><<<gp_blocks(dm.ndof), GP_THREADS, 0, st>>>(dm, adj, nu, u, out)
  GP_DISPATCH_ELEM(dm, degree, CALL_L);
#undef CALL_L
  return launched("Laplace term kernel");
}

int launch_laplace_term_grad_nu(const DevMesh& dm, int degree, const double* u, const double* grad_out, double* grad_nu, cudaStream_t st) {
  if (dm.ne == 0) return 0;
#define CALL_LG(DIM, DEG) k_laplace_term_grad_nu<DIM, DEG><<<gp_blocks(dm.ne), GP_THREADS, 0, st>>>(dm, u, grad_out, grad_nu)
  GP_DISPATCH_ELEM(dm, degree, CALL_LG);
#undef CALL_LG
  return launched("Laplace term coefficient-gradient kernel");
}

int launch_plane_matrix(int mode, long long n, const double* E, const double* nu, double* H, cudaStream_t st) {
  if (mode != 0 && mode != 1) return fail("plane matrix: mode must be 0 (PlaneStrainMatrix) or 1 (PlaneStressMatrix)");
  if (n <= 0) return 0;
  k_plane_matrix<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mode, n, E, nu, H);
  return launched("plane matrix kernel");
}
int launch_plane_matrix_grad(int mode, long long n, const double* E, const double* nu, const double* grad_H, double* grad_E, double* grad_nu,
                             cudaStream_t st) {
  if (mode != 0 && mode != 1) return fail("plane matrix: mode must be 0 (PlaneStrainMatrix) or 1 (PlaneStressMatrix)");
  if (n <= 0) return 0;
  k_plane_matrix_grad<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mode, n, E, nu, grad_H, grad_E, grad_nu);
  return launched("plane matrix gradient kernel");
}

int launch_grid_gp_scatter(const DevMesh& dm, const GridTri& gt, int basis, bool weighted, const double* in, double* out, cudaStream_t st, const double* xy) {
  const long long nn = (long long)(gt.m + 1) * (gt.n + 1);
  const unsigned nb = gp_blocks(nn);
  if (basis < GB_P1SHAPE || basis > GB_STRAIN || (weighted && basis != GB_STRAIN)) return fail("gauss-point scatter: unknown operator");
  switch (basis) {
    case GB_P1SHAPE: k_grid_gp_scatter<GB_P1SHAPE, false><<<nb, GP_THREADS, 0, st>>>(dm, gt, xy, in, out); break;
    case GB_SHAPE: k_grid_gp_scatter<GB_SHAPE, false><<<nb, GP_THREADS, 0, st>>>(dm, gt, xy, in, out); break;
    case GB_GRAD: k_grid_gp_scatter<GB_GRAD, false><<<nb, GP_THREADS, 0, st>>>(dm, gt, xy, in, out); break;
    default:
      if (weighted) k_grid_gp_scatter<GB_STRAIN, true><<<nb, GP_THREADS, 0, st>>>(dm, gt, xy, in, out);
      else k_grid_gp_scatter<GB_STRAIN, false><<<nb, GP_THREADS, 0, st>>>(dm, gt, xy, in, out);
  }
  return launched("structured gauss-point scatter kernel");
}
int launch_grid_laplace_term(const DevMesh& dm, const GridTri& gt, const double* nu, const double* u, double* out, cudaStream_t st, const double* xy) {
  k_grid_laplace_term<<<gp_blocks((long long)(gt.m + 1) * (gt.n + 1)), GP_THREADS, 0, st>>>(dm, gt, xy, nu, u, out);
  return launched("structured Laplace term kernel");
}

int launch_tet_gp_scatter(const DevMesh& dm, const GridTet& gt, int basis, bool weighted, const double* in, double* out, cudaStream_t st) {
  const long long n1 = gt.n + 1, nn = n1 * n1 * (gt.l + 1);
  const unsigned nb = gp_blocks(nn);
  if (basis < GB_P1SHAPE || basis > GB_STRAIN || (weighted && basis != GB_STRAIN)) return fail("gauss-point scatter: unknown operator");
  switch (basis) {
    case GB_P1SHAPE: k_tet_gp_scatter<GB_P1SHAPE, false><<<nb, GP_THREADS, 0, st>>>(dm, gt, in, out); break;
    case GB_SHAPE: k_tet_gp_scatter<GB_SHAPE, false><<<nb, GP_THREADS, 0, st>>>(dm, gt, in, out); break;
    case GB_GRAD: k_tet_gp_scatter<GB_GRAD, false><<<nb, GP_THREADS, 0, st>>>(dm, gt, in, out); break;
    default:
      if (weighted) k_tet_gp_scatter<GB_STRAIN, true><<<nb, GP_THREADS, 0, st>>>(dm, gt, in, out);
      else k_tet_gp_scatter<GB_STRAIN, false><<<nb, GP_THREADS, 0, st>>>(dm, gt, in, out);
  }
  return launched("structured tetrahedral gauss-point scatter kernel");
}
int launch_tet_laplace_term(const DevMesh& dm, const GridTet& gt, const double* nu, const double* u, double* out, cudaStream_t st) {
  const long long n1 = gt.n + 1;
  k_tet_laplace_term<<<gp_blocks(n1 * n1 * (gt.l + 1)), GP_THREADS, 0, st>>>(dm, gt, nu, u, out);
  return launched("structured tetrahedral Laplace term kernel");
}

int launch_presum_coef(const DevMesh& dm, int ns2, const double* coef, double* hbar, cudaStream_t st) {
  const long long n = (long long)dm.ne * ns2;
  if (n <= 0) return 0;
  k_presum_coef<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dm, ns2, n, coef, hbar);
  return launched("coefficient pre-sum kernel");
}
int launch_expand_grad(const DevMesh& dm, int ns2, const double* gbar, double* grad, cudaStream_t st) {
  const long long n = (long long)dm.ne * dm.g * ns2;
  if (n <= 0) return 0;
  k_expand_grad<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dm, ns2, n, gbar, grad);
  return launched("gradient expansion kernel");
}

int launch_presum_plane(const DevMesh& dm, int mode, const double* E, const double* nu, double* hbar, cudaStream_t st) {
  if (mode != 0 && mode != 1) return fail("plane matrix: mode must be 0 (PlaneStrainMatrix) or 1 (PlaneStressMatrix)");
  const long long n = (long long)dm.ne * 9;
  if (n <= 0) return 0;
  k_presum_plane<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dm, mode, n, E, nu, hbar);
  return launched("fused plane-matrix pre-sum kernel");
}
int launch_expand_plane_grad(const DevMesh& dm, int mode, const double* E, const double* nu, const double* gbar, double* grad_E, double* grad_nu,
                             cudaStream_t st) {
  if (mode != 0 && mode != 1) return fail("plane matrix: mode must be 0 (PlaneStrainMatrix) or 1 (PlaneStressMatrix)");
  const long long n = (long long)dm.ne * dm.g;
  if (n <= 0) return 0;
  k_expand_plane_grad<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dm, mode, n, E, nu, gbar, grad_E, grad_nu);
  return launched("fused plane-matrix gradient kernel");
}

}  // namespace adfem
