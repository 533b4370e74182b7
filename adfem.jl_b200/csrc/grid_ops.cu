// Structured-grid Q1 operators on an m x n grid of h x h cells:
//   UnivariateFemStiffness  deps/FemStiffness1/UnivariateFemStiffness.h:7-247   (compute_fem_stiffness_matrix1, src/InvCore.jl:67-76)
//   FemStiffness            deps/FemStiffness/FemStiffness.h:7-109              (constant 3x3 H)
//   SpatialFemStiffness     deps/SpatialFemStiffness/SpatialFemStiffness.h:7-133 (per-Gauss H)
//   SpatialVaryingTangentElastic  deps/SpatialVaryingTangentElastic/...h:1-70
// Index arithmetic only, no connectivity.  Output slot order, 1-based ii/jj and every quirk (cell loop i-outer /
// j-inner but cell id j*m+i, Q8; transposed Gauss pairing of SpatialFemStiffness, Q6; column-major constant H of
// FemStiffness, Q7) follow the reference.  The constant-coefficient adjoints reduce over all cells with a fixed-order
// two-pass reduction (no atomics: bit-reproducible).
#include <cuda_runtime.h>

#include <algorithm>
#include <string>

#include "../../include/adfem_cuda.h"
#include "internal.h"
#include "quad_ops.cuh"

using namespace adfem;

namespace {

#define CU_TRY(call)                                                                                   \
  do {                                                                                                 \
    cudaError_t _e = (call);                                                                           \
    if (_e != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(_e));            \
  } while (0)

__device__ __forceinline__ double gpt(int i) { return i == 0 ? (-1 / sqrt(3.0) + 1.0) / 2.0 : (1 / sqrt(3.0) + 1.0) / 2.0; }   // pts[], UnivariateFemStiffness.h:5

// rows of the scalar gradient matrix at (xi, eta): d/dx and d/dy of the 4 bilinear shapes (UnivariateFemStiffness.h:22-24)
__device__ __forceinline__ void grad_rows(double h, double xi, double eta, double r0[4], double r1[4]) {
  r0[0] = -1 / h * (1 - eta); r0[1] = 1 / h * (1 - eta); r0[2] = -1 / h * eta; r0[3] = 1 / h * eta;
  r1[0] = -1 / h * (1 - xi);  r1[1] = -1 / h * xi;       r1[2] = 1 / h * (1 - xi); r1[3] = 1 / h * xi;
}
// 3x8 strain-displacement matrix (FemStiffness.h:24-26)
__device__ __forceinline__ void bmat3x8(double h, double xi, double eta, double B[3][8]) {
  double r0[4], r1[4]; grad_rows(h, xi, eta, r0, r1);
#pragma unroll
  for (int c = 0; c < 4; c++) { B[0][c] = r0[c]; B[0][c + 4] = 0; B[1][c] = 0; B[1][c + 4] = r1[c]; B[2][c] = r1[c]; B[2][c + 4] = r0[c]; }
}
__device__ __forceinline__ void cell_nodes(int i, int j, int m, long long idx[4]) {
  idx[0] = (long long)j * (m + 1) + i; idx[1] = idx[0] + 1; idx[2] = (long long)(j + 1) * (m + 1) + i; idx[3] = idx[2] + 1;
}

// ---- UnivariateFemStiffness: one thread per (cell sequence, Gauss sequence) = 16 slots ------------------------
__global__ void k_quad_stiff1_fwd(const double* __restrict__ hmat, int rank3, int m, int n, double h, long long* __restrict__ ii,
                                  long long* __restrict__ jj, double* __restrict__ vv) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= 4LL * m * n) return;
  const int gs = (int)(t & 3), ei = gs >> 1, ej = gs & 1;          // loops: for ei; for ej (:32-33)
  const long long cs = t >> 2;
  const int i = (int)(cs / n), j = (int)(cs % n);                   // loops: for i<m; for j<n (:29-30)
  double r0[4], r1[4]; grad_rows(h, gpt(ei), gpt(ej), r0, r1);
  const double* K = rank3 ? hmat + 16 * ((long long)i + (long long)j * m) + 4 * (ei + ej * 2) : hmat;   // ids (:34)
  const double k00 = K[0], k01 = K[1], k10 = K[2], k11 = K[3], sc = 0.25 * h * h;
  long long idx[4]; cell_nodes(i, j, m, idx);
  const long long z0 = t * 16;
#pragma unroll
  for (int p = 0; p < 4; p++)
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const double kb0 = k00 * r0[q] + k01 * r1[q], kb1 = k10 * r0[q] + k11 * r1[q];
      vv[z0 + p * 4 + q] = (r0[p] * kb0 + r1[p] * kb1) * sc;
      if (ii) { ii[z0 + p * 4 + q] = idx[p] + 1; jj[z0 + p * 4 + q] = idx[q] + 1; }
    }
}
// per-Gauss K adjoint (backward, :80-124)
__global__ void k_quad_stiff1_bwd(const double* __restrict__ grad_vv, int m, int n, double h, double* __restrict__ grad_hmat) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= 4LL * m * n) return;
  const int gs = (int)(t & 3), ei = gs >> 1, ej = gs & 1;
  const long long cs = t >> 2;
  const int i = (int)(cs / n), j = (int)(cs % n);
  double r0[4], r1[4]; grad_rows(h, gpt(ei), gpt(ej), r0, r1);
  const double* g = grad_vv + t * 16;
  double d00 = 0, d01 = 0, d10 = 0, d11 = 0;
#pragma unroll
  for (int p = 0; p < 4; p++)
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const double v = g[p * 4 + q];
      d00 += r0[p] * v * r0[q]; d01 += r0[p] * v * r1[q]; d10 += r1[p] * v * r0[q]; d11 += r1[p] * v * r1[q];
    }
  const double sc = 0.25 * h * h;
  double* o = grad_hmat + 16 * ((long long)i + (long long)j * m) + 4 * (ei + ej * 2);
  o[0] = d00 * sc; o[1] = d01 * sc; o[2] = d10 * sc; o[3] = d11 * sc;
}

// ---- elasticity: FemStiffness (constant H, 64 slots per cell) / SpatialFemStiffness (per-Gauss H, 256 per cell) --
__global__ void k_quad_elast_const_fwd(const double* __restrict__ hmat, int m, int n, double h, long long* __restrict__ ii,
                                       long long* __restrict__ jj, double* __restrict__ vv) {
  __shared__ double Om[64];
  if (threadIdx.x < 64) {                                            // Omega is the same for every cell (FemStiffness.h:17-28)
    const int p = threadIdx.x >> 3, q = threadIdx.x & 7;
    double K[3][3];
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) K[a][b] = hmat[b * 3 + a];    // column-major read (quirk Q7)
    double acc = 0;
    for (int gi = 0; gi < 2; gi++) for (int gj = 0; gj < 2; gj++) {
      double B[3][8]; bmat3x8(h, gpt(gi), gpt(gj), B);
      double a2 = 0;
      for (int r = 0; r < 3; r++) for (int s = 0; s < 3; s++) a2 += B[r][p] * K[r][s] * B[s][q];
      acc += a2 * 0.25 * h * h;
    }
    Om[threadIdx.x] = acc;
  }
  __syncthreads();
  const long long total = 64LL * m * n, N = (long long)(m + 1) * (n + 1);
  for (long long z = blockIdx.x * (long long)blockDim.x + threadIdx.x; z < total; z += (long long)gridDim.x * blockDim.x) {
    const long long cs = z >> 6;
    const int pq = (int)(z & 63), p = pq >> 3, q = pq & 7, i = (int)(cs / n), j = (int)(cs % n);
    vv[z] = Om[pq];
    if (ii) { long long idx[4]; cell_nodes(i, j, m, idx); ii[z] = idx[p & 3] + (p >> 2) * N + 1; jj[z] = idx[q & 3] + (q >> 2) * N + 1; }
  }
}
// per-Gauss H: one thread per (cell sequence, Gauss sequence (p outer, q inner), output row r) = 8 slots
__global__ void k_quad_elast_gauss_fwd(const double* __restrict__ hmat, int m, int n, double h, long long* __restrict__ ii,
                                       long long* __restrict__ jj, double* __restrict__ vv) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= 32LL * m * n) return;
  const int r = (int)(t & 7), gs = (int)((t >> 3) & 3), p = gs >> 1, q = gs & 1, k = 2 * q + p;    // loops: for p; for q; k = 2q+p (:34-36)
  const long long cs = t >> 5;
  const int i = (int)(cs / n), j = (int)(cs % n);
  const long long elem = (long long)j * m + i, N = (long long)(m + 1) * (n + 1);
  double B[3][8]; bmat3x8(h, gpt(k >> 1), gpt(k & 1), B);            // Bs[k]: xi = pts[k/2], eta = pts[k%2] (quirk Q6, :18-26)
  const double* K = hmat + 36 * elem + 9 * k;
  double kb[3];                                                       // row r of B^T K
#pragma unroll
  for (int y = 0; y < 3; y++) kb[y] = B[0][r] * K[y] + B[1][r] * K[3 + y] + B[2][r] * K[6 + y];
  long long idx[4]; cell_nodes(i, j, m, idx);
  const long long z0 = (cs * 4 + gs) * 64 + r * 8;
#pragma unroll
  for (int s = 0; s < 8; s++) {
    vv[z0 + s] = (kb[0] * B[0][s] + kb[1] * B[1][s] + kb[2] * B[2][s]) * 0.25 * h * h;
    if (ii) { ii[z0 + s] = idx[r & 3] + (r >> 2) * N + 1; jj[z0 + s] = idx[s & 3] + (s >> 2) * N + 1; }
  }
}
// SFS_backward (:85-133): one thread per (cell sequence, Gauss sequence)
__global__ void k_quad_elast_gauss_bwd(const double* __restrict__ grad_vv, int m, int n, double h, double* __restrict__ grad_hmat) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= 4LL * m * n) return;
  const int gs = (int)(t & 3), p = gs >> 1, q = gs & 1, k = 2 * q + p;
  const long long cs = t >> 2;
  const int i = (int)(cs / n), j = (int)(cs % n);
  const long long elem = (long long)j * m + i;
  double B[3][8]; bmat3x8(h, gpt(k >> 1), gpt(k & 1), B);
  const double* G = grad_vv + t * 64;
  double bg[3][8];                                                    // B * G
#pragma unroll
  for (int x = 0; x < 3; x++)
#pragma unroll
    for (int s = 0; s < 8; s++) { double a = 0; for (int r = 0; r < 8; r++) a += B[x][r] * G[r * 8 + s]; bg[x][s] = a; }
  double* o = grad_hmat + 36 * elem + 9 * k;
#pragma unroll
  for (int x = 0; x < 3; x++)
#pragma unroll
    for (int y = 0; y < 3; y++) { double a = 0; for (int s = 0; s < 8; s++) a += bg[x][s] * B[y][s]; o[3 * x + y] = a * 0.25 * h * h; }
}

// ---- fixed-order reduction of `ncell` rows of 64 doubles -----------------------------------------------------------
constexpr int RED_BLOCKS = 592;   // 4 x 148 SMs
__global__ void k_reduce64_partial(const double* __restrict__ in, long long ncell, double* __restrict__ partial) {
  __shared__ double sh[4][64];
  const int k = threadIdx.x & 63, g = threadIdx.x >> 6;               // 256 threads: 4 groups of 64 components
  const long long per = (ncell + gridDim.x - 1) / gridDim.x, c0 = blockIdx.x * per, c1 = min(ncell, c0 + per);
  double a = 0;
  for (long long c = c0 + g; c < c1; c += 4) a += in[c * 64 + k];
  sh[g][k] = a;
  __syncthreads();
  if (g == 0) partial[blockIdx.x * 64 + k] = ((sh[0][k] + sh[1][k]) + sh[2][k]) + sh[3][k];
}
// final pass + the tiny dense contraction; mode 0: UnivariateFemStiffness constant K (4 outputs), 1: FemStiffness (9 outputs)
__global__ void k_reduce64_final(const double* __restrict__ partial, int nblocks, int mode, double h, double* __restrict__ out) {
  __shared__ double S[64];
  const int k = threadIdx.x;
  double a = 0;
  for (int b = 0; b < nblocks; b++) a += partial[b * 64 + k];
  S[k] = a;
  __syncthreads();
  const double sc = 0.25 * h * h;
  if (mode == 0 && k < 4) {                                          // backward2 (:199-247): sum over the 4 Gauss points, t = 2*ei+ej
    const int a2 = k >> 1, b2 = k & 1;
    double acc = 0;
    for (int t = 0; t < 4; t++) {
      double r0[4], r1[4]; grad_rows(h, gpt(t >> 1), gpt(t & 1), r0, r1);
      const double* ra = a2 == 0 ? r0 : r1; const double* rb = b2 == 0 ? r0 : r1;
      double d = 0;
      for (int p = 0; p < 4; p++) for (int q = 0; q < 4; q++) d += ra[p] * S[t * 16 + p * 4 + q] * rb[q];
      acc += d * sc;
    }
    out[k] = acc;
  } else if (mode == 1 && k < 9) {                                   // FS_backward (:73-109): grad_hmat[j*3+i] = dK(i,j)
    const int jj = k / 3, ii = k % 3;
    double acc = 0;
    for (int gi = 0; gi < 2; gi++) for (int gj = 0; gj < 2; gj++) {
      double B[3][8]; bmat3x8(h, gpt(gi), gpt(gj), B);
      double d = 0;
      for (int p = 0; p < 8; p++) for (int q = 0; q < 8; q++) d += B[ii][p] * S[p * 8 + q] * B[jj][q];
      acc += d * sc;
    }
    out[k] = acc;
  }
}

// ---- SpatialVaryingTangentElastic -----------------------------------------------------------------------------------
__global__ void k_svt_fwd(const double* __restrict__ mu, long long off, int type, double* __restrict__ hmat) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= off) return;
  double a = mu[i], d = type == 1 ? a : mu[i + off], b = type == 3 ? mu[i + 2 * off] : 0.0;
  double* o = hmat + 4 * i;
  o[0] = a; o[1] = b; o[2] = b; o[3] = d;
}
__global__ void k_svt_bwd(const double* __restrict__ gh, long long off, int type, double* __restrict__ gmu) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= off) return;
  const double* g = gh + 4 * i;
  if (type == 1) gmu[i] = g[0] + g[3];
  else { gmu[i] = g[0]; gmu[i + off] = g[3]; if (type == 3) gmu[i + 2 * off] = g[1] + g[2]; }
}

inline unsigned nblk(long long n, int bs) { return (unsigned)std::max<long long>(1, (n + bs - 1) / bs); }

int check_grid(int m, int n, double h) {
  if (m <= 0 || n <= 0 || !(h > 0)) return fail("structured grid: need m > 0, n > 0, h > 0");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) { cudaGetLastError(); return fail("no CUDA device available (libadfem_cuda has no CPU fallback)"); }
  return 0;
}

int reduce_cells(const double* grad_vv, long long ncell, int mode, double h, double* out, cudaStream_t st) {
  double* partial = nullptr;
  const int nb = (int)std::min<long long>(RED_BLOCKS, ncell);
  CU_TRY(cudaMallocAsync((void**)&partial, sizeof(double) * 64 * nb, st));
  k_reduce64_partial<<<nb, 256, 0, st>>>(grad_vv, ncell, partial);
  k_reduce64_final<<<1, 64, 0, st>>>(partial, nb, mode, h, out);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaFreeAsync(partial, st));
  return 0;
}

// ---- FemLaplace / FemMass / FemSource (quad_ops.cuh): one thread per (cell, Gauss point), one per node for the source gather ------
__global__ void k_quad_scalar_fwd(int op, const double* __restrict__ coef, int m, long long total, double h, long long* __restrict__ ii,
                                  long long* __restrict__ jj, double* __restrict__ vv) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < total) quad_scalar_fwd_body(op, t, coef, m, h, ii, jj, vv);
}
__global__ void k_quad_scalar_bwd(int op, const double* __restrict__ grad_vv, long long total, double h, double* __restrict__ grad_coef) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < total) grad_coef[t] = quad_scalar_bwd_body(op, t, grad_vv, h);
}
__global__ void k_quad_stiff1_svt_fwd(const double* __restrict__ mu, int type, int m, int n, double h, long long* __restrict__ ii,
                                      long long* __restrict__ jj, double* __restrict__ vv) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < 4LL * m * n) quad_stiff1_svt_fwd_body(t, mu, type, m, n, h, ii, jj, vv);
}
__global__ void k_quad_stiff1_svt_bwd(const double* __restrict__ grad_vv, int type, int m, int n, double h, double* __restrict__ grad_mu) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < 4LL * m * n) quad_stiff1_svt_bwd_body(t, grad_vv, type, m, n, h, grad_mu);
}
__global__ void k_quad_source_fwd(const double* __restrict__ f, int m, int n, double h, double* __restrict__ rhs) {
  const long long node = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (node < (long long)(m + 1) * (n + 1)) rhs[node] = quad_source_node(node, f, m, n, h);
}
__global__ void k_quad_source_bwd(const double* __restrict__ grad_rhs, int m, long long total, double h, double* __restrict__ grad_f) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < total) grad_f[t] = quad_source_bwd_body(t, grad_rhs, m, h);
}

}  // namespace

extern "C" {

int adfem_quad_stiffness1(const double* hmat, int rank3, int m, int n, double h, long long* ii, long long* jj, double* vv, void* stream) {
  if (int rc = check_grid(m, n, h)) return rc;
  if ((ii == nullptr) != (jj == nullptr)) return fail("ii and jj must both be given or both be NULL");
  k_quad_stiff1_fwd<<<nblk(4LL * m * n, 128), 128, 0, (cudaStream_t)stream>>>(hmat, rank3, m, n, h, ii, jj, vv);
  CU_TRY(cudaGetLastError());
  return 0;
}
int adfem_quad_stiffness1_grad(const double* grad_vv, int rank3, int m, int n, double h, double* grad_hmat, void* stream) {
  if (int rc = check_grid(m, n, h)) return rc;
  if (rank3) {
    k_quad_stiff1_bwd<<<nblk(4LL * m * n, 128), 128, 0, (cudaStream_t)stream>>>(grad_vv, m, n, h, grad_hmat);
    CU_TRY(cudaGetLastError());
    return 0;
  }
  return reduce_cells(grad_vv, (long long)m * n, 0, h, grad_hmat, (cudaStream_t)stream);
}
int adfem_quad_elasticity(const double* hmat, int per_gauss, int m, int n, double h, long long* ii, long long* jj, double* vv, void* stream) {
  if (int rc = check_grid(m, n, h)) return rc;
  if ((ii == nullptr) != (jj == nullptr)) return fail("ii and jj must both be given or both be NULL");
  if (per_gauss) k_quad_elast_gauss_fwd<<<nblk(32LL * m * n, 128), 128, 0, (cudaStream_t)stream>>>(hmat, m, n, h, ii, jj, vv);
  else k_quad_elast_const_fwd<<<(unsigned)std::min<long long>(nblk(64LL * m * n, 256), 148 * 16), 256, 0, (cudaStream_t)stream>>>(hmat, m, n, h, ii, jj, vv);
  CU_TRY(cudaGetLastError());
  return 0;
}
int adfem_quad_elasticity_grad(const double* grad_vv, int per_gauss, int m, int n, double h, double* grad_hmat, void* stream) {
  if (int rc = check_grid(m, n, h)) return rc;
  if (per_gauss) {
    k_quad_elast_gauss_bwd<<<nblk(4LL * m * n, 128), 128, 0, (cudaStream_t)stream>>>(grad_vv, m, n, h, grad_hmat);
    CU_TRY(cudaGetLastError());
    return 0;
  }
  return reduce_cells(grad_vv, (long long)m * n, 1, h, grad_hmat, (cudaStream_t)stream);
}
int adfem_svt(const double* mu, long long m, long long n, int type, double* hmat, void* stream) {
  if (type < 1 || type > 3) return fail("SpatialVaryingTangentElastic: type must be 1, 2 or 3");
  if (int rc = check_grid((int)m, (int)n, 1.0)) return rc;
  k_svt_fwd<<<nblk(4 * m * n, 256), 256, 0, (cudaStream_t)stream>>>(mu, 4 * m * n, type, hmat);
  CU_TRY(cudaGetLastError());
  return 0;
}
int adfem_svt_grad(const double* grad_hmat, long long m, long long n, int type, double* grad_mu, void* stream) {
  if (type < 1 || type > 3) return fail("SpatialVaryingTangentElastic: type must be 1, 2 or 3");
  if (int rc = check_grid((int)m, (int)n, 1.0)) return rc;
  k_svt_bwd<<<nblk(4 * m * n, 256), 256, 0, (cudaStream_t)stream>>>(grad_hmat, 4 * m * n, type, grad_mu);
  CU_TRY(cudaGetLastError());
  return 0;
}

int adfem_quad_scalar(int op, const double* coef, long long m, long long n, double h, long long* ii, long long* jj, double* vv, void* stream) {
  if (op != 0 && op != 1) return fail("adfem_quad_scalar: op must be 0 (FemLaplace) or 1 (FemMass)");
  if (int rc = check_grid((int)m, (int)n, h)) return rc;
  if ((ii == nullptr) != (jj == nullptr)) return fail("ii and jj must both be given or both be NULL");
  k_quad_scalar_fwd<<<nblk(4 * m * n, 128), 128, 0, (cudaStream_t)stream>>>(op, coef, (int)m, 4 * m * n, h, ii, jj, vv);
  CU_TRY(cudaGetLastError());
  return 0;
}
int adfem_quad_scalar_grad(int op, const double* grad_vv, long long m, long long n, double h, double* grad_coef, void* stream) {
  if (op != 0 && op != 1) return fail("adfem_quad_scalar_grad: op must be 0 (FemLaplace) or 1 (FemMass)");
  if (int rc = check_grid((int)m, (int)n, h)) return rc;
  k_quad_scalar_bwd<<<nblk(4 * m * n, 128), 128, 0, (cudaStream_t)stream>>>(op, grad_vv, 4 * m * n, h, grad_coef);
  CU_TRY(cudaGetLastError());
  return 0;
}
int adfem_quad_source(const double* f, long long m, long long n, double h, double* rhs, void* stream) {
  if (int rc = check_grid((int)m, (int)n, h)) return rc;
  k_quad_source_fwd<<<nblk((m + 1) * (n + 1), 128), 128, 0, (cudaStream_t)stream>>>(f, (int)m, (int)n, h, rhs);
  CU_TRY(cudaGetLastError());
  return 0;
}
int adfem_quad_source_grad(const double* grad_rhs, long long m, long long n, double h, double* grad_f, void* stream) {
  if (int rc = check_grid((int)m, (int)n, h)) return rc;
  k_quad_source_bwd<<<nblk(4 * m * n, 128), 128, 0, (cudaStream_t)stream>>>(grad_rhs, (int)m, 4 * m * n, h, grad_f);
  CU_TRY(cudaGetLastError());
  return 0;
}

int adfem_quad_stiffness1_svt(const double* mu, int type, int m, int n, double h, long long* ii, long long* jj, double* vv, void* stream) {
  if (type < 1 || type > 3) return fail("SpatialVaryingTangentElastic: type must be 1, 2 or 3");
  if (int rc = check_grid(m, n, h)) return rc;
  if ((ii == nullptr) != (jj == nullptr)) return fail("ii and jj must both be given or both be NULL");
  k_quad_stiff1_svt_fwd<<<nblk(4LL * m * n, 128), 128, 0, (cudaStream_t)stream>>>(mu, type, m, n, h, ii, jj, vv);
  CU_TRY(cudaGetLastError());
  return 0;
}
int adfem_quad_stiffness1_svt_grad(const double* grad_vv, int type, int m, int n, double h, double* grad_mu, void* stream) {
  if (type < 1 || type > 3) return fail("SpatialVaryingTangentElastic: type must be 1, 2 or 3");
  if (int rc = check_grid(m, n, h)) return rc;
  k_quad_stiff1_svt_bwd<<<nblk(4LL * m * n, 128), 128, 0, (cudaStream_t)stream>>>(grad_vv, type, m, n, h, grad_mu);
  CU_TRY(cudaGetLastError());
  return 0;
}

}  // extern "C"
