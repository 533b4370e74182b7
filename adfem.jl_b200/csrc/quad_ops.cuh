// Structured-grid Q1 scalar operators on an m x n grid of h x h cells (SURVEY §8 f, rank 4):
//   FemLaplace  deps/FemLaplace/FemLaplace.h:11-82   compute_fem_laplace_matrix1(K, m, n, h)  src/InvCore.jl:443-449
//   FemMass     deps/FemMass/FemMass.h:10-79         compute_fem_mass_matrix1(rho, m, n, h)   src/InvCore.jl:364-369
//   FemSource   deps/FemSource/FemSource.h:8-47      compute_fem_source_term1(f, m, n, h)     src/InvCore.jl:307-312
// Cell (i, j) has id j*m + i and nodes j(m+1)+i, +1, (j+1)(m+1)+i, +1; Gauss point k = 2q + p sits at (xi, eta) = (pts[p], pts[q]);
// coefficient arrays are indexed 4*cell + k; the COO slot of (cell, k, a, b) is (4*cell + k)*16 + 4a + b and ii / jj are 0-BASED
// (the Julia wrappers add 1, src/InvCore.jl:368,448) — unlike the stiffness siblings of grid_ops.cu, which emit 1-based indices.
// Bodies are __host__ __device__ (tests/host_emul/ runs them on the host against the oracle).
#pragma once
#include "device_fem.cuh"

namespace adfem {

ADFEM_HD double quad_pt(int i) { return i == 0 ? (-1 / sqrt(3.0) + 1.0) / 2.0 : (1 / sqrt(3.0) + 1.0) / 2.0; }   // pts[], FemLaplace.h:8

// the 4 bilinear shapes at (xi, eta) (FemMass.h:17) and the rows of their gradient (FemLaplace.h:17-19)
ADFEM_HD void quad_shapes(double xi, double eta, double A[4]) { A[0] = (1 - xi) * (1 - eta); A[1] = xi * (1 - eta); A[2] = (1 - xi) * eta; A[3] = xi * eta; }
ADFEM_HD void quad_grad_rows(double h, double xi, double eta, double r0[4], double r1[4]) {
  r0[0] = -1 / h * (1 - eta); r0[1] = 1 / h * (1 - eta); r0[2] = -1 / h * eta; r0[3] = 1 / h * eta;
  r1[0] = -1 / h * (1 - xi);  r1[1] = -1 / h * xi;       r1[2] = 1 / h * (1 - xi); r1[3] = 1 / h * xi;
}
// local 4x4 matrix of Gauss point k: op 0 = B^T B h^2/4 (Laplace), op 1 = A A^T h^2/4 (mass)
ADFEM_HD void quad_local(int op, double h, int k, double M[16]) {
  const double xi = quad_pt(k & 1), eta = quad_pt(k >> 1), sc = 0.25 * h * h;
  if (op == 0) {
    double r0[4], r1[4]; quad_grad_rows(h, xi, eta, r0, r1);
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) M[4 * a + b] = (r0[a] * r0[b] + r1[a] * r1[b]) * sc;
  } else {
    double A[4]; quad_shapes(xi, eta, A);
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) M[4 * a + b] = A[a] * A[b] * sc;
  }
}

// forward, t = 4*cell + k: 16 slots
ADFEM_HD void quad_scalar_fwd_body(int op, long long t, const double* coef, int m, double h, long long* ii, long long* jj, double* vv) {
  const long long cell = t >> 2;
  const int k = (int)(t & 3), i = (int)(cell % m);
  const long long j = cell / m;
  double M[16]; quad_local(op, h, k, M);
  const double c = ldg(coef + t);
  long long idx[4];
  idx[0] = j * (m + 1) + i; idx[1] = idx[0] + 1; idx[2] = (j + 1) * (m + 1) + i; idx[3] = idx[2] + 1;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      vv[t * 16 + 4 * a + b] = c * M[4 * a + b];
      if (ii) { ii[t * 16 + 4 * a + b] = idx[a]; jj[t * 16 + 4 * a + b] = idx[b]; }
    }
}
// adjoint: grad_coef[t] = sum_ab M_ab grad_vv[slot] (FemLaplace.h:64-77, FemMass.h:62-72)
ADFEM_HD double quad_scalar_bwd_body(int op, long long t, const double* grad_vv, double h) {
  double M[16]; quad_local(op, h, (int)(t & 3), M);
  double s = 0.0;
#pragma unroll
  for (int ab = 0; ab < 16; ab++) s += M[ab] * ldg(grad_vv + t * 16 + ab);
  return s;
}

// FemSource forward as a gather: node (ni, nj) sums its (up to) four cells in the reference's order of arrival (cell loop i outer, j inner;
// within a cell p outer, q inner — FemSource.h:9-24)
ADFEM_HD double quad_source_node(long long node, const double* f, int m, int n, double h) {
  const int ni = (int)(node % (m + 1));
  const long long nj = node / (m + 1);
  double acc = 0.0;
  for (int di = 1; di >= 0; di--)          // cell column i = ni-1 (node is its right side), then ni
    for (int dj = 1; dj >= 0; dj--) {      // cell row j = nj-1 (node is its top side), then nj
      const int i = ni - di;
      const long long j = nj - dj;
      if (i < 0 || i >= m || j < 0 || j >= n) continue;
      const long long cell = j * m + i;
      for (int p = 0; p < 2; p++)
        for (int q = 0; q < 2; q++) {
          const double xi = quad_pt(p), eta = quad_pt(q);
          const double val1 = ldg(f + cell * 4 + 2 * q + p) * h * h * 0.25;
          acc += val1 * (di ? xi : 1 - xi) * (dj ? eta : 1 - eta);
        }
    }
  return acc;
}
// FemSource adjoint, t = 4*cell + k (FemSource.h:30-46)
ADFEM_HD double quad_source_bwd_body(long long t, const double* grad_rhs, int m, double h) {
  const long long cell = t >> 2;
  const int k = (int)(t & 3), i = (int)(cell % m);
  const long long j = cell / m;
  const double xi = quad_pt(k & 1), eta = quad_pt(k >> 1), sc = h * h * 0.25;
  const long long n0 = j * (m + 1) + i, n2 = (j + 1) * (m + 1) + i;
  double s = 0.0;
  s += sc * (1 - xi) * (1 - eta) * ldg(grad_rhs + n0);
  s += sc * xi * (1 - eta) * ldg(grad_rhs + n0 + 1);
  s += sc * (1 - xi) * eta * ldg(grad_rhs + n2);
  s += sc * xi * eta * ldg(grad_rhs + n2 + 1);
  return s;
}

// ---- SpatialVaryingTangentElastic fused into UnivariateFemStiffness ---------------------------------------------
// compute_fem_stiffness_matrix1(compute_space_varying_tangent_elasticity_matrix(mu, m, n, h, type), m, n, h) in one pass: the 4mn x 2 x 2 tensor
// is never written.  K of Gauss point gi (deps/SpatialVaryingTangentElastic/SpatialVaryingTangentElastic.h:1-31): type 1 mu_gi I,
// type 2 diag(mu_gi, mu_{gi+4mn}), type 3 [[mu_gi, mu_{gi+8mn}], [mu_{gi+8mn}, mu_{gi+4mn}]].
// Slot order and 1-BASED ii / jj of UnivariateFemStiffness (deps/FemStiffness1/UnivariateFemStiffness.h:29-43): t = 4*cs + (2*ei + ej), cell
// sequence cs = i*n + j (i outer), xi = pts[ei], eta = pts[ej], K index gi = 4*(i + j*m) + ei + 2*ej.
ADFEM_HD void quad_stiff1_svt_fwd_body(long long t, const double* mu, int type, int m, int n, double h, long long* ii, long long* jj, double* vv) {
  const int gs = (int)(t & 3), ei = gs >> 1, ej = gs & 1;
  const long long cs = t >> 2, off = 4LL * m * n;
  const int i = (int)(cs / n), j = (int)(cs % n);
  const long long gi = 4 * ((long long)i + (long long)j * m) + ei + 2 * ej;
  const double k00 = ldg(mu + gi), k11 = type == 1 ? k00 : ldg(mu + gi + off), k01 = type == 3 ? ldg(mu + gi + 2 * off) : 0.0;
  double r0[4], r1[4]; quad_grad_rows(h, quad_pt(ei), quad_pt(ej), r0, r1);
  const double sc = 0.25 * h * h;
  long long idx[4];
  idx[0] = (long long)j * (m + 1) + i; idx[1] = idx[0] + 1; idx[2] = (long long)(j + 1) * (m + 1) + i; idx[3] = idx[2] + 1;
#pragma unroll
  for (int p = 0; p < 4; p++)
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const double kb0 = k00 * r0[q] + k01 * r1[q], kb1 = k01 * r0[q] + k11 * r1[q];
      vv[t * 16 + p * 4 + q] = (r0[p] * kb0 + r1[p] * kb1) * sc;
      if (ii) { ii[t * 16 + p * 4 + q] = idx[p] + 1; jj[t * 16 + p * 4 + q] = idx[q] + 1; }
    }
}
// adjoint: grad_mu (length 4mn * type), every entry written exactly once
ADFEM_HD void quad_stiff1_svt_bwd_body(long long t, const double* grad_vv, int type, int m, int n, double h, double* grad_mu) {
  const int gs = (int)(t & 3), ei = gs >> 1, ej = gs & 1;
  const long long cs = t >> 2, off = 4LL * m * n;
  const int i = (int)(cs / n), j = (int)(cs % n);
  const long long gi = 4 * ((long long)i + (long long)j * m) + ei + 2 * ej;
  double r0[4], r1[4]; quad_grad_rows(h, quad_pt(ei), quad_pt(ej), r0, r1);
  double d00 = 0, d01 = 0, d10 = 0, d11 = 0;
#pragma unroll
  for (int p = 0; p < 4; p++)
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const double v = ldg(grad_vv + t * 16 + p * 4 + q);
      d00 += r0[p] * v * r0[q]; d01 += r0[p] * v * r1[q]; d10 += r1[p] * v * r0[q]; d11 += r1[p] * v * r1[q];
    }
  const double sc = 0.25 * h * h;
  if (type == 1) grad_mu[gi] = (d00 + d11) * sc;
  else {
    grad_mu[gi] = d00 * sc; grad_mu[gi + off] = d11 * sc;
    if (type == 3) grad_mu[gi + 2 * off] = (d01 + d10) * sc;
  }
}

}  // namespace adfem
