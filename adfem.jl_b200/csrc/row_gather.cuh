// Row-gather forward for the scalar operators (Laplace, mass) on ANY mesh, written for the P2 case (BASELINE config 4), option "row_gather".
// One THREAD per dof row: it walks the dof -> (element, local dof) adjacency of the symbolic phase in ascending element order (the summation
// order of the tile kernels), evaluates ROW p of each incident element's local matrix in registers, finds the CSR position of every column
// dof by binary search in the row's (sorted) column indices and accumulates into the CTA's shared-memory copy of its rows; the CTA then writes
// its rows — one contiguous run of the values array — with coalesced stores.  No tile plan is needed (the P2 tile blobs are 330 B per element,
// more than the algorithmic traffic): the mesh-static inputs are the adjacency (5 B per incidence), the connectivity and the CSR pattern.
// A local matrix row is evaluated once per dof of the element (d times the geometry / basis work of the tile kernels, ~3 kflop per P2 triangle,
// far below the fp64 roof at this traffic).  Rows are limited by the shared-memory staging: RG_CAP values per CTA of 128, 64 or 32 rows (the
// host picks the largest count whose CTAs all fit; P2 tetrahedra run with 32).
// Host + device bodies (tests/host_emul/).
#pragma once
#include "device_fem.cuh"

namespace adfem {

constexpr int RG_THREADS = 128;
constexpr int RG_CAP = 5120;             // doubles of shared memory per CTA (40 KB): the CSR entries of its 128 rows must fit

// row p of the local matrix of one element: row[q] = sum_k coef_k w_k (grad phi_p . grad phi_q | phi_p phi_q)
template <int DIM, int DEG, int OP>
ADFEM_HD void rg_local_row(const Geom<DIM>& G, const QuadRule& rule, int g, int p, const double* ce, double* row) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
#pragma unroll
  for (int q = 0; q < D; q++) row[q] = 0.0;
  for (int k = 0; k < g; k++) {
    double L[DIM + 1]; bary<DIM>(rule, k, L);
    const double cw = ldg(ce + k) * (rule.w[k] * G.wscale);
    if (OP == OP_LAPLACE) {
      double gp[D][DIM]; basis_grad<DIM, DEG>(G, L, gp);
      double gs[DIM];
#pragma unroll
      for (int i = 0; i < DIM; i++) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < D; q++) v = (q == p) ? gp[q][i] : v;
        gs[i] = v;
      }
#pragma unroll
      for (int q = 0; q < D; q++) row[q] += dotg<DIM>(gs, gp[q]) * cw;
    } else {
      double phi[D]; basis_val<DIM, DEG>(L, phi);
      double ps = 0.0;
#pragma unroll
      for (int q = 0; q < D; q++) ps = (q == p) ? phi[q] : ps;
#pragma unroll
      for (int q = 0; q < D; q++) row[q] += ps * phi[q] * cw;
    }
  }
}

// phase 1, thread <-> row r (rows of the CTA start at row r0, their entries at rs0): accumulate the row into acc[rowptr[r] - rs0 + j]
template <int DIM, int DEG, int OP>
ADFEM_HD void rg_row(const DevMesh& m, const long long* adj_ptr, const int* adj_elem, const uint8_t* adj_loc, const long long* rowptr, const int* colind,
                     int r, long long rs0, const double* coef, double* acc) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  const long long rs = rowptr[r];
  const int len = (int)(rowptr[r + 1] - rs);
  double* a = acc + (rs - rs0);
  for (int j = 0; j < len; j++) a[j] = 0.0;
  const int* cols = colind + rs;
  for (long long t = adj_ptr[r]; t < adj_ptr[r + 1]; t++) {
    const int e = adj_elem[t], p = adj_loc[t];
    Geom<DIM> G; load_geom(m, e, G);
    double row[D];
    rg_local_row<DIM, DEG, OP>(G, m.rule, m.g, p, coef + (size_t)e * m.g, row);
#pragma unroll
    for (int q = 0; q < D; q++) {
      const int c = ldg(m.conn + (size_t)q * m.ne + e);
      int lo = 0, hi = len;                                   // first position with cols[pos] >= c
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (ldg(cols + mid) < c) lo = mid + 1; else hi = mid; }
      a[lo] += row[q];
    }
  }
}

// ---- P1 elasticity (any mesh): the same row walk for the NC x NC block operator, from the Gauss-summed tangents of option "coef_presum" ---------
// hbar[e*NS*NS + ...] = sum_k w_k H_{e,k} (k_presum_coef).  Thread <-> scalar row r; it fills the NC*NC blocks of its row in the CTA's copy of
// the NC component runs: acc[a*NC*T + NC*(rs - rs0) + b*len + j], T = CSR entries of the CTA's rows, i.e. the layout of adfem_assemble_csr.
constexpr int RGE_THREADS = 64;

template <int DIM>
ADFEM_HD void rg_row_elast(const DevMesh& m, const long long* adj_ptr, const int* adj_elem, const uint8_t* adj_loc, const long long* rowptr,
                           const int* colind, int r, long long rs0, int T, const double* hbar, double* acc) {
  constexpr int NC = DIM, D = DIM + 1, NS = Voigt<DIM>::NS;
  const long long rs = rowptr[r];
  const int len = (int)(rowptr[r + 1] - rs), base = NC * (int)(rs - rs0);
  for (int a = 0; a < NC; a++)
    for (int t = 0; t < NC * len; t++) acc[a * NC * T + base + t] = 0.0;
  const int* cols = colind + rs;
  for (long long t = adj_ptr[r]; t < adj_ptr[r + 1]; t++) {
    const int e = adj_elem[t], p = adj_loc[t];
    Geom<DIM> G; load_geom(m, e, G);
    double H[NS * NS];
#pragma unroll
    for (int c = 0; c < NS * NS; c++) H[c] = ldg(hbar + (size_t)e * (NS * NS) + c) * G.wscale;
    double gp[DIM];
#pragma unroll
    for (int c = 0; c < DIM; c++) {
      double v = 0.0;
#pragma unroll
      for (int q = 0; q < D; q++) v = (q == p) ? G.gL[q][c] : v;
      gp[c] = v;
    }
#pragma unroll
    for (int q = 0; q < D; q++) {
      const int c = ldg(m.conn + (size_t)q * m.ne + e);
      int lo = 0, hi = len;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (ldg(cols + mid) < c) lo = mid + 1; else hi = mid; }
#pragma unroll
      for (int b = 0; b < NC; b++) {
        double hb[NS];
#pragma unroll
        for (int i = 0; i < NS; i++) hb[i] = bdot<DIM>(b, G.gL[q], &H[NS * i]);
#pragma unroll
        for (int a = 0; a < NC; a++) acc[a * NC * T + base + b * len + lo] += bdot<DIM>(a, gp, hb);
      }
    }
  }
}

#ifdef __CUDACC__
// blockDim.x = rows per CTA (128, 64 or 32: the largest whose CTAs all fit RG_CAP, chosen on the host from the pattern)
template <int DIM, int DEG, int OP>
__global__ void __launch_bounds__(RG_THREADS) k_row_gather_fwd(DevMesh m, const long long* __restrict__ adj_ptr, const int* __restrict__ adj_elem,
                                                                const uint8_t* __restrict__ adj_loc, const long long* __restrict__ rowptr,
                                                                const int* __restrict__ colind, const double* __restrict__ coef, double* __restrict__ vals) {
  __shared__ double acc[RG_CAP];
  const int rows = blockDim.x, r0 = blockIdx.x * rows, r = r0 + threadIdx.x, r1 = min(r0 + rows, m.ndof);
  const long long rs0 = rowptr[r0];
  if (r < m.ndof) rg_row<DIM, DEG, OP>(m, adj_ptr, adj_elem, adj_loc, rowptr, colind, r, rs0, coef, acc);
  __syncthreads();
  const int total = (int)(rowptr[r1] - rs0);
  for (int idx = threadIdx.x; idx < total; idx += rows) vals[rs0 + idx] = acc[idx];
}

template <int DIM>
__global__ void __launch_bounds__(RGE_THREADS) k_row_gather_elast_fwd(DevMesh m, const long long* __restrict__ adj_ptr, const int* __restrict__ adj_elem,
                                                                       const uint8_t* __restrict__ adj_loc, const long long* __restrict__ rowptr,
                                                                       const int* __restrict__ colind, long long nnz, const double* __restrict__ hbar,
                                                                       double* __restrict__ vals) {
  extern __shared__ __align__(16) double rge_acc[];
  constexpr int NC = DIM;
  const int r0 = blockIdx.x * RGE_THREADS, r = r0 + threadIdx.x, r1 = min(r0 + RGE_THREADS, m.ndof);
  const long long rs0 = rowptr[r0];
  const int T = (int)(rowptr[r1] - rs0);
  if (r < m.ndof) rg_row_elast<DIM>(m, adj_ptr, adj_elem, adj_loc, rowptr, colind, r, rs0, T, hbar, rge_acc);
  __syncthreads();
  for (int a = 0; a < NC; a++) {
    double* out = vals + NC * ((long long)a * nnz + rs0);
    for (int idx = threadIdx.x; idx < NC * T; idx += RGE_THREADS) out[idx] = rge_acc[a * NC * T + idx];
  }
}
#endif

}  // namespace adfem
