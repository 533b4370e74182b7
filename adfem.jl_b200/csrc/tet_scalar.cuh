// CSR assembly / adjoint of the SCALAR P1 operators (FemLaplaceScalarT, ComputeFemMassMatrixMfemT — the reference's own 3-D ops, deps/MFEM3) on
// the structured tetrahedral grid `Mesh3(n, n, l, h)`, option "structured_elasticity" (the switch of the opt-in structured kernels).
// Forward: one thread per node walks its 8 or 32 incident tetrahedra (tet_grid_tables.h) in ascending element order, evaluates row p of each
// local matrix in registers and accumulates into the CTA's shared-memory copy of its 128 rows at positions given by the 27-bit neighbour mask;
// the CTA writes its rows as one contiguous run.  Adjoint: one thread per tetrahedron gathers its 16 upstream values at computed CSR positions.
// No adjacency, connectivity, column indices or coordinates are read (row pointers: 8 B per node).  Host + device bodies (tests/host_emul/).
#pragma once
#include "row_gather.cuh"
#include "tet_gauss.cuh"

namespace adfem {

// forward, thread <-> node (i, j, k) = row r; rows of the CTA start at entry rs0
template <int OP>
ADFEM_HD void tgs_row(const GridTet& gt, const QuadRule& rule, int g, int i, int j, int k, long long rs, long long rs0, const double* coef, double* acc) {
  const int par = (i + j + k) & 1, mask = tg_row_mask(gt, par, i, j, k), len = tg_popc(mask);
  const long long n1 = gt.n + 1;
  double* a = acc + (rs - rs0);
  for (int t = 0; t < len; t++) a[t] = 0.0;
  for (int t = 0; t < gt.tab->ninc[par]; t++) {
    Geom<3> G; int p; long long e, v[4];
    if (!tg_incident(gt, par, t, i, j, k, G, p, e, v)) continue;
    double row[4];
    rg_local_row<3, 1, OP>(G, rule, g, p, coef + (size_t)e * g, row);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int vk = (int)(v[q] / (n1 * n1)), vj = (int)((v[q] / n1) % n1), vi = (int)(v[q] % n1);
      const int s = (vk - k + 1) * 9 + (vj - j + 1) * 3 + (vi - i + 1);
      a[tg_popc(mask & ((1 << s) - 1))] += row[q];
    }
  }
}

// adjoint, thread <-> tetrahedron e: grad_coef[e*g + k]
template <int OP>
ADFEM_HD void tgs_tet_adjoint(const GridTet& gt, const QuadRule& rule, int g, long long e, const long long* rowptr, const double* dvals, double* grad) {
  const TetGridTables& T = *gt.tab;
  const long long cube = e / 5, n1 = gt.n + 1;
  const int t = (int)(e - 5 * cube), ck = (int)(cube % gt.l), cj = (int)((cube / gt.l) % gt.n), ci = (int)(cube / ((long long)gt.l * gt.n));
  const int ev = ((ci + cj + ck + 3) & 1) == 0;
  int vi[4], vj[4], vk[4];
  double X[4][3];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int v = T.te[ev][t][q];
    vi[q] = ci + (v & 1); vj[q] = cj + ((v >> 1) & 1); vk[q] = ck + (v >> 2);
    X[q][0] = ldg(gt.xs + vi[q]); X[q][1] = ldg(gt.ys + vj[q]); X[q][2] = ldg(gt.zs + vk[q]);
  }
  Geom<3> G; geom_tet(X, G);
  if (G.wscale < 0) {                                          // orientation fix: swap local vertices 0 and 1 (Gauss points follow the mesh's order)
#pragma unroll
    for (int c = 0; c < 3; c++) { const double x = X[0][c]; X[0][c] = X[1][c]; X[1][c] = x; }
    int s;
    s = vi[0]; vi[0] = vi[1]; vi[1] = s; s = vj[0]; vj[0] = vj[1]; vj[1] = s; s = vk[0]; vk[0] = vk[1]; vk[1] = s;
    geom_tet(X, G);
  }
  double dK[4][4];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int mask = tg_row_mask(gt, (vi[p] + vj[p] + vk[p]) & 1, vi[p], vj[p], vk[p]);
    const double* row = dvals + rowptr[((long long)vk[p] * n1 + vj[p]) * n1 + vi[p]];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int s = (vk[q] - vk[p] + 1) * 9 + (vj[q] - vj[p] + 1) * 3 + (vi[q] - vi[p] + 1);
      dK[p][q] = ldg(row + tg_popc(mask & ((1 << s) - 1)));
    }
  }
  for (int k = 0; k < g; k++) {
    double L[4]; bary<3>(rule, k, L);
    double s = 0.0;
    if (OP == OP_LAPLACE) {
#pragma unroll
      for (int p = 0; p < 4; p++)
#pragma unroll
        for (int q = 0; q < 4; q++) s += dK[p][q] * dotg<3>(G.gL[p], G.gL[q]);
    } else {
#pragma unroll
      for (int p = 0; p < 4; p++)
#pragma unroll
        for (int q = 0; q < 4; q++) s += dK[p][q] * (L[p] * L[q]);
    }
    grad[(size_t)e * g + k] = s * (rule.w[k] * G.wscale);
  }
}

#ifdef __CUDACC__
template <int OP>
__global__ void __launch_bounds__(RG_THREADS) k_tet_grid_scalar_fwd(GridTet gt, QuadRule rule, int g, const long long* __restrict__ rowptr,
                                                                     const double* __restrict__ coef, double* __restrict__ vals) {
  __shared__ double acc[RG_CAP];
  const long long n1 = gt.n + 1, nn = n1 * n1 * (gt.l + 1);
  const long long r0 = (long long)blockIdx.x * RG_THREADS, r = r0 + threadIdx.x, r1 = r0 + RG_THREADS < nn ? r0 + RG_THREADS : nn;
  const long long rs0 = rowptr[r0];
  if (r < nn) tgs_row<OP>(gt, rule, g, (int)(r % n1), (int)((r / n1) % n1), (int)(r / (n1 * n1)), rowptr[r], rs0, coef, acc);
  __syncthreads();
  const int total = (int)(rowptr[r1] - rs0);
  for (int idx = threadIdx.x; idx < total; idx += RG_THREADS) vals[rs0 + idx] = acc[idx];
}
template <int OP>
__global__ void __launch_bounds__(128) k_tet_grid_scalar_adj(GridTet gt, QuadRule rule, int g, long long ne, const long long* __restrict__ rowptr,
                                                             const double* __restrict__ dvals, double* __restrict__ grad) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < ne) tgs_tet_adjoint<OP>(gt, rule, g, e, rowptr, dvals, grad);
}
#endif

}  // namespace adfem
