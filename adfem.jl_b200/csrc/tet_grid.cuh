// Forward CSR assembly of the P1 tetrahedral ELASTICITY operator (3-D extension of ComputeFemStiffnessMatrixMfem, 6x6 Voigt tangent per Gauss
// point) on the reference's structured grid `Mesh3(n, n, l, h)` — BASELINE config 5 — without tile blobs, connectivity or coordinates.
//
// Input: the Gauss-summed tangents Hbar_e = sum_k w_k H_{e,k} (36 doubles per tetrahedron) produced by the streaming pre-pass of option
// "coef_presum" (gauss_ops.cu, k_presum_coef).  One WARP per node:
//   phase 1  lane t evaluates the 3 x 12 row block of the node in its t-th incident tetrahedron (8 or 32 of them, by the parity of i+j+k,
//            tet_grid_tables.h) from Hbar and the rectilinear coordinates, and parks it in shared memory;
//   phase 2  the (up to) 19 x 9 values of the node's three CSR rows are gathered from those blocks in ascending element order (the summation
//            order of the general tile kernels) into CSR order;
//   phase 3  the three contiguous runs (component a = 0, 1, 2) are written with coalesced stores.
// Row pointers are read from the scalar pattern (8 B per node); column positions are index arithmetic.
// CSR layout as everywhere: scalar row r (start rs, length len) holds entry (a, b, j) at 3*(a*nnz + rs) + b*len + j.
// Adjoint: one warp per 32 consecutive tetrahedra (a contiguous run of the gradient array): lane t gathers the 144 upstream values of its
// tetrahedron at computed CSR positions, forms B dK B^T |det| (36 values) into shared memory, and the warp writes the 32 x 4 x 36 gradients
// (times the Gauss weights) coalesced.
// All phases are __host__ __device__ functions of the lane index (tests/host_emul/).
#pragma once
#include "device_fem.cuh"
#include "tet_grid_tables.h"

namespace adfem {

struct GridTet {
  int n, l;                 // cubes in x and y (n) and in z (l); node (i, j, k) = k*(n+1)^2 + j*(n+1) + i, cube (ci, cj, ck) = (ci*n + cj)*l + ck
  const double* xs;         // n+1, n+1, l+1 node coordinates per axis
  const double* ys;
  const double* zs;
  const TetGridTables* tab;
};

constexpr int TG_BLK = 36;                        // row block of one tetrahedron: [a][q][b]
constexpr int TG_WARP_DOUBLES = 32 * TG_BLK + 3 * 27 * 3;
constexpr int TG_WARPS = 8;

// 27-bit mask of the row slots of node (i, j, k) that exist: structurally present for its parity and inside the grid (slot = (dk+1)*9 + (dj+1)*3
// + (di+1); the in-range part is a product of three 3-bit masks)
ADFEM_HD int tg_row_mask(const GridTet& gt, int par, int i, int j, int k) {
  const int mx = (i > 0 ? 1 : 0) | 2 | (i < gt.n ? 4 : 0), my = (j > 0 ? 1 : 0) | 2 | (j < gt.n ? 4 : 0), mz = (k > 0 ? 1 : 0) | 2 | (k < gt.l ? 4 : 0);
  const int row9 = ((my & 1) ? mx : 0) | (mx << 3) | ((my & 4) ? mx << 6 : 0);
  const int in27 = ((mz & 1) ? row9 : 0) | (row9 << 9) | ((mz & 4) ? row9 << 18 : 0);
  return in27 & gt.tab->present[par];
}
#ifdef __CUDA_ARCH__
ADFEM_HD int tg_popc(int x) { return __popc((unsigned)x); }
#else
ADFEM_HD int tg_popc(int x) { int c = 0; for (; x; x &= x - 1) c++; return c; }
#endif

// phase 1: lane t -> tb[t*36 + (a*4 + q)*3 + b] = bvec(a, g_p)^T (Hbar |det|) bvec(b, g_q); zeros when the tetrahedron does not exist
ADFEM_HD void tg_tet_block(int lane, const GridTet& gt, int par, int i, int j, int k, const double* hbar, double* tb) {
  const TetGridTables& T = *gt.tab;
  double* out = tb + lane * TG_BLK;
  const int ci = i + T.inc[par][lane][0], cj = j + T.inc[par][lane][1], ck = k + T.inc[par][lane][2];
  if (lane >= T.ninc[par] || ci < 0 || ci >= gt.n || cj < 0 || cj >= gt.n || ck < 0 || ck >= gt.l) {
    for (int c = 0; c < TG_BLK; c++) out[c] = 0.0;
    return;
  }
  double X[4][3];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    X[q][0] = ldg(gt.xs + i + T.voff[par][lane][q][0]);
    X[q][1] = ldg(gt.ys + j + T.voff[par][lane][q][1]);
    X[q][2] = ldg(gt.zs + k + T.voff[par][lane][q][2]);
  }
  Geom<3> G; geom_tet(X, G);
  const double ws = G.wscale < 0 ? -G.wscale : G.wscale;          // the table order is the generator's, before MFEM's orientation fix
  const double* he = hbar + ((size_t)5 * (((size_t)ci * gt.n + cj) * gt.l + ck) + T.inc[par][lane][3]) * 36;
  double H[36];
#pragma unroll
  for (int c = 0; c < 36; c++) H[c] = ldg(he + c) * ws;
  const int p = T.inc[par][lane][4];
  double gp[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 4; q++) v = (q == p) ? G.gL[q][c] : v;
    gp[c] = v;
  }
#pragma unroll
  for (int q = 0; q < 4; q++)
#pragma unroll
    for (int b = 0; b < 3; b++) {
      double hb[6];
#pragma unroll
      for (int r = 0; r < 6; r++) hb[r] = bdot<3>(b, G.gL[q], &H[6 * r]);
#pragma unroll
      for (int a = 0; a < 3; a++) out[(a * 4 + q) * 3 + b] = bdot<3>(a, gp, hb);
    }
}

// phase 2: the existing slots of the three rows into stage[a*81 + b*len + pos]
ADFEM_HD void tg_gather_rows(int lane, const GridTet& gt, int par, int mask, const double* tb, double* stage) {
  const TetGridTables& T = *gt.tab;
  const int len = tg_popc(mask);
  for (int idx = lane; idx < 27 * 9; idx += 32) {
    const int s = idx / 9, ab = idx - 9 * s, a = ab / 3, b = ab - 3 * a;
    if (!((mask >> s) & 1)) continue;
    double v = 0.0;
    for (int c = 0; c < T.nsrc[par][s]; c++) v += tb[T.src[par][s][c][0] * TG_BLK + (a * 4 + T.src[par][s][c][1]) * 3 + b];
    stage[a * 81 + b * len + tg_popc(mask & ((1 << s) - 1))] = v;
  }
}

// phase 3: three contiguous runs of 3*len values; rs = rowptr[node]
ADFEM_HD void tg_store_rows(int lane, long long rs, int len, long long nnz, const double* stage, double* vals) {
  for (int a = 0; a < 3; a++) {
    double* out = vals + 3 * ((long long)a * nnz + rs);
    for (int idx = lane; idx < 3 * len; idx += 32) out[idx] = stage[a * 81 + idx];
  }
}

// ---- adjoint ----------------------------------------------------------------------------------------------------------------------
// leading dimension of the per-warp gradient staging: 38 doubles = 19 16-byte units (odd), so that BOTH phases move 16-byte pairs without bank
// conflicts — the lanes' pair stores of phase 1 (lane stride 19 units) and the consecutive-pair loads of phase 2.  The kernel runs at 94 % of the
// L1 data-pipe wavefront peak (profiles/ncu_r02_cfg5_v3.md: 8-byte accesses with a leading dimension of 37 took a quarter of it).
constexpr int TG_ADJ_LD = 38;
constexpr int TG_ADJ_WARP_DOUBLES = 32 * TG_ADJ_LD;
struct alignas(16) TgPair { double x, y; };
constexpr int TG_ADJ_WARPS = 4;

// phase 1: lane -> tetrahedron e0 + lane: st[lane*36 + r*6 + c] = (B dK B^T)_{rc} |det|
ADFEM_HD void tg_tet_adjoint(int lane, const GridTet& gt, long long e0, long long ne, long long nnz, const long long* rowptr, const double* dvals,
                             double* st) {
  const TetGridTables& T = *gt.tab;
  const long long e = e0 + lane;
  if (e >= ne) return;
  const long long cube = e / 5, n1 = gt.n + 1;
  const int t = (int)(e - 5 * cube), ck = (int)(cube % gt.l), cj = (int)((cube / gt.l) % gt.n), ci = (int)(cube / ((long long)gt.l * gt.n));
  const int ev = ((ci + cj + ck + 3) & 1) == 0;
  int vi[4], vj[4], vk[4], mask[4];
  long long rs[4];
  double X[4][3];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int v = T.te[ev][t][q];
    vi[q] = ci + (v & 1); vj[q] = cj + ((v >> 1) & 1); vk[q] = ck + (v >> 2);
    X[q][0] = ldg(gt.xs + vi[q]); X[q][1] = ldg(gt.ys + vj[q]); X[q][2] = ldg(gt.zs + vk[q]);
    mask[q] = tg_row_mask(gt, (vi[q] + vj[q] + vk[q]) & 1, vi[q], vj[q], vk[q]);
    rs[q] = rowptr[((long long)vk[q] * n1 + vj[q]) * n1 + vi[q]];
  }
  Geom<3> G; geom_tet(X, G);
  const double ws = G.wscale < 0 ? -G.wscale : G.wscale;
  // CSR positions of the four vertex columns inside each vertex row, one byte each (rows have at most 19 entries)
  unsigned pos4[4]; int len[4];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    len[p] = tg_popc(mask[p]);
    unsigned pk = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int s = (vk[q] - vk[p] + 1) * 9 + (vj[q] - vj[p] + 1) * 3 + (vi[q] - vi[p] + 1);
      pk |= (unsigned)tg_popc(mask[p] & ((1 << s) - 1)) << (8 * q);
    }
    pos4[p] = pk;
  }
  // grad H = sum over (row component a, vertex p) of b(a, grad lambda_p) (x) t(a, p),  t = sum over (b, q) of dK[(p,a),(q,b)] b(b, grad lambda_q).
  // b(a, .) has three non-zero Voigt rows (device_fem.cuh badd<3>), so component a only touches rows ROW[a]: 18 accumulators live at a
  // time instead of 36 (the kernel ran 12 warps per SM at 168 registers).  Rows 3, 4, 5 receive two components: the second one adds in place.
  constexpr int ROW[3][3] = {{0, 4, 5}, {1, 3, 5}, {2, 3, 4}}, AX[3][3] = {{0, 2, 1}, {1, 2, 0}, {2, 1, 0}};
  double* out = st + lane * TG_ADJ_LD;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    double acc[3][6];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int c = 0; c < 6; c++) acc[i][c] = 0.0;
#pragma unroll
    for (int p = 0; p < 4; p++) {
      const double* row = dvals + 3 * ((long long)a * nnz + rs[p]);
      double tl[6];
#pragma unroll
      for (int c = 0; c < 6; c++) tl[c] = 0.0;
#pragma unroll
      for (int b = 0; b < 3; b++)
#pragma unroll
        for (int q = 0; q < 4; q++) badd<3>(b, G.gL[q], ldg(row + b * len[p] + (int)((pos4[p] >> (8 * q)) & 0xffu)), tl);
#pragma unroll
      for (int i = 0; i < 3; i++) {
        const double bl = G.gL[p][AX[a][i]];
#pragma unroll
        for (int c = 0; c < 6; c++) acc[i][c] += bl * tl[c];
      }
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const int r = ROW[a][i];
      const bool first = (r < 3) || (a == 0) || (a == 1 && r == 3);      // rows 3, 4, 5 are first written by components 1, 0, 0
      TgPair* o2 = reinterpret_cast<TgPair*>(out + 6 * r);          // 16-byte aligned: lane * 38 + 6 r is even
#pragma unroll
      for (int c = 0; c < 3; c++) {
        TgPair v{acc[i][2 * c] * ws, acc[i][2 * c + 1] * ws};
        if (!first) { const TgPair w = o2[c]; v.x += w.x; v.y += w.y; }
        o2[c] = v;
      }
    }
  }
}

// phase 2: grad[(e*g + k)*36 + c] = st[t*36 + c] * w_k over the contiguous run of the warp's tetrahedra
ADFEM_HD void tg_store_grad(int lane, const QuadRule& rule, int g, long long e0, long long ne, const double* st, double* grad) {
  const int nt = (int)(ne - e0 < 32 ? ne - e0 : 32), per = 36 * g;
  double* out = grad + (size_t)e0 * per;
  if (g == 4 && nt == 32 && (reinterpret_cast<size_t>(out) & 15) == 0) {
    // full warp, 4 Gauss points: 16-byte stores; pair q = lane + 32*it covers tetrahedron q / 72, Gauss point (q % 72) / 18, entries 2*(q % 18), +1.
    // 32 * 9 pairs = 4 tetrahedra, so the decomposition of an iteration repeats every 9 iterations with the tetrahedron advanced by 4.
    int off[9]; double w[9];
#pragma unroll
    for (int j = 0; j < 9; j++) {
      const int q = lane + 32 * j, t = q / 72, r = q - 72 * t, k = r / 18, c2 = r - 18 * k;
      off[j] = t * TG_ADJ_LD + 2 * c2; w[j] = rule.w[k];
    }
    TgPair* o2 = reinterpret_cast<TgPair*>(out) + lane;
#pragma unroll 1
    for (int tt = 0; tt < 8; tt++) {
      const double* s4 = st + tt * 4 * TG_ADJ_LD;
#pragma unroll
      for (int j = 0; j < 9; j++) {
        const TgPair v = *reinterpret_cast<const TgPair*>(s4 + off[j]);      // t * 38 + 2 c2: 16-byte aligned
        o2[(tt * 9 + j) * 32] = TgPair{v.x * w[j], v.y * w[j]};
      }
    }
    return;
  }
  for (int idx = lane; idx < nt * per; idx += 32) {
    const int t = idx / per, r = idx - t * per, k = r / 36, c = r - 36 * k;
    out[idx] = st[t * TG_ADJ_LD + c] * rule.w[k];
  }
}

#ifdef __CUDACC__
static __global__ void __launch_bounds__(TG_WARPS * 32) k_tet_grid_elast_fwd(GridTet gt, long long nnz, const long long* __restrict__ rowptr,
                                                                      const double* __restrict__ hbar, double* __restrict__ vals) {
  extern __shared__ __align__(16) double tg_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long node = (long long)blockIdx.x * TG_WARPS + wib, n1 = gt.n + 1;
  if (node >= n1 * n1 * (gt.l + 1)) return;
  const int i = (int)(node % n1), j = (int)((node / n1) % n1), k = (int)(node / (n1 * n1)), par = (i + j + k) & 1;
  double* tb = tg_smem + (size_t)wib * TG_WARP_DOUBLES;
  double* stage = tb + 32 * TG_BLK;
  tg_tet_block(lane, gt, par, i, j, k, hbar, tb);
  const int mask = tg_row_mask(gt, par, i, j, k);
  __syncwarp();
  tg_gather_rows(lane, gt, par, mask, tb, stage);
  __syncwarp();
  tg_store_rows(lane, rowptr[node], tg_popc(mask), nnz, stage, vals);
}

// Work unit = (cube column (ci, cj), chunk c of 32 consecutive tetrahedra of that column): the warp's gradients stay one contiguous run, and units
// are rasterised chunk-slowest / ci-fastest, u = (c*n + cj)*n + ci, so that the warps in flight at any time cover a slab of ~6 cube layers over
// a few rows of columns — a working set of upstream values that fits L2 at any mesh size.  (In element order a wave of warps covers a whole
// (cj, ck) plane at fixed ci, 250 MB of CSR rows for Mesh3(215,215,208): the adjoint fell from 0.39 ns per tetrahedron at 1.3 M to 0.60 ns at 48 M.)
template <int MINB>
static __global__ void __launch_bounds__(TG_ADJ_WARPS * 32, MINB) k_tet_grid_elast_adj(GridTet gt, QuadRule rule, int g, long long ne, long long nnz,
                                                                      const long long* __restrict__ rowptr, const double* __restrict__ dvals,
                                                                      double* __restrict__ grad) {
  extern __shared__ __align__(16) double tg_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long u = (long long)blockIdx.x * TG_ADJ_WARPS + wib, ncol = (long long)gt.n * gt.n;
  const int per_col = 5 * gt.l, nchunk = (per_col + 31) / 32;
  if (u >= ncol * nchunk) return;
  const int c = (int)(u / ncol), cj = (int)((u / gt.n) % gt.n), ci = (int)(u % gt.n);
  const long long col0 = ((long long)ci * gt.n + cj) * per_col, e0 = col0 + 32 * c, e1 = min(col0 + per_col, ne);
  double* st = tg_smem + (size_t)wib * TG_ADJ_WARP_DOUBLES;
  tg_tet_adjoint(lane, gt, e0, e1, nnz, rowptr, dvals, st);
  __syncwarp();
  tg_store_grad(lane, rule, g, e0, e1, st, grad);
}
#endif

}  // namespace adfem
