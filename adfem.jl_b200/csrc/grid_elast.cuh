// Index-free CSR assembly / adjoint of the P1 ELASTICITY operator (ComputeFemStiffnessMatrixMfem, deps/MFEM/ComputeFemStiffnessMatrixMfem/
// ComputeFemStiffnessMatrixMfem.h:4-81) on the reference's structured triangulation `Mesh(m, n, h)` — BASELINE config 3.  Like tri_grid.cuh
// for the scalar operators: connectivity, CSR positions and vertex coordinates are index arithmetic, the only DRAM streams are the 3x3 tangent
// H per Gauss point (216 B per element) and the CSR values (112 B per element); no tile blobs, no halo re-evaluation.
//
// Work decomposition.  A warp owns a strip of 32 node columns and marches over node rows.  Per cell row it forms the Gauss-summed tangents
// Hbar_t = sum_k w_k H_{t,k} of the 2 x 33 triangles under / above its node columns with warp-wide, flat-indexed loads (every byte of the
// contiguous run is used, each H block is read once per strip) and keeps the previous and the current cell row in shared memory.  Lane l then
// gathers the 7 x (2 x 2) entries of node column j0 + l from its six incident triangles in ascending element order (the summation order of
// the general tile kernels), parks them in shared memory in CSR order, and the warp writes the two contiguous runs (component a = 0, 1) of
// the node-row segment with coalesced stores.  Only __syncwarp() separates the three phases.
//
// CSR layout (adfem_assemble_csr, = canonical CSR of the component-blocked matrix): scalar row r of length len starting at rs holds entry
// (a, b, j) at 2*(a*nnz + rs) + b*len + j.
//
// Every phase is a __host__ __device__ function of the LANE index, so tests/host_emul/ can run a warp as three loops over lanes against
// the oracle where there is no GPU.
#pragma once
#include "device_fem.cuh"
#include "gauss_ops.cuh"
#include "grid_index.cuh"

namespace adfem {

constexpr int GE_COLS = 32;                              // node columns (forward) / cell columns (adjoint) per warp
constexpr int GE_TRI = 2 * (GE_COLS + 1);                // triangles of one cell-row segment: cell columns j0-1 .. j0+31
constexpr int GE_HROW = GE_TRI * 9;                      // doubles of Gauss-summed tangents per cell-row segment
constexpr int GE_STAGE = GE_COLS * 14;                   // forward staging per component: <= 14 values per node
constexpr int GE_WARPS = 4;
constexpr int GE_G = 3;                                   // Gauss points per triangle: the structured layout is only detected for the order-2 rule
constexpr int GE_FWD_WARP_DOUBLES = 2 * GE_HROW + 2 * GE_STAGE;
constexpr int GE_NROW = (GE_COLS + 1) * 14;              // adjoint: staged CSR values of one node-row segment (33 nodes) per component
constexpr int GE_ADJ_WARP_DOUBLES = 2 * 2 * GE_NROW + 2 * GE_COLS * 9;

ADFEM_HD int ge_prefix(int j, int m, int A, int B) {    // 32-bit grid_row_prefix: scalar CSR entries of a node row before node column j
  const int jm = j < m ? j : m, j1 = j > 0 ? j - 1 : 0;
  return j * (1 + A + B) + (A + 1) * jm + (1 + B) * j1;
}
ADFEM_HD int ge_min(int a, int b) { return a < b ? a : b; }

// ---- forward ----------------------------------------------------------------------------------------------------------------------
// phase 1: Gauss-summed tangents of cell row ci, cell columns j0-1 .. j0+31, into buf[t*9 + c] (t = 2*(cc - j0 + 1) + triangle); cells
// outside the mesh give zeros
template <int G>
ADFEM_HD void ge_load_cell_row(int lane, const QuadRule& rule, int m, int n, int ci, int j0, const double* coef, double* buf) {
  for (int idx = lane; idx < GE_HROW; idx += 32) {
    const int t = idx / 9, c = idx - 9 * t, cc = j0 - 1 + (t >> 1);
    double s = 0.0;
    if (ci >= 0 && ci < n && cc >= 0 && cc < m) {
      const double* p = coef + ((size_t)(2 * ((size_t)ci * m + cc) + (t & 1)) * G) * 9 + c;
#pragma unroll
      for (int k = 0; k < G; k++) s += ldg(p + 9 * k) * rule.w[k];
    }
    buf[idx] = s;
  }
}

// phase 1, fused constitutive step: the same Gauss-summed tangents straight from the moduli, H_k = plane matrix(E_k, nu_k)
// (gauss_ops.cuh); a lane owns whole triangles, so every plane matrix is evaluated once.  16 B per Gauss point are read instead of 72.
template <int G>
ADFEM_HD void ge_load_cell_row_plane(int lane, const QuadRule& rule, int m, int n, int ci, int j0, int mode, const double* E, const double* nu,
                                     double* buf) {
  for (int t = lane; t < GE_TRI; t += 32) {
    const int cc = j0 - 1 + (t >> 1);
    double s[9];
#pragma unroll
    for (int c = 0; c < 9; c++) s[c] = 0.0;
    if (ci >= 0 && ci < n && cc >= 0 && cc < m) {
      const size_t g0 = (size_t)(2 * ((size_t)ci * m + cc) + (t & 1)) * G;
#pragma unroll
      for (int k = 0; k < G; k++) {
        double H[9];
        plane_matrix_body(mode, ldg(E + g0 + k), ldg(nu + g0 + k), H);
#pragma unroll
        for (int c = 0; c < 9; c++) s[c] += H[c] * rule.w[k];
      }
    }
#pragma unroll
    for (int c = 0; c < 9; c++) buf[t * 9 + c] = s[c];
  }
}

// contribution of one triangle (vertices v0 v1 v2, Gauss-summed tangent Hb) to the row of its local node P: the 2x2 blocks against its three
// nodes go to the pattern slots S0 S1 S2 of acc[slot][a*2 + b]
template <int P, int S0, int S1, int S2>
ADFEM_HD void ge_add_triangle(const double* Hb, double2 v0, double2 v1, double2 v2, int heron, double acc[7][4]) {
  Geom<2> G; geom_tri(v0, v1, v2, heron, G);
  double H[9];
#pragma unroll
  for (int c = 0; c < 9; c++) H[c] = Hb[c] * G.wscale;
  constexpr int SL[3] = {S0, S1, S2};
#pragma unroll
  for (int q = 0; q < 3; q++)
#pragma unroll
    for (int b = 0; b < 2; b++) {
      double hb[3];
#pragma unroll
      for (int r = 0; r < 3; r++) hb[r] = bdot<2>(b, G.gL[q], &H[3 * r]);
#pragma unroll
      for (int a = 0; a < 2; a++) acc[SL[q]][a * 2 + b] += bdot<2>(a, G.gL[P], hb);
    }
}

// phase 2: node (i, j0 + lane).  Pattern slots: 0 (i-1,j)  1 (i-1,j+1)  2 (i,j-1)  3 (i,j)  4 (i,j+1)  5 (i+1,j-1)  6 (i+1,j).
// P / C = tangents of cell rows i-1 / i (ge_load_cell_row).  Writes stage[a*GE_STAGE + 2*o + b*len + pos].
// MAPPED (structured connectivity on arbitrary node positions): the positions of the seven stencil nodes come from the coordinate array `xy`
// ([node][2]) instead of the axis tables.
template <bool MAPPED = false>
ADFEM_HD void ge_node(int lane, int heron, const GridTri& gt, int i, int j0, const double* P, const double* C, double* stage, const double* xy = nullptr) {
  const int m = gt.m, n = gt.n, jn = j0 + lane;
  if (jn > m) return;
  const int A = i > 0, B = i < n, jl = jn > 0, jr = jn < m;
  double acc[7][4];
#pragma unroll
  for (int s = 0; s < 7; s++)
#pragma unroll
    for (int ab = 0; ab < 4; ab++) acc[s][ab] = 0.0;
  if constexpr (MAPPED) {
    auto Pt = [&](int di, int dj) { const double* p = xy + 2 * ((size_t)(i + di) * (m + 1) + (jn + dj)); return make_double2(ldg(p), ldg(p + 1)); };
    const double2 pc = Pt(0, 0);
    if (A) {                                                 // cell row i-1
      const double2 pd = Pt(-1, 0);
      if (jl) ge_add_triangle<2, 2, 0, 3>(P + (2 * lane + 1) * 9, Pt(0, -1), pd, pc, heron, acc);                                                     // T1(i-1, j-1) = [TL BR TR], node = TR
      if (jr) {
        const double2 pdr = Pt(-1, 1);
        ge_add_triangle<2, 0, 1, 3>(P + (2 * lane + 2) * 9, pd, pdr, pc, heron, acc);                                                                   // T0(i-1, j)   = [BL BR TL], node = TL
        ge_add_triangle<0, 3, 1, 4>(P + (2 * lane + 3) * 9, pc, pdr, Pt(0, 1), heron, acc);                                                             // T1(i-1, j)   = [TL BR TR], node = TL
      }
    }
    if (B) {                                                 // cell row i
      const double2 pu = Pt(1, 0);
      if (jl) {
        const double2 pl = Pt(0, -1), pul = Pt(1, -1);
        ge_add_triangle<1, 2, 3, 5>(C + (2 * lane) * 9, pl, pc, pul, heron, acc);                                                                       // T0(i, j-1)   = [BL BR TL], node = BR
        ge_add_triangle<1, 5, 3, 6>(C + (2 * lane + 1) * 9, pul, pc, pu, heron, acc);                                                                   // T1(i, j-1)   = [TL BR TR], node = BR
      }
      if (jr) ge_add_triangle<0, 3, 4, 6>(C + (2 * lane + 2) * 9, pc, Pt(0, 1), pu, heron, acc);                                                       // T0(i, j)     = [BL BR TL], node = BL
    }
  } else {
  const double xc = ldg(gt.xs + jn), xl = jl ? ldg(gt.xs + jn - 1) : 0.0, xr = jr ? ldg(gt.xs + jn + 1) : 0.0;
  const double yc = ldg(gt.ys + i);
  if (A) {                                                   // cell row i-1: this node is on its top side
    const double yd = ldg(gt.ys + i - 1);
    if (jl) ge_add_triangle<2, 2, 0, 3>(P + (2 * lane + 1) * 9, make_double2(xl, yc), make_double2(xc, yd), make_double2(xc, yc), heron, acc);         // T1(i-1, j-1) = [TL BR TR], node = TR
    if (jr) {
      ge_add_triangle<2, 0, 1, 3>(P + (2 * lane + 2) * 9, make_double2(xc, yd), make_double2(xr, yd), make_double2(xc, yc), heron, acc);               // T0(i-1, j)   = [BL BR TL], node = TL
      ge_add_triangle<0, 3, 1, 4>(P + (2 * lane + 3) * 9, make_double2(xc, yc), make_double2(xr, yd), make_double2(xr, yc), heron, acc);               // T1(i-1, j)   = [TL BR TR], node = TL
    }
  }
  if (B) {                                                   // cell row i: this node is on its bottom side
    const double yu = ldg(gt.ys + i + 1);
    if (jl) {
      ge_add_triangle<1, 2, 3, 5>(C + (2 * lane) * 9, make_double2(xl, yc), make_double2(xc, yc), make_double2(xl, yu), heron, acc);                   // T0(i, j-1)   = [BL BR TL], node = BR
      ge_add_triangle<1, 5, 3, 6>(C + (2 * lane + 1) * 9, make_double2(xl, yu), make_double2(xc, yc), make_double2(xc, yu), heron, acc);               // T1(i, j-1)   = [TL BR TR], node = BR
    }
    if (jr) ge_add_triangle<0, 3, 4, 6>(C + (2 * lane + 2) * 9, make_double2(xc, yc), make_double2(xr, yc), make_double2(xc, yu), heron, acc);         // T0(i, j)     = [BL BR TL], node = BL
  }
  }
  const int ex[7] = {A, A && jr, jl, 1, jr, B && jl, B};
  const int len = ex[0] + ex[1] + ex[2] + 1 + ex[4] + ex[5] + ex[6];
  const int o = ge_prefix(jn, m, A, B) - ge_prefix(j0, m, A, B);
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int b = 0; b < 2; b++) {
      double* dst = stage + a * GE_STAGE + 2 * o + b * len;
      int pos = 0;
#pragma unroll
      for (int s = 0; s < 7; s++)
        if (ex[s]) dst[pos++] = acc[s][a * 2 + b];
    }
}

// phase 3: the two contiguous runs of the node-row segment; rowbase = CSR offset of node (i, 0)
ADFEM_HD void ge_store_node_row(int lane, int m, int n, int i, int j0, long long rowbase, long long nnz, const double* stage, double* vals) {
  const int A = i > 0, B = i < n, pbase = ge_prefix(j0, m, A, B);
  const int total = 2 * (ge_prefix(ge_min(j0 + GE_COLS, m + 1), m, A, B) - pbase);
  for (int a = 0; a < 2; a++) {
    double* out = vals + 2 * ((long long)a * nnz + rowbase + pbase);
    for (int idx = lane; idx < total; idx += 32) out[idx] = stage[a * GE_STAGE + idx];
  }
}

// ---- adjoint ----------------------------------------------------------------------------------------------------------------------
// phase 1: upstream CSR values of node row i, node columns c0 .. c0+32, both components, into buf[a*GE_NROW + idx] (contiguous runs)
ADFEM_HD void ge_load_node_row(int lane, int m, int n, int i, int c0, long long rowbase, long long nnz, const double* dvals, double* buf) {
  if (i < 0 || i > n) return;
  const int A = i > 0, B = i < n, pbase = ge_prefix(c0, m, A, B);
  const int total = 2 * (ge_prefix(ge_min(c0 + GE_COLS + 1, m + 1), m, A, B) - pbase);
  for (int a = 0; a < 2; a++) {
    const double* in = dvals + 2 * ((long long)a * nnz + rowbase + pbase);
    for (int idx = lane; idx < total; idx += 32) buf[a * GE_NROW + idx] = ldg(in + idx);
  }
}

// upstream value of entry (row node (i, j) component a, column node (i + di, j + dj) component b) from the staged node row `buf` of row i
ADFEM_HD double ge_entry(const double* buf, int m, int n, int c0, int i, int j, int di, int dj, int a, int b) {
  const int A = i > 0, B = i < n, jl = j > 0, jr = j < m;
  const int ex[7] = {A, A && jr, jl, 1, jr, B && jl, B};
  const int s = di < 0 ? (dj == 0 ? 0 : 1) : (di == 0 ? 3 + dj : (dj < 0 ? 5 : 6));       // pattern slot of the offset
  int pos = 0, len = 0;
#pragma unroll
  for (int t = 0; t < 7; t++) { pos += (t < s) ? ex[t] : 0; len += ex[t]; }
  const int o = ge_prefix(j, m, A, B) - ge_prefix(c0, m, A, B);
  return buf[a * GE_NROW + 2 * o + b * len + pos];
}

// gradient block B dK B^T wscale of one triangle with nodes (iv[q], jv[q]); lo / hi = staged node rows ci / ci + 1
ADFEM_HD void ge_triangle_adjoint(const double* lo, const double* hi, int m, int n, int c0, int ci, const int iv[3], const int jv[3], double2 v0,
                                  double2 v1, double2 v2, int heron, double* out9) {
  Geom<2> G; geom_tri(v0, v1, v2, heron, G);
  double gH[9];
#pragma unroll
  for (int c = 0; c < 9; c++) gH[c] = 0.0;
#pragma unroll
  for (int a = 0; a < 2; a++)
#pragma unroll
    for (int p = 0; p < 3; p++) {
      const double* rowbuf = iv[p] == ci ? lo : hi;
      double tl[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int b = 0; b < 2; b++)
#pragma unroll
        for (int q = 0; q < 3; q++) badd<2>(b, G.gL[q], ge_entry(rowbuf, m, n, c0, iv[p], jv[p], iv[q] - iv[p], jv[q] - jv[p], a, b), tl);
      double bl[3] = {0.0, 0.0, 0.0};
      badd<2>(a, G.gL[p], 1.0, bl);
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) gH[3 * r + c] += bl[r] * tl[c];
    }
#pragma unroll
  for (int c = 0; c < 9; c++) out9[c] = gH[c] * G.wscale;
}

// phase 2: both triangles of cell (ci, c0 + lane) -> gst[(2*lane + t)*9 + c]
template <bool MAPPED = false>
ADFEM_HD void ge_cell_adjoint(int lane, int heron, const GridTri& gt, int ci, int c0, const double* lo, const double* hi, double* gst, const double* xy = nullptr) {
  const int m = gt.m, n = gt.n, cc = c0 + lane;
  if (cc >= m) return;
  double2 BL, BR, TL, TR;
  if constexpr (MAPPED) {
    const double* p0 = xy + 2 * ((size_t)ci * (m + 1) + cc);
    const double* p1 = xy + 2 * ((size_t)(ci + 1) * (m + 1) + cc);
    BL = make_double2(ldg(p0), ldg(p0 + 1)); BR = make_double2(ldg(p0 + 2), ldg(p0 + 3));
    TL = make_double2(ldg(p1), ldg(p1 + 1)); TR = make_double2(ldg(p1 + 2), ldg(p1 + 3));
  } else {
    const double x0 = ldg(gt.xs + cc), x1 = ldg(gt.xs + cc + 1), y0 = ldg(gt.ys + ci), y1 = ldg(gt.ys + ci + 1);
    BL = make_double2(x0, y0); BR = make_double2(x1, y0); TL = make_double2(x0, y1); TR = make_double2(x1, y1);
  }
  {                                                          // T0 = [BL BR TL]
    const int iv[3] = {ci, ci, ci + 1}, jv[3] = {cc, cc + 1, cc};
    ge_triangle_adjoint(lo, hi, m, n, c0, ci, iv, jv, BL, BR, TL, heron, gst + (2 * lane) * 9);
  }
  {                                                          // T1 = [TL BR TR]
    const int iv[3] = {ci + 1, ci, ci + 1}, jv[3] = {cc, cc + 1, cc + 1};
    ge_triangle_adjoint(lo, hi, m, n, c0, ci, iv, jv, TL, BR, TR, heron, gst + (2 * lane + 1) * 9);
  }
}

// phase 3: grad[(e*g + k)*9 + c] = gst[t*9 + c] * w_k over the contiguous run of the segment's elements (e = 2*(ci*m + c0) + t)
template <int G>
ADFEM_HD void ge_store_cell_row(int lane, const QuadRule& rule, int m, int ci, int c0, const double* gst, double* grad) {
  constexpr int per = 9 * G;
  const int ncell = ge_min(GE_COLS, m - c0), total = 2 * ncell * per;
  double* out = grad + (size_t)(2 * ((size_t)ci * m + c0)) * per;
  for (int idx = lane; idx < total; idx += 32) {
    const int t = idx / per, r = idx - t * per, k = r / 9, c = r - 9 * k;
    out[idx] = gst[t * 9 + c] * rule.w[k];
  }
}

// phase 3, fused constitutive step: (dE, dnu)[e*g + k] = d plane matrix / d(E, nu)^T (w_k gst[t]) over the contiguous run of the segment's Gauss points
template <int G>
ADFEM_HD void ge_store_cell_row_plane(int lane, const QuadRule& rule, int m, int ci, int c0, int mode, const double* E, const double* nu,
                                      const double* gst, double* grad_E, double* grad_nu) {
  const int ncell = ge_min(GE_COLS, m - c0), total = 2 * ncell * G;
  const size_t g0 = (size_t)(2 * ((size_t)ci * m + c0)) * G;
  for (int idx = lane; idx < total; idx += 32) {
    const int t = idx / G, k = idx - t * G;
    double gk[9];
#pragma unroll
    for (int c = 0; c < 9; c++) gk[c] = gst[t * 9 + c] * rule.w[k];
    plane_matrix_grad_body(mode, ldg(E + g0 + idx), ldg(nu + g0 + idx), gk, grad_E + g0 + idx, grad_nu + g0 + idx);
  }
}

#ifdef __CUDACC__
// L2 prefetch of a contiguous run of doubles [p, p + count), one 128-byte line per lane and step (no effect on results)
__device__ __forceinline__ void ge_prefetch_run(int lane, const double* p, long long count) {
  for (long long o = 16LL * lane; o < count; o += 16 * 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
}

// Forward kernel: node rows [0, n], strips of 32 node columns; a warp handles `rows_per_warp` consecutive node rows of one strip.
// PLANE: the tangents come from the moduli (coef = E, coef2 = nu, mode = 0 | 1) instead of from H (coef).
template <bool PLANE, bool MAPPED = false>
__global__ void __launch_bounds__(GE_WARPS * 32, 3) k_grid_elast_fwd(DevMesh dm, GridTri gt, long long nnz, int rows_per_warp, int mode,
                                                                  const double* __restrict__ coef, const double* __restrict__ coef2, double* __restrict__ vals) {
  extern __shared__ __align__(16) double ge_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, m = gt.m, n = gt.n;
  const int strips = (m + 1 + GE_COLS - 1) / GE_COLS, chunks = (n + 1 + rows_per_warp - 1) / rows_per_warp;
  const long long gw = (long long)blockIdx.x * GE_WARPS + wib;
  if (gw >= (long long)strips * chunks) return;
  const int strip = (int)(gw % strips), chunk = (int)(gw / strips), j0 = strip * GE_COLS;
  const int i0 = chunk * rows_per_warp, i1 = ge_min(i0 + rows_per_warp, n + 1);
  double* P = ge_smem + (size_t)wib * GE_FWD_WARP_DOUBLES;
  double* C = P + GE_HROW;
  double* stage = C + GE_HROW;
  auto load = [&](int ci, double* buf) {
    if constexpr (PLANE) ge_load_cell_row_plane<GE_G>(lane, dm.rule, m, n, ci, j0, mode, coef, coef2, buf);
    else ge_load_cell_row<GE_G>(lane, dm.rule, m, n, ci, j0, coef, buf);
  };
  load(i0 - 1, P);
  long long rowbase = grid_rowptr(i0, 0, m, n);
  const int pc0 = j0 > 0 ? j0 - 1 : 0, pc1 = ge_min(j0 + GE_COLS, m);            // cell columns of the strip that exist
  const int per_cell = PLANE ? 2 * GE_G : 2 * GE_G * 9;                          // doubles per cell in coef (and coef2)
  for (int i = i0; i < i1; i++) {
    load(i, C);
    if (i + 2 < n && i + 2 <= i1 && pc1 > pc0) {                                 // the run cell row i+2 will read: in L2 by the time it is needed
      const size_t off = ((size_t)(i + 2) * m + pc0) * per_cell;
      ge_prefetch_run(lane, coef + off, (long long)(pc1 - pc0) * per_cell);
      if (PLANE) ge_prefetch_run(lane, coef2 + off, (long long)(pc1 - pc0) * per_cell);
    }
    __syncwarp();
    ge_node<MAPPED>(lane, dm.heron, gt, i, j0, P, C, stage, dm.coords);
    __syncwarp();
    ge_store_node_row(lane, m, n, i, j0, rowbase, nnz, stage, vals);
    __syncwarp();
    rowbase += ge_prefix(m + 1, m, i > 0, i < n);
    double* t = P; P = C; C = t;
  }
}

// Adjoint kernel: cell rows [0, n), strips of 32 cell columns; a warp handles `rows_per_warp` consecutive cell rows of one strip.
// PLANE: gradients with respect to the moduli (E, nu in; grad = dE, grad2 = dnu) instead of with respect to H.
template <bool PLANE, bool MAPPED = false>
__global__ void __launch_bounds__(GE_WARPS * 32, 2) k_grid_elast_adj(DevMesh dm, GridTri gt, long long nnz, int rows_per_warp, int mode,
                                                                  const double* __restrict__ E, const double* __restrict__ nu,
                                                                  const double* __restrict__ dvals, double* __restrict__ grad, double* __restrict__ grad2) {
  extern __shared__ __align__(16) double ge_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, m = gt.m, n = gt.n;
  const int strips = (m + GE_COLS - 1) / GE_COLS, chunks = (n + rows_per_warp - 1) / rows_per_warp;
  const long long gw = (long long)blockIdx.x * GE_WARPS + wib;
  if (gw >= (long long)strips * chunks) return;
  const int strip = (int)(gw % strips), chunk = (int)(gw / strips), c0 = strip * GE_COLS;
  const int r0 = chunk * rows_per_warp, r1 = ge_min(r0 + rows_per_warp, n);
  double* lo = ge_smem + (size_t)wib * GE_ADJ_WARP_DOUBLES;
  double* hi = lo + 2 * GE_NROW;
  double* gst = hi + 2 * GE_NROW;
  long long rowbase = grid_rowptr(r0, 0, m, n);
  ge_load_node_row(lane, m, n, r0, c0, rowbase, nnz, dvals, lo);
  for (int ci = r0; ci < r1; ci++) {
    rowbase += ge_prefix(m + 1, m, ci > 0, ci < n);          // CSR offset of node row ci + 1
    ge_load_node_row(lane, m, n, ci + 1, c0, rowbase, nnz, dvals, hi);
    if (ci + 2 <= n && ci + 2 <= r1) {                       // the two runs node row ci+2 will read (it has a row below: A = 1)
      const int B2 = ci + 2 < n, pb = ge_prefix(c0, m, 1, B2);
      const long long cnt = 2LL * (ge_prefix(ge_min(c0 + GE_COLS + 1, m + 1), m, 1, B2) - pb);
      const long long rb2 = rowbase + ge_prefix(m + 1, m, 1, ci + 1 < n);      // CSR offset of node row ci + 2
      ge_prefetch_run(lane, dvals + 2 * (rb2 + pb), cnt);
      ge_prefetch_run(lane, dvals + 2 * (nnz + rb2 + pb), cnt);
    }
    __syncwarp();
    ge_cell_adjoint<MAPPED>(lane, dm.heron, gt, ci, c0, lo, hi, gst, dm.coords);
    __syncwarp();
    if constexpr (PLANE) ge_store_cell_row_plane<GE_G>(lane, dm.rule, m, ci, c0, mode, E, nu, gst, grad, grad2);
    else ge_store_cell_row<GE_G>(lane, dm.rule, m, ci, c0, gst, grad);
    __syncwarp();
    double* t = lo; lo = hi; hi = t;
  }
}
#endif

}  // namespace adfem
