// CSR assembly / adjoint kernels for the reference's STRUCTURED triangulation `Mesh(m, n, h)` version 1
// (src/MFEM/MFEM.jl:134-146): (m+1) x (n+1) nodes numbered row-major (node (i,j) = i*(m+1)+j), cell (ci,cj) split by its
// anti-diagonal into elements 2*(ci*m+cj) = T0 = [BL, BR, TL] and 2*(ci*m+cj)+1 = T1 = [TL, BR, TR] (T1 after MFEM's
// orientation fix).  The mesh handle detects this layout from the (coords, elems) arrays it is given (any rectilinear node
// coordinates x = xs[j], y = ys[i]) and then NO mesh-static index data is read at all: connectivity, CSR positions and
// vertex coordinates are index arithmetic, so the only DRAM streams are the coefficients in and the values out
// (kappa 24 B + values 28 B per element instead of the general path's 72 B + tile blobs).
//
// Work decomposition: a warp owns a strip of 31 node columns and marches over node rows.  Lane l holds the local matrices
// of cell column j0-1+l for the previous and the current cell row in REGISTERS; the contributions of the cell to the left
// come over warp shuffles.  No shared-memory exchange and no CTA-wide barrier: shared memory is used only as a per-warp
// transpose buffer so that the 7 entries of 31 consecutive CSR rows leave (forward) or arrive (adjoint) as fully
// coalesced 256-byte accesses.  Per-entry summation order is ascending element id, i.e. the order of the general path.
#pragma once
#include "kernels.cuh"

namespace adfem {

struct GridTri {
  int m, n;                    // cells in x and y
  const double* xs;            // m+1 node abscissae
  const double* ys;            // n+1 node ordinates
};

constexpr int GRID_STRIP = 31;   // node columns per warp (lane 0 / lane 31 carry the neighbouring strip's cell / node column)
constexpr int GRID_WARPS = 8;

// CSR row pointer of node (i, j), j in [0, m+1] (j = m+1: end of node row i), closed form for the 7-point pattern
//   row = [ (i-1,j), (i-1,j+1), (i,j-1), (i,j), (i,j+1), (i+1,j-1), (i+1,j) ]  restricted to existing nodes
__host__ __device__ __forceinline__ long long grid_row_prefix(int j, int m, int A, int B) {
  const int jm = j < m ? j : m, j1 = j > 0 ? j - 1 : 0;
  return (long long)j * (1 + A + B) + (long long)(A + 1) * jm + (long long)(1 + B) * j1;
}
__host__ __device__ __forceinline__ long long grid_rowptr(int i, int j, int m, int n) {
  const int A = i > 0, B = i < n;
  long long before = 0;
  if (i > 0) {
    before = grid_row_prefix(m + 1, m, 0, n > 0);                                   // node row 0
    if (i > 1) before += (long long)(i - 1) * grid_row_prefix(m + 1, m, 1, 1);      // node rows 1 .. i-1 (all have a row above and below)
  }
  return before + grid_row_prefix(j, m, A, B);
}

__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_down1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// packed symmetric index of (p,q), p <= q, 3x3: 00 01 02 11 12 22
#define S00 0
#define S01 1
#define S02 2
#define S11 3
#define S12 4
#define S22 5

// local matrices (packed upper triangle) of both triangles of cell (ci, cj); zero outside the grid
template <int OP>
__device__ __forceinline__ void grid_cell_matrices(const DevMesh& m, const GridTri& gt, int ci, int cj, double x0, double x1, const double kap[6],
                                                   double T0[6], double T1[6]) {
#pragma unroll
  for (int s = 0; s < 6; s++) { T0[s] = 0.0; T1[s] = 0.0; }
  if (ci < 0 || ci >= gt.n || cj < 0 || cj >= gt.m) return;
  const double y0 = __ldg(gt.ys + ci), y1 = __ldg(gt.ys + ci + 1);
  const double2 BL = make_double2(x0, y0), BR = make_double2(x1, y0), TL = make_double2(x0, y1), TR = make_double2(x1, y1);
  Geom<2> G;
  geom_tri(BL, BR, TL, m.heron, G);
  local_matrix_scalar<2, 1, OP, 3>(m, G, [&](int k) { return kap[k]; }, [&](int s, double v) { T0[s] = v; });
  geom_tri(TL, BR, TR, m.heron, G);
  local_matrix_scalar<2, 1, OP, 3>(m, G, [&](int k) { return kap[3 + k]; }, [&](int s, double v) { T1[s] = v; });
}

// the 6 coefficients (2 triangles x 3 Gauss points) of cell (ci, cj): 48 contiguous, 16-byte aligned bytes
__device__ __forceinline__ void grid_load_coef(const GridTri& gt, const double* __restrict__ coef, int ci, int cj, double kap[6]) {
  if (ci < 0 || ci >= gt.n || cj < 0 || cj >= gt.m) {
#pragma unroll
    for (int k = 0; k < 6; k++) kap[k] = 0.0;
    return;
  }
  const double2* p = reinterpret_cast<const double2*>(coef + 6 * ((size_t)ci * gt.m + cj));
  const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  kap[0] = a.x; kap[1] = a.y; kap[2] = b.x; kap[3] = b.y; kap[4] = c.x; kap[5] = c.y;
}

// Forward: vals[nnz] of the scalar operator OP on the structured triangulation (3 Gauss points per element).
// grid: ceil(strips * chunks / GRID_WARPS) CTAs of GRID_WARPS warps; `rows_per_warp` node rows per warp.
template <int OP>
__global__ void __launch_bounds__(GRID_WARPS * 32) k_grid_fwd(DevMesh m, GridTri gt, int rows_per_warp, const double* __restrict__ coef,
                                                              double* __restrict__ vals) {
  __shared__ double stage_all[GRID_WARPS][GRID_STRIP * 7 + 1];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int strips = (gt.m + 1 + GRID_STRIP - 1) / GRID_STRIP, chunks = (gt.n + 1 + rows_per_warp - 1) / rows_per_warp;
  const long long gw = (long long)blockIdx.x * GRID_WARPS + wib;
  if (gw >= (long long)strips * chunks) return;
  const int strip = (int)(gw % strips), chunk = (int)(gw / strips);
  const int j0 = strip * GRID_STRIP, cj = j0 - 1 + lane;            // this lane's cell column = its node column
  const int i0 = chunk * rows_per_warp, i1 = min(i0 + rows_per_warp, gt.n + 1);
  double* stage = stage_all[wib];
  const double x0 = (cj >= 0 && cj <= gt.m) ? __ldg(gt.xs + cj) : 0.0, x1 = (cj >= 0 && cj < gt.m) ? __ldg(gt.xs + cj + 1) : 0.0;
  const bool has_node = lane >= 1 && cj <= gt.m;                     // lanes 1..31 write node column cj
  const int jend = min(j0 + GRID_STRIP, gt.m + 1);                   // one past the last node column of the strip

  double kap[6], kn1[6], pT0[6], pT1[6], cT0[6], cT1[6];
  grid_load_coef(gt, coef, i0 - 1, cj, kap);
  grid_cell_matrices<OP>(m, gt, i0 - 1, cj, x0, x1, kap, pT0, pT1);
  grid_load_coef(gt, coef, i0, cj, kap);
  grid_load_coef(gt, coef, i0 + 1 < i1 ? i0 + 1 : -1, cj, kn1);
  for (int i = i0; i < i1; i++) {
    double kn2[6];
    grid_load_coef(gt, coef, i + 2 < i1 ? i + 2 : -1, cj, kn2);    // two cell rows ahead: in flight while rows i, i+1 are computed
    grid_cell_matrices<OP>(m, gt, i, cj, x0, x1, kap, cT0, cT1);
    // contributions of the cell column to the left (lane - 1)
    const double l_pT1_21 = shfl_up1(pT1[S12]), l_pT1_20 = shfl_up1(pT1[S02]), l_pT1_22 = shfl_up1(pT1[S22]);
    const double l_cT0_10 = shfl_up1(cT0[S01]), l_cT0_11 = shfl_up1(cT0[S11]), l_cT1_11 = shfl_up1(cT1[S11]);
    const double l_cT0_12 = shfl_up1(cT0[S12]), l_cT1_10 = shfl_up1(cT1[S01]), l_cT1_12 = shfl_up1(cT1[S12]);
    // the 7 entries of node (i, cj), each summed in ascending element order
    const double vD = l_pT1_21 + pT0[S02];
    const double vDR = pT0[S12] + pT1[S01];
    const double vL = l_pT1_20 + l_cT0_10;
    const double vC = ((((l_pT1_22 + pT0[S22]) + pT1[S00]) + l_cT0_11) + l_cT1_11) + cT0[S00];
    const double vR = pT1[S02] + cT0[S01];
    const double vUL = l_cT0_12 + l_cT1_10;
    const double vU = l_cT1_12 + cT0[S02];
    const int A = i > 0, B = i < gt.n;
    const long long pbase = grid_row_prefix(j0, gt.m, A, B);
    if (has_node) {
      int o = (int)(grid_row_prefix(cj, gt.m, A, B) - pbase);
      const bool jl = cj > 0, jr = cj < gt.m;
      if (A) stage[o++] = vD;
      if (A && jr) stage[o++] = vDR;
      if (jl) stage[o++] = vL;
      stage[o++] = vC;
      if (jr) stage[o++] = vR;
      if (B && jl) stage[o++] = vUL;
      if (B) stage[o++] = vU;
    }
    __syncwarp();
    const int total = (int)(grid_row_prefix(jend, gt.m, A, B) - pbase);
    double* out = vals + (grid_rowptr(i, 0, gt.m, gt.n) + pbase);
    for (int t = lane; t < total; t += 32) out[t] = stage[t];
    __syncwarp();
#pragma unroll
    for (int s = 0; s < 6; s++) { pT0[s] = cT0[s]; pT1[s] = cT1[s]; kap[s] = kn1[s]; kn1[s] = kn2[s]; }
  }
}

// Adjoint: grad_coef[e*3 + k] from upstream dvals[nnz].  Lane l owns node column j0+l (32 columns, the last one only feeds
// its left neighbour) and cell column j0+l for l < 31.
template <int OP>
__global__ void __launch_bounds__(GRID_WARPS * 32) k_grid_adj(DevMesh m, GridTri gt, int rows_per_warp, const double* __restrict__ dvals,
                                                              double* __restrict__ grad_coef) {
  __shared__ double stage_all[GRID_WARPS][32 * 7];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int strips = (gt.m + GRID_STRIP - 1) / GRID_STRIP, chunks = (gt.n + rows_per_warp - 1) / rows_per_warp;   // over CELLS
  const long long gw = (long long)blockIdx.x * GRID_WARPS + wib;
  if (gw >= (long long)strips * chunks) return;
  const int strip = (int)(gw % strips), chunk = (int)(gw / strips);
  const int j0 = strip * GRID_STRIP, j = j0 + lane;                  // node column == cell column of this lane
  const int c0 = chunk * rows_per_warp, c1 = min(c0 + rows_per_warp, gt.n);
  double* stage = stage_all[wib];
  const bool node_ok = j <= gt.m, cell_ok = lane < GRID_STRIP && j < gt.m;
  const double x0 = node_ok ? __ldg(gt.xs + j) : 0.0, x1 = (j < gt.m) ? __ldg(gt.xs + j + 1) : 0.0;
  const int jend = min(j0 + 32, gt.m + 1);

  // node row i of the strip: its CSR entries are one contiguous run; every lane fetches elements lane, lane+32, ... (coalesced)
  auto fetch = [&](int i, double raw[7]) {
#pragma unroll
    for (int k = 0; k < 7; k++) raw[k] = 0.0;
    if (i > gt.n) return;
    const int A = i > 0, B = i < gt.n;
    const long long pbase = grid_row_prefix(j0, gt.m, A, B);
    const int total = (int)(grid_row_prefix(jend, gt.m, A, B) - pbase);
    const double* in = dvals + (grid_rowptr(i, 0, gt.m, gt.n) + pbase);
#pragma unroll
    for (int k = 0; k < 7; k++) if (lane + 32 * k < total) raw[k] = __ldg(in + lane + 32 * k);
  };
  // transpose through shared memory: the 7 logical entries [D, DR, L, C, R, UL, U] of node (i, j), missing ones = 0
  auto unpack = [&](int i, const double raw[7], double r[7]) {
    const int A = i > 0, B = i < gt.n;
    const long long pbase = grid_row_prefix(j0, gt.m, A, B);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 7; k++) stage[lane + 32 * k] = raw[k];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 7; k++) r[k] = 0.0;
    if (node_ok && i <= gt.n) {
      int o = (int)(grid_row_prefix(j, gt.m, A, B) - pbase);
      const bool jl = j > 0, jr = j < gt.m;
      if (A) r[0] = stage[o++];
      if (A && jr) r[1] = stage[o++];
      if (jl) r[2] = stage[o++];
      r[3] = stage[o++];
      if (jr) r[4] = stage[o++];
      if (B && jl) r[5] = stage[o++];
      if (B) r[6] = stage[o++];
    }
  };
  double lo[7], hi[7], ra[7], rb[7];                                 // node rows ci, ci+1 (unpacked) and ci+1, ci+2 (raw, in flight)
  fetch(c0, ra);
  fetch(c0 + 1, rb);
  unpack(c0, ra, lo);
#pragma unroll
  for (int k = 0; k < 7; k++) ra[k] = rb[k];
  fetch(c0 + 2, rb);
  for (int ci = c0; ci < c1; ci++) {
    unpack(ci + 1, ra, hi);
#pragma unroll
    for (int k = 0; k < 7; k++) ra[k] = rb[k];
    fetch(ci + 3, rb);                                               // two node rows ahead
    // rows of the right-hand node column (lane + 1): BR = (ci, j+1), TR = (ci+1, j+1)
    const double br_L = shfl_down1(lo[2]), br_C = shfl_down1(lo[3]), br_UL = shfl_down1(lo[5]), br_U = shfl_down1(lo[6]);
    const double tr_D = shfl_down1(hi[0]), tr_L = shfl_down1(hi[2]), tr_C = shfl_down1(hi[3]);
    double gk[6];
#pragma unroll
    for (int s = 0; s < 6; s++) gk[s] = 0.0;
    if (cell_ok) {
      const double y0 = __ldg(gt.ys + ci), y1 = __ldg(gt.ys + ci + 1);
      const double2 BL = make_double2(x0, y0), BR = make_double2(x1, y0), TL = make_double2(x0, y1), TR = make_double2(x1, y1);
      Geom<2> G;
      // T0 = [BL, BR, TL]: dK(p,q) = entry (row of local p, column of local q)
      geom_tri(BL, BR, TL, m.heron, G);
      const double t0[9] = {lo[3], lo[4], lo[6], br_L, br_C, br_UL, hi[0], hi[1], hi[3]};
      local_adjoint_scalar<2, 1, OP, 3>(m, G, [&](int p, int q) { return t0[p * 3 + q]; }, [&](int k, double v) { gk[k] = v; });
      // T1 = [TL, BR, TR]
      geom_tri(TL, BR, TR, m.heron, G);
      const double t1[9] = {hi[3], hi[1], hi[4], br_UL, br_C, br_U, tr_L, tr_D, tr_C};
      local_adjoint_scalar<2, 1, OP, 3>(m, G, [&](int p, int q) { return t1[p * 3 + q]; }, [&](int k, double v) { gk[3 + k] = v; });
    }
    // the 6 gradients of every cell of the strip leave through the transpose buffer as one contiguous, coalesced run
    __syncwarp();
    if (lane < GRID_STRIP) {
#pragma unroll
      for (int s = 0; s < 6; s++) stage[lane * 7 + s] = gk[s];
    }
    __syncwarp();
    const int ncell = min(GRID_STRIP, gt.m - j0);
    double* out = grad_coef + 6 * ((size_t)ci * gt.m + j0);
    for (int t = lane; t < 6 * ncell; t += 32) out[t] = stage[(t / 6) * 7 + t % 6];
#pragma unroll
    for (int k = 0; k < 7; k++) lo[k] = hi[k];
  }
}

#undef S00
#undef S01
#undef S02
#undef S11
#undef S12
#undef S22

}  // namespace adfem
