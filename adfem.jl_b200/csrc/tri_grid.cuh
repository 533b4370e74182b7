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
#include "grid_index.cuh"
#include "kernels.cuh"

namespace adfem {

constexpr int GRID_STRIP = 31;   // node columns per warp (lane 0 / lane 31 carry the neighbouring strip's cell / node column)
constexpr int GRID_WARPS = 8;

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ double shfl_up1(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_down1(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// packed symmetric index of (p,q), p <= q, 3x3: 00 01 02 11 12 22
#define S00 0
#define S01 1
#define S02 2
#define S11 3
#define S12 4
#define S22 5

// 32-bit variant of grid_row_prefix for offsets inside one node row
__device__ __forceinline__ int grid_row_prefix32(int j, int m, int A, int B) {
  const int jm = j < m ? j : m, j1 = j > 0 ? j - 1 : 0;
  return j * (1 + A + B) + (A + 1) * jm + (1 + B) * j1;
}

// local matrices (packed upper triangle) of both triangles of a VALID cell with corners (x0|x1, y0|y1)
template <int OP>
__device__ __forceinline__ void grid_cell_matrices(const DevMesh& m, double x0, double x1, double y0, double y1, double inv, const double kap[6], double T0[6],
                                                   double T1[6]) {
  if (OP == OP_LAPLACE && !m.heron) {
    // axis-aligned right triangles: grad lambda of T0 = [BL,BR,TL] is {(-a,-b),(a,0),(0,b)}, of T1 = [TL,BR,TR] {(-a,0),(0,-b),(a,b)},
    // a = dy/det, b = dx/det, det = dx*dy (exactly what geom_tri computes for these vertices, one reciprocal for both triangles)
    // inv = 1 / det is passed in: the callers compute it one row ahead, off the critical path
    const double dx = x1 - x0, dy = y1 - y0, det = dx * dy, a = dy * inv, b = dx * inv;
    const double a2 = a * a, b2 = b * b, ab = a * a + b * b;
    double c0 = 0.0, c1 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) { const double w = m.rule.w[k] * det; c0 += kap[k] * w; c1 += kap[3 + k] * w; }
    T0[0] = ab * c0; T0[1] = -a2 * c0; T0[2] = -b2 * c0; T0[3] = a2 * c0; T0[4] = 0.0; T0[5] = b2 * c0;
    T1[0] = a2 * c1; T1[1] = 0.0; T1[2] = -a2 * c1; T1[3] = b2 * c1; T1[4] = -b2 * c1; T1[5] = ab * c1;
  } else {
    const double2 BL = make_double2(x0, y0), BR = make_double2(x1, y0), TL = make_double2(x0, y1), TR = make_double2(x1, y1);
    Geom<2> G;
    geom_tri(BL, BR, TL, m.heron, G);
    local_matrix_scalar<2, 1, OP, 3>(m, G, [&](int k) { return kap[k]; }, [&](int s, double v) { T0[s] = v; });
    geom_tri(TL, BR, TR, m.heron, G);
    local_matrix_scalar<2, 1, OP, 3>(m, G, [&](int k) { return kap[3 + k]; }, [&](int s, double v) { T1[s] = v; });
  }
}

// the same for a cell with ARBITRARY corner positions (structured connectivity on mapped / jittered coordinates, template flag MAPPED below)
template <int OP>
__device__ __forceinline__ void grid_cell_matrices_pts(const DevMesh& m, double2 BL, double2 BR, double2 TL, double2 TR, const double kap[6], double T0[6],
                                                       double T1[6]) {
  Geom<2> G;
  geom_tri(BL, BR, TL, m.heron, G);
  local_matrix_scalar<2, 1, OP, 3>(m, G, [&](int k) { return kap[k]; }, [&](int s, double v) { T0[s] = v; });
  geom_tri(TL, BR, TR, m.heron, G);
  local_matrix_scalar<2, 1, OP, 3>(m, G, [&](int k) { return kap[3 + k]; }, [&](int s, double v) { T1[s] = v; });
}

// per-warp shared memory of k_grid_fwd: the output transpose buffer
constexpr int GRID_FWD_WARP_DOUBLES = GRID_STRIP * 7 + 1;
constexpr int GRID_FWD_SMEM = GRID_WARPS * GRID_FWD_WARP_DOUBLES * 8;

// Forward: vals[nnz] of the scalar operator OP on the structured triangulation (3 Gauss points per element).
// Node rows [r0, r1); grid: ceil(strips * chunks / GRID_WARPS) CTAs of GRID_WARPS warps; `rows_per_warp` node rows per warp.
// The coefficients of cell row i+2 are loaded into registers while row i is processed (three rotating buffers).
// MAPPED: structured connectivity, arbitrary node positions — the corner positions of the lane's cell column come from the coordinate array
// (16 B per node and strip, loaded one node row ahead) instead of xs / ys; everything else (index arithmetic, shuffles, stores) is unchanged.
template <int OP, int MINB, bool MAPPED = false>
__global__ void __launch_bounds__(GRID_WARPS * 32, MINB) k_grid_fwd(DevMesh m, GridTri gt, int r0, int r1, int rows_per_warp,
                                                                    const double* __restrict__ coef, double* __restrict__ vals) {
  extern __shared__ __align__(16) double grid_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int strips = (gt.m + 1 + GRID_STRIP - 1) / GRID_STRIP, chunks = (r1 - r0 + rows_per_warp - 1) / rows_per_warp;   // node rows [r0, r1)
  const long long gw = (long long)blockIdx.x * GRID_WARPS + wib;
  if (gw >= (long long)strips * chunks) return;
  const int strip = (int)(gw % strips), chunk = (int)(gw / strips);
  const int j0 = strip * GRID_STRIP, cj = j0 - 1 + lane;            // this lane's cell column = its node column
  const int i0 = r0 + chunk * rows_per_warp, i1 = min(i0 + rows_per_warp, r1);
  double* stage = grid_smem + (size_t)wib * GRID_FWD_WARP_DOUBLES;
  const bool colok = cj >= 0 && cj < gt.m;                           // this lane's cell column exists
  const double x0 = (!MAPPED && colok) ? __ldg(gt.xs + cj) : 0.0, x1 = (!MAPPED && colok) ? __ldg(gt.xs + cj + 1) : 1.0;
  // MAPPED: positions of nodes (r, cj) and (r, cj + 1)
  auto points = [&](int r, double2& a, double2& b) {
    if (colok && r >= 0 && r <= gt.n) {
      const double2* p = reinterpret_cast<const double2*>(m.coords) + ((size_t)r * (gt.m + 1) + cj);
      a = __ldg(p); b = __ldg(p + 1);
    }
  };
  double2 qb0 = make_double2(0.0, 0.0), qb1 = make_double2(1.0, 0.0), qt0 = make_double2(0.0, 1.0), qt1 = make_double2(1.0, 1.0);   // node rows i and i + 1
  const bool has_node = lane >= 1 && cj <= gt.m;                     // lanes 1..31 write node column cj
  const bool jl = cj > 0, jr = cj < gt.m;
  const int jend = min(j0 + GRID_STRIP, gt.m + 1);                   // one past the last node column of the strip
  const size_t kstride = (size_t)6 * gt.m;
  const double* kcol = coef + 6 * (size_t)max(cj, 0);                // + cell row * kstride

  // the 6 coefficients (2 triangles x 3 Gauss points) of cell (ci, cj): 48 contiguous, 16-byte aligned bytes
  auto load = [&](int ci, double k[6]) {
    if (colok && ci >= 0 && ci < gt.n) {
      const double2* p = reinterpret_cast<const double2*>(kcol + (size_t)ci * kstride);
      const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
      k[0] = a.x; k[1] = a.y; k[2] = b.x; k[3] = b.y; k[4] = c.x; k[5] = c.y;
    }
  };
  // coefficient buffers of cell rows i, i+1, i+2 rotate through kA, kB, kC (row loop unrolled by 3): a buffer is written only
  // by the load issued two rows ahead and never copied, so nothing waits for an in-flight load before its row is due
  double kA[6], kB[6], kC[6], pT0[6], pT1[6], cT0[6], cT1[6];
#pragma unroll
  for (int s = 0; s < 6; s++) { pT0[s] = pT1[s] = 0.0; kA[s] = kB[s] = kC[s] = 0.0; }
  double ya = 0.0, yb = 1.0;
  if constexpr (!MAPPED) { ya = __ldg(gt.ys + i0); yb = __ldg(gt.ys + min(i0 + 1, gt.n)); }   // ordinates of node rows i, i+1 (loaded one row ahead)
  else { points(i0, qb0, qb1); points(min(i0 + 1, gt.n), qt0, qt1); }
  if (i0 > 0) {                                                      // cell row i0-1 (below the first node row of the chunk)
    load(i0 - 1, kC);
    if constexpr (!MAPPED) {
      const double ym = __ldg(gt.ys + i0 - 1);
      if (colok) grid_cell_matrices<OP>(m, x0, x1, ym, ya, 1.0 / ((x1 - x0) * (ya - ym)), kC, pT0, pT1);
    } else {
      double2 qm0 = qb0, qm1 = qb1;
      points(i0 - 1, qm0, qm1);
      if (colok) grid_cell_matrices_pts<OP>(m, qm0, qm1, qb0, qb1, kC, pT0, pT1);
    }
  }
  load(i0, kA);
  if (i0 + 1 < i1) load(i0 + 1, kB);
  long long rowbase = grid_rowptr(i0, 0, gt.m, gt.n);
  double inv = 1.0 / ((x1 - x0) * (yb - ya));                        // 1 / det of cell row i, computed one row ahead
  auto row = [&](int i, const double kcur[6], double kfill[6]) {
    if (i + 2 < i1) load(i + 2, kfill);                              // two cell rows ahead: in flight while rows i, i+1 are computed
    double yc = 0.0, inv_next = 0.0;
    double2 qn0 = qt0, qn1 = qt1;
    if constexpr (!MAPPED) {
      yc = __ldg(gt.ys + min(i + 2, gt.n));                          // used by the next row
      inv_next = 1.0 / ((x1 - x0) * (yc - yb));                      // independent of this row's dependency chain
    } else {
      points(min(i + 2, gt.n), qn0, qn1);                            // node row i + 2: used by the next row
    }
    if (colok && i < gt.n) {
      if constexpr (!MAPPED) grid_cell_matrices<OP>(m, x0, x1, ya, yb, inv, kcur, cT0, cT1);
      else grid_cell_matrices_pts<OP>(m, qb0, qb1, qt0, qt1, kcur, cT0, cT1);
    } else {
#pragma unroll
      for (int s = 0; s < 6; s++) cT0[s] = cT1[s] = 0.0;
    }
    // contributions of the cell column to the left (lane - 1); sums of two left-hand values are formed by the sender
    const double l_pT1_21 = shfl_up1(pT1[S12]), l_pT1_22 = shfl_up1(pT1[S22]);
    const double l_cT0_11 = shfl_up1(cT0[S11]), l_cT1_11 = shfl_up1(cT1[S11]), l_cT1_12 = shfl_up1(cT1[S12]);
    const double vL = shfl_up1(pT1[S02] + cT0[S01]);
    const double vUL = shfl_up1(cT0[S12] + cT1[S01]);
    // the 7 entries of node (i, cj), each summed in ascending element order
    const double vD = l_pT1_21 + pT0[S02];
    const double vDR = pT0[S12] + pT1[S01];
    const double vC = ((((l_pT1_22 + pT0[S22]) + pT1[S00]) + l_cT0_11) + l_cT1_11) + cT0[S00];
    const double vR = pT1[S02] + cT0[S01];
    const double vU = l_cT1_12 + cT0[S02];
    const int A = i > 0, B = i < gt.n;
    const int pbase = grid_row_prefix32(j0, gt.m, A, B);
    if (has_node) {
      int o = grid_row_prefix32(cj, gt.m, A, B) - pbase;
      if (A) stage[o++] = vD;
      if (A && jr) stage[o++] = vDR;
      if (jl) stage[o++] = vL;
      stage[o++] = vC;
      if (jr) stage[o++] = vR;
      if (B && jl) stage[o++] = vUL;
      if (B) stage[o++] = vU;
    }
    __syncwarp();
    const int total = grid_row_prefix32(jend, gt.m, A, B) - pbase;
    double* out = vals + (rowbase + pbase) + lane;
#pragma unroll
    for (int k = 0; k < 7; k++) if (lane + 32 * k < total) out[32 * k] = stage[lane + 32 * k];
    __syncwarp();
    rowbase += grid_row_prefix32(gt.m + 1, gt.m, A, B);
    ya = yb; yb = yc; inv = inv_next;
    if constexpr (MAPPED) { qb0 = qt0; qb1 = qt1; qt0 = qn0; qt1 = qn1; }
#pragma unroll
    for (int s = 0; s < 6; s++) { pT0[s] = cT0[s]; pT1[s] = cT1[s]; }
  };
  for (int i = i0; i < i1; i += 3) {
    row(i, kA, kC);
    if (i + 1 < i1) row(i + 1, kB, kA);
    if (i + 2 < i1) row(i + 2, kC, kB);
  }
}

// gradients w.r.t. the 6 coefficients of a VALID cell from the upstream values t0/t1[p*3+q] of its two triangles
template <int OP>
__device__ __forceinline__ void grid_cell_adjoint(const DevMesh& m, double x0, double x1, double y0, double y1, double inv, const double t0[9], const double t1[9],
                                                  double gk[6]) {
  if (OP == OP_LAPLACE && !m.heron) {
    // sum_pq dK(p,q) grad_p . grad_q with the axis-aligned gradients of grid_cell_matrices
    const double dx = x1 - x0, dy = y1 - y0, det = dx * dy, a = dy * inv, b = dx * inv;
    const double a2 = a * a, b2 = b * b;
    const double s0 = a2 * (((t0[0] - t0[1]) - t0[3]) + t0[4]) + b2 * (((t0[0] - t0[2]) - t0[6]) + t0[8]);
    const double s1 = a2 * (((t1[0] - t1[2]) - t1[6]) + t1[8]) + b2 * (((t1[4] - t1[5]) - t1[7]) + t1[8]);
#pragma unroll
    for (int k = 0; k < 3; k++) { const double w = m.rule.w[k] * det; gk[k] = s0 * w; gk[3 + k] = s1 * w; }
  } else {
    const double2 BL = make_double2(x0, y0), BR = make_double2(x1, y0), TL = make_double2(x0, y1), TR = make_double2(x1, y1);
    Geom<2> G;
    geom_tri(BL, BR, TL, m.heron, G);
    local_adjoint_scalar<2, 1, OP, 3>(m, G, [&](int p, int q) { return t0[p * 3 + q]; }, [&](int k, double v) { gk[k] = v; });
    geom_tri(TL, BR, TR, m.heron, G);
    local_adjoint_scalar<2, 1, OP, 3>(m, G, [&](int p, int q) { return t1[p * 3 + q]; }, [&](int k, double v) { gk[3 + k] = v; });
  }
}

template <int OP>
__device__ __forceinline__ void grid_cell_adjoint_pts(const DevMesh& m, double2 BL, double2 BR, double2 TL, double2 TR, const double t0[9], const double t1[9],
                                                      double gk[6]) {
  Geom<2> G;
  geom_tri(BL, BR, TL, m.heron, G);
  local_adjoint_scalar<2, 1, OP, 3>(m, G, [&](int p, int q) { return t0[p * 3 + q]; }, [&](int k, double v) { gk[k] = v; });
  geom_tri(TL, BR, TR, m.heron, G);
  local_adjoint_scalar<2, 1, OP, 3>(m, G, [&](int p, int q) { return t1[p * 3 + q]; }, [&](int k, double v) { gk[3 + k] = v; });
}

// per-warp shared memory of k_grid_adj: a ring of 3 raw node rows (up to 32*7 CSR entries each) + the output transpose buffer
constexpr int GRID_ADJ_WARP_DOUBLES = 4 * 32 * 7;
constexpr int GRID_ADJ_SMEM = GRID_WARPS * GRID_ADJ_WARP_DOUBLES * 8;

// Adjoint: grad_coef[e*3 + k] from upstream dvals[nnz].  Lane l owns node column j0+l (32 columns, the last one only feeds
// its left neighbour) and cell column j0+l for l < 31.  The CSR entries of node row ci+3 (one contiguous run per strip) are
// requested with asynchronous copies while cell row ci is processed.
template <int OP, int MINB, bool MAPPED = false>
__global__ void __launch_bounds__(GRID_WARPS * 32, MINB) k_grid_adj(DevMesh m, GridTri gt, int r0, int r1, int rows_per_warp,
                                                                    const double* __restrict__ dvals, double* __restrict__ grad_coef) {
  extern __shared__ __align__(16) double grid_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int strips = (gt.m + GRID_STRIP - 1) / GRID_STRIP, chunks = (r1 - r0 + rows_per_warp - 1) / rows_per_warp;   // cell rows [r0, r1)
  const long long gw = (long long)blockIdx.x * GRID_WARPS + wib;
  if (gw >= (long long)strips * chunks) return;
  const int strip = (int)(gw % strips), chunk = (int)(gw / strips);
  const int j0 = strip * GRID_STRIP, j = j0 + lane;                  // node column == cell column of this lane
  const int c0 = r0 + chunk * rows_per_warp, c1 = min(c0 + rows_per_warp, r1);
  double* ring = grid_smem + (size_t)wib * GRID_ADJ_WARP_DOUBLES;    // ring[slot][32*7]
  double* stage = ring + 3 * 32 * 7;
  const bool node_ok = j <= gt.m, cell_ok = lane < GRID_STRIP && j < gt.m;
  const bool jl = j > 0, jr = j < gt.m;
  const double x0 = (!MAPPED && cell_ok) ? __ldg(gt.xs + j) : 0.0, x1 = (!MAPPED && cell_ok) ? __ldg(gt.xs + j + 1) : 1.0;
  auto points = [&](int r, double2& a, double2& b) {                 // MAPPED: positions of nodes (r, j) and (r, j + 1)
    if (cell_ok && r >= 0 && r <= gt.n) {
      const double2* p = reinterpret_cast<const double2*>(m.coords) + ((size_t)r * (gt.m + 1) + j);
      a = __ldg(p); b = __ldg(p + 1);
    }
  };
  double2 qb0 = make_double2(0.0, 0.0), qb1 = make_double2(1.0, 0.0), qt0 = make_double2(0.0, 1.0), qt1 = make_double2(1.0, 1.0);
  const int jend = min(j0 + 32, gt.m + 1);
  const int ncell6 = 6 * min(GRID_STRIP, gt.m - j0);

  long long fb = grid_rowptr(c0, 0, gt.m, gt.n);                     // CSR offset of the node row that is requested next
  int fi = c0;                                                       // ... and its index
  // request node row fi of the strip (one contiguous run of CSR entries, coalesced) into ring slot `slot`; one commit group per call
  auto request = [&](int slot) {
    if (fi <= gt.n) {
      const int A = fi > 0, B = fi < gt.n;
      const int pbase = grid_row_prefix32(j0, gt.m, A, B);
      const int total = grid_row_prefix32(jend, gt.m, A, B) - pbase;
      const double* in = dvals + (fb + pbase) + lane;
      double* dst = ring + slot * (32 * 7) + lane;
#pragma unroll
      for (int k = 0; k < 7; k++) if (lane + 32 * k < total) cp_async8(dst + 32 * k, in + 32 * k);
      fb += grid_row_prefix32(gt.m + 1, gt.m, A, B);
    }
    fi++;
    cp_async_commit();
  };
  // the 7 logical entries [D, DR, L, C, R, UL, U] of node (i, j) from ring slot `slot`, missing ones = 0
  auto unpack = [&](int i, int slot, double r[7]) {
    const int A = i > 0, B = i < gt.n;
    const double* row = ring + slot * (32 * 7);
#pragma unroll
    for (int k = 0; k < 7; k++) r[k] = 0.0;
    if (node_ok && i <= gt.n) {
      int o = grid_row_prefix32(j, gt.m, A, B) - grid_row_prefix32(j0, gt.m, A, B);
      if (A) r[0] = row[o++];
      if (A && jr) r[1] = row[o++];
      if (jl) r[2] = row[o++];
      r[3] = row[o++];
      if (jr) r[4] = row[o++];
      if (B && jl) r[5] = row[o++];
      if (B) r[6] = row[o++];
    }
  };
  double lo[7], hi[7];                                               // node rows ci and ci+1
  request(0); request(1); request(2);
  cp_async_wait<2>();
  __syncwarp();
  unpack(c0, 0, lo);
  double* out = grad_coef + 6 * ((size_t)c0 * gt.m + j0) + lane;
  const size_t ostride = (size_t)6 * gt.m;
  double ya = 0.0, yb = 1.0;
  if constexpr (!MAPPED) { ya = __ldg(gt.ys + c0); yb = __ldg(gt.ys + c0 + 1); }   // ordinates of node rows ci, ci+1 (loaded one row ahead)
  else { points(c0, qb0, qb1); points(c0 + 1, qt0, qt1); }
  int slot = 0;                                                      // ring slot of node row ci
  double inv = 1.0 / ((x1 - x0) * (yb - ya));                        // 1 / det of cell row ci, computed one row ahead
  for (int ci = c0; ci < c1; ci++, out += ostride) {
    double yc = 0.0, inv_next = 0.0;
    double2 qn0 = qt0, qn1 = qt1;
    if constexpr (!MAPPED) { yc = __ldg(gt.ys + min(ci + 2, gt.n)); inv_next = 1.0 / ((x1 - x0) * (yc - yb)); }
    else points(min(ci + 2, gt.n), qn0, qn1);
    const int slot1 = slot == 2 ? 0 : slot + 1;
    cp_async_wait<1>();                                              // node row ci+1 has landed (row ci+2 may be in flight)
    __syncwarp();                                                    // ... for every lane; every lane is also done reading slot (row ci)
    unpack(ci + 1, slot1, hi);
    request(slot);                                                   // node row ci+3 into the slot of row ci
    // rows of the right-hand node column (lane + 1): BR = (ci, j+1), TR = (ci+1, j+1)
    const double br_L = shfl_down1(lo[2]), br_C = shfl_down1(lo[3]), br_UL = shfl_down1(lo[5]), br_U = shfl_down1(lo[6]);
    const double tr_D = shfl_down1(hi[0]), tr_L = shfl_down1(hi[2]), tr_C = shfl_down1(hi[3]);
    double gk[6];
#pragma unroll
    for (int s = 0; s < 6; s++) gk[s] = 0.0;
    if (cell_ok) {
      // T0 = [BL, BR, TL], T1 = [TL, BR, TR]: dK(p,q) = entry (row of local p, column of local q)
      const double t0[9] = {lo[3], lo[4], lo[6], br_L, br_C, br_UL, hi[0], hi[1], hi[3]};
      const double t1[9] = {hi[3], hi[1], hi[4], br_UL, br_C, br_U, tr_L, tr_D, tr_C};
      if constexpr (!MAPPED) grid_cell_adjoint<OP>(m, x0, x1, ya, yb, inv, t0, t1, gk);
      else grid_cell_adjoint_pts<OP>(m, qb0, qb1, qt0, qt1, t0, t1, gk);
    }
    // the 6 gradients of every cell of the strip leave through the transpose buffer as one contiguous, coalesced run
    if (lane < GRID_STRIP) {
#pragma unroll
      for (int s = 0; s < 6; s++) stage[lane * 7 + s] = gk[s];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 6; k++) { const int t = lane + 32 * k; if (t < ncell6) out[32 * k] = stage[(t / 6) * 7 + t % 6]; }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 7; k++) lo[k] = hi[k];
    ya = yb; yb = yc; inv = inv_next;
    if constexpr (MAPPED) { qb0 = qt0; qb1 = qt1; qt0 = qn0; qt1 = qn1; }
    slot = slot1;
  }
}

// ---- source term on the structured triangulation (FemSourceScalar_forward/_backward, deps/MFEM/FemSource1/FemSourceScalar.h:4-32) -------------
// per-vertex load integrals of both triangles of a VALID cell: v[r] = sum_k f_k lambda_r(x_k) w_k
__device__ __forceinline__ void grid_cell_loads_pts(const DevMesh& m, double2 BL, double2 BR, double2 TL, double2 TR, const double f[6], double T0[3], double T1[3]) {
  Geom<2> G0, G1;
  geom_tri(BL, BR, TL, m.heron, G0);
  geom_tri(TL, BR, TR, m.heron, G1);
#pragma unroll
  for (int r = 0; r < 3; r++) T0[r] = T1[r] = 0.0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double L[3]; bary<2>(m.rule, k, L);
    const double w0 = f[k] * (m.rule.w[k] * G0.wscale), w1 = f[3 + k] * (m.rule.w[k] * G1.wscale);
#pragma unroll
    for (int r = 0; r < 3; r++) { T0[r] += L[r] * w0; T1[r] += L[r] * w1; }
  }
}
__device__ __forceinline__ void grid_cell_loads(const DevMesh& m, double x0, double x1, double y0, double y1, const double f[6], double T0[3], double T1[3]) {
  grid_cell_loads_pts(m, make_double2(x0, y0), make_double2(x1, y0), make_double2(x0, y1), make_double2(x1, y1), f, T0, T1);
}

// rhs[node] for node rows [r0, r1): lane l owns cell column / node column j0-1+l as in k_grid_fwd; one value per node, so
// consecutive lanes write consecutive addresses and no transpose is needed
template <bool MAPPED = false>
__global__ void __launch_bounds__(GRID_WARPS * 32) k_grid_source_fwd(DevMesh m, GridTri gt, int r0, int r1, int rows_per_warp, const double* __restrict__ f,
                                                                     double* __restrict__ rhs) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int strips = (gt.m + 1 + GRID_STRIP - 1) / GRID_STRIP, chunks = (r1 - r0 + rows_per_warp - 1) / rows_per_warp;
  const long long gw = (long long)blockIdx.x * GRID_WARPS + wib;
  if (gw >= (long long)strips * chunks) return;
  const int strip = (int)(gw % strips), chunk = (int)(gw / strips);
  const int j0 = strip * GRID_STRIP, cj = j0 - 1 + lane;
  const int i0 = r0 + chunk * rows_per_warp, i1 = min(i0 + rows_per_warp, r1);
  const bool colok = cj >= 0 && cj < gt.m, has_node = lane >= 1 && cj <= gt.m;
  const double x0 = (!MAPPED && colok) ? __ldg(gt.xs + cj) : 0.0, x1 = (!MAPPED && colok) ? __ldg(gt.xs + cj + 1) : 1.0;
  auto points = [&](int r, double2& a, double2& b) {                 // MAPPED: positions of nodes (r, cj) and (r, cj + 1)
    if (colok && r >= 0 && r <= gt.n) {
      const double2* p = reinterpret_cast<const double2*>(m.coords) + ((size_t)r * (gt.m + 1) + cj);
      a = __ldg(p); b = __ldg(p + 1);
    }
  };
  double2 qb0 = make_double2(0.0, 0.0), qb1 = make_double2(1.0, 0.0), qt0 = make_double2(0.0, 1.0), qt1 = make_double2(1.0, 1.0);
  const size_t kstride = (size_t)6 * gt.m;
  const double* kcol = f + 6 * (size_t)max(cj, 0);
  auto load = [&](int ci, double k[6]) {
    if (colok && ci >= 0 && ci < gt.n) {
      const double2* p = reinterpret_cast<const double2*>(kcol + (size_t)ci * kstride);
      const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
      k[0] = a.x; k[1] = a.y; k[2] = b.x; k[3] = b.y; k[4] = c.x; k[5] = c.y;
    }
  };
  double kA[6], kB[6], kC[6], pT0[3], pT1[3], cT0[3], cT1[3];
#pragma unroll
  for (int s = 0; s < 6; s++) kA[s] = kB[s] = kC[s] = 0.0;
#pragma unroll
  for (int s = 0; s < 3; s++) pT0[s] = pT1[s] = 0.0;
  double ya = 0.0, yb = 1.0;
  if constexpr (!MAPPED) { ya = __ldg(gt.ys + i0); yb = __ldg(gt.ys + min(i0 + 1, gt.n)); }
  else { points(i0, qb0, qb1); points(min(i0 + 1, gt.n), qt0, qt1); }
  if (i0 > 0) {
    load(i0 - 1, kC);
    if constexpr (!MAPPED) { if (colok) grid_cell_loads(m, x0, x1, __ldg(gt.ys + i0 - 1), ya, kC, pT0, pT1); }
    else {
      double2 qm0 = qb0, qm1 = qb1;
      points(i0 - 1, qm0, qm1);
      if (colok) grid_cell_loads_pts(m, qm0, qm1, qb0, qb1, kC, pT0, pT1);
    }
  }
  load(i0, kA);
  if (i0 + 1 < i1) load(i0 + 1, kB);
  auto row = [&](int i, const double kcur[6], double kfill[6]) {
    if (i + 2 < i1) load(i + 2, kfill);
    double yc = 0.0;
    double2 qn0 = qt0, qn1 = qt1;
    if constexpr (!MAPPED) yc = __ldg(gt.ys + min(i + 2, gt.n));
    else points(min(i + 2, gt.n), qn0, qn1);
    if (colok && i < gt.n) {
      if constexpr (!MAPPED) grid_cell_loads(m, x0, x1, ya, yb, kcur, cT0, cT1);
      else grid_cell_loads_pts(m, qb0, qb1, qt0, qt1, kcur, cT0, cT1);
    } else {
#pragma unroll
      for (int s = 0; s < 3; s++) cT0[s] = cT1[s] = 0.0;
    }
    // incident (cell, triangle, local vertex) in ascending element order: (i-1,j-1).T1.2, (i-1,j).T0.2, (i-1,j).T1.0, (i,j-1).T0.1, (i,j-1).T1.1, (i,j).T0.0
    const double l_p = shfl_up1(pT1[2]), l_c0 = shfl_up1(cT0[1]), l_c1 = shfl_up1(cT1[1]);
    const double v = ((((l_p + pT0[2]) + pT1[0]) + l_c0) + l_c1) + cT0[0];
    if (has_node) rhs[(size_t)i * (gt.m + 1) + cj] = v;
    ya = yb; yb = yc;
    if constexpr (MAPPED) { qb0 = qt0; qb1 = qt1; qt0 = qn0; qt1 = qn1; }
#pragma unroll
    for (int s = 0; s < 3; s++) { pT0[s] = cT0[s]; pT1[s] = cT1[s]; }
  };
  for (int i = i0; i < i1; i += 3) {
    row(i, kA, kC);
    if (i + 1 < i1) row(i + 1, kB, kA);
    if (i + 2 < i1) row(i + 2, kC, kB);
  }
}

// grad_f[e*3 + k] = sum_r lambda_r(x_k) w_k grad_rhs[node_r] for cell rows [r0, r1): lane l owns node column = cell column j0+l (l < 31)
template <bool MAPPED = false>
__global__ void __launch_bounds__(GRID_WARPS * 32) k_grid_source_adj(DevMesh m, GridTri gt, int r0, int r1, int rows_per_warp, const double* __restrict__ grad_rhs,
                                                                     double* __restrict__ grad_f) {
  __shared__ double stage_all[GRID_WARPS][GRID_STRIP * 7];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int strips = (gt.m + GRID_STRIP - 1) / GRID_STRIP, chunks = (r1 - r0 + rows_per_warp - 1) / rows_per_warp;
  const long long gw = (long long)blockIdx.x * GRID_WARPS + wib;
  if (gw >= (long long)strips * chunks) return;
  const int strip = (int)(gw % strips), chunk = (int)(gw / strips);
  const int j0 = strip * GRID_STRIP, j = j0 + lane;
  const int c0 = r0 + chunk * rows_per_warp, c1 = min(c0 + rows_per_warp, r1);
  double* stage = stage_all[wib];
  const bool node_ok = j <= gt.m, cell_ok = lane < GRID_STRIP && j < gt.m;
  const double x0 = (!MAPPED && cell_ok) ? __ldg(gt.xs + j) : 0.0, x1 = (!MAPPED && cell_ok) ? __ldg(gt.xs + j + 1) : 1.0;
  auto points = [&](int r, double2& a, double2& b) {                 // MAPPED: positions of nodes (r, j) and (r, j + 1)
    if (cell_ok && r >= 0 && r <= gt.n) {
      const double2* p = reinterpret_cast<const double2*>(m.coords) + ((size_t)r * (gt.m + 1) + j);
      a = __ldg(p); b = __ldg(p + 1);
    }
  };
  double2 qb0 = make_double2(0.0, 0.0), qb1 = make_double2(1.0, 0.0), qt0 = make_double2(0.0, 1.0), qt1 = make_double2(1.0, 1.0);
  const int ncell6 = 6 * min(GRID_STRIP, gt.m - j0);
  auto node = [&](int i) { return (node_ok && i <= gt.n) ? __ldg(grad_rhs + (size_t)i * (gt.m + 1) + j) : 0.0; };
  double glo = node(c0), ghi = node(c0 + 1), gnx = node(c0 + 2);     // node rows ci, ci+1, ci+2 (one row ahead)
  double ya = 0.0, yb = 1.0;
  if constexpr (!MAPPED) { ya = __ldg(gt.ys + c0); yb = __ldg(gt.ys + c0 + 1); }
  else { points(c0, qb0, qb1); points(c0 + 1, qt0, qt1); }
  double* out = grad_f + 6 * ((size_t)c0 * gt.m + j0) + lane;
  const size_t ostride = (size_t)6 * gt.m;
  for (int ci = c0; ci < c1; ci++, out += ostride) {
    const double gn2 = node(ci + 3);
    double yc = 0.0;
    double2 qn0 = qt0, qn1 = qt1;
    if constexpr (!MAPPED) yc = __ldg(gt.ys + min(ci + 2, gt.n));
    else points(min(ci + 2, gt.n), qn0, qn1);
    const double g_br = shfl_down1(glo), g_tr = shfl_down1(ghi);
    double gk[6];
#pragma unroll
    for (int s = 0; s < 6; s++) gk[s] = 0.0;
    if (cell_ok) {
      const double2 BL = MAPPED ? qb0 : make_double2(x0, ya), BR = MAPPED ? qb1 : make_double2(x1, ya), TL = MAPPED ? qt0 : make_double2(x0, yb),
                    TR = MAPPED ? qt1 : make_double2(x1, yb);
      Geom<2> G0, G1;
      geom_tri(BL, BR, TL, m.heron, G0);
      geom_tri(TL, BR, TR, m.heron, G1);
      const double t0[3] = {glo, g_br, ghi}, t1[3] = {ghi, g_br, g_tr};   // T0 = [BL, BR, TL], T1 = [TL, BR, TR]
#pragma unroll
      for (int k = 0; k < 3; k++) {
        double L[3]; bary<2>(m.rule, k, L);
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int r = 0; r < 3; r++) { a += L[r] * (m.rule.w[k] * G0.wscale) * t0[r]; b += L[r] * (m.rule.w[k] * G1.wscale) * t1[r]; }
        gk[k] = a; gk[3 + k] = b;
      }
    }
    __syncwarp();
    if (lane < GRID_STRIP) {
#pragma unroll
      for (int s = 0; s < 6; s++) stage[lane * 7 + s] = gk[s];
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 6; k++) { const int t = lane + 32 * k; if (t < ncell6) out[32 * k] = stage[(t / 6) * 7 + t % 6]; }
    glo = ghi; ghi = gnx; gnx = gn2;
    ya = yb; yb = yc;
    if constexpr (MAPPED) { qb0 = qt0; qb1 = qt1; qt0 = qn0; qt1 = qn1; }
  }
}

#undef S00
#undef S01
#undef S02
#undef S11
#undef S12
#undef S22

}  // namespace adfem
