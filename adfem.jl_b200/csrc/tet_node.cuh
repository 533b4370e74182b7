// Forward CSR assembly of the P1 tetrahedral ELASTICITY operator on the structured grid `Mesh3(n, n, l, h)` (BASELINE config 5), second
// generation (round 2).  Replaces the one-warp-per-node kernel of tet_grid.cuh, which profiled at 1950 warp instructions per node (idle lanes
// for the 8-tetrahedron parity, table reads from global memory, a source-list gather with 32 dependent shared-memory reads per lane):
// profiles/ncu_r02_cfg5_start.md.
//
// Pass 1  k_tet_presum_x   streams the per-Gauss-point 6x6 tangents once (1152 B per tetrahedron, contiguous per cube), sums them with the Gauss
//         weights and writes the 36 summed entries in an X-FASTEST, parity-split layout
//             hx[((((cj*l + ck)*5 + t) * NBLK + (px >> 5)) * 36 + c) * 32 + (px & 31)],   px = (ci >> 1) + (ci & 1) * HALF
//         (blocks of 32 x-positions, so that the 18 loads of a lane are one pointer + immediate offsets) through a shared-memory transpose, so that pass 2's lanes (same-parity nodes of one grid line, cube index stride 2) read
//         consecutive doubles.
// Pass 2  k_tet_node_fwd   one THREAD per (node, column component b); a CTA covers 64 consecutive nodes: warps 0-2 the 32 nodes of the parity with
//         32 incident tetrahedra and 19 row slots, warps 3-5 those of the other parity (8 tetrahedra, 7 slots), warp w % 3 = b.  All lanes of a warp
//         walk the same incidence table (constant memory, warp-uniform) in ascending element order, so there are no idle lanes and no
//         per-lane table reads; a lane evaluates column b of the 3 x 12 row block of its node in the tetrahedron (only the three columns
//         of the tangent that component b touches are loaded) and adds the 4 x 3 results into its own column of a shared-memory accumulator
//         acc[slot][a][lane] (conflict-free, no atomics, fixed summation order = ascending element id).  The CTA then writes the three
//         contiguous CSR runs (a = 0, 1, 2) of its 64 nodes with coalesced stores.
// Geometry: on the rectilinear grid grad lambda_q = gcoef[q] / (hx, hy, hz) with gcoef in {0, +-1, +-1/2} per (parity, tetrahedron, vertex)
// and |det| = detfac * hx*hy*hz (tables built on the host from the unit cube with the same geom_tet).
// CSR layout as everywhere: scalar row r (start rs, length len) holds entry (a, b, j) at 3*(a*nnz + rs) + b*len + j.
#pragma once
#include "tet_grid.cuh"

namespace adfem {

struct TetNodeConst {
  int ninc[2], nslot[2], present[2];
  int heavy;                       // the parity with 32 incident tetrahedra / 19 row slots (the other one has 8 / 7)
  signed char inc[2][32][5];       // cube offset (3), tetrahedron of the cube, local index of the node
  unsigned char cslot[2][32][4];   // compact structural slot (rank of the 27-neighbourhood id among the parity's present slots) of vertex q
  unsigned char slot27[2][19];     // 27-neighbourhood id of compact slot
  // corner tetrahedra (apex A + the three neighbours X, Y, Z of A along the axes): grad lambda_X = (sgn_x / hx, 0, 0) etc., grad lambda_A = -(sum)
  signed char kind[2][32];         // 1 = corner tetrahedron, 0 = central tetrahedron of the cube (generic path)
  signed char qax[2][32][4];       // local vertex index of X, Y, Z and A
  signed char axp[2][32];          // axis of the node itself (0, 1, 2) or -1 when the node is the apex
  double sgn[2][32][3];            // +-1
  double gcoef[2][32][4][3];       // unit-cube barycentric gradients (0, +-1, +-0.5)
  double detfac[2][32];            // |det| of the unit-cube tetrahedron (1 or 2)
};

inline void build_tet_node_const(const TetGridTables& T, TetNodeConst& C) {
  std::memset(&C, 0, sizeof(C));
  for (int par = 0; par < 2; par++) {
    C.ninc[par] = T.ninc[par];
    C.present[par] = T.present[par];
    int ns = 0;
    for (int s = 0; s < 27; s++) if ((T.present[par] >> s) & 1) C.slot27[par][ns++] = (unsigned char)s;
    C.nslot[par] = ns;
    for (int t = 0; t < T.ninc[par]; t++) {
      for (int c = 0; c < 5; c++) C.inc[par][t][c] = (signed char)T.inc[par][t][c];
      double X[4][3];
      for (int q = 0; q < 4; q++) {
        int cs = 0;
        for (int s = 0; s < T.vslot[par][t][q]; s++) cs += (T.present[par] >> s) & 1;
        C.cslot[par][t][q] = (unsigned char)cs;
        for (int c = 0; c < 3; c++) X[q][c] = (double)(T.voff[par][t][q][c] - T.inc[par][t][c]);      // position in the unit cube
      }
      Geom<3> G; geom_tet(X, G);
      for (int q = 0; q < 4; q++)
        for (int c = 0; c < 3; c++) C.gcoef[par][t][q][c] = G.gL[q][c];
      C.detfac[par][t] = G.wscale < 0 ? -G.wscale : G.wscale;
      // classify: a corner tetrahedron has three vertices whose gradient has exactly one non-zero (+-1) component, one per axis
      int ax_of[4], nax = 0, seen[3] = {0, 0, 0}, apex = -1;
      for (int q = 0; q < 4; q++) {
        int nz = 0, which = -1;
        for (int c = 0; c < 3; c++) if (G.gL[q][c] != 0.0) { nz++; which = c; }
        ax_of[q] = nz == 1 ? which : -1;
        if (nz == 1 && (G.gL[q][which] == 1.0 || G.gL[q][which] == -1.0) && !seen[which]) { seen[which] = 1; nax++; } else apex = q;
      }
      if (nax == 3 && apex >= 0) {
        C.kind[par][t] = 1;
        C.qax[par][t][3] = (signed char)apex;
        for (int q = 0; q < 4; q++)
          if (q != apex) { C.qax[par][t][ax_of[q]] = (signed char)q; C.sgn[par][t][ax_of[q]] = G.gL[q][ax_of[q]]; }
        const int p = T.inc[par][t][4];
        C.axp[par][t] = (signed char)(p == apex ? -1 : ax_of[p]);
      }
    }
  }
  C.heavy = C.ninc[1] > C.ninc[0] ? 1 : 0;
}

// x-extent of the summed-tangent scratch: two parity halves, each padded to a multiple of 32 so that 32-blocks never straddle the halves
ADFEM_HD int tn_half(int n) { return (((n + 1) >> 1) + 31) & ~31; }
ADFEM_HD int tn_nblk(int n) { return 2 * tn_half(n) / 32; }
ADFEM_HD size_t tn_scratch_doubles(int n, int l) { return (size_t)n * l * 5 * tn_nblk(n) * 36 * 32; }
// per-axis spacing tables of the node kernel: [1/hx (n) | 1/hy (n) | 1/hz (l) | hx (n) | hy (n) | hz (l)]
struct TetSpacing { const double *ihx, *ihy, *ihz, *hx, *hy, *hz; };

#ifdef __CUDACC__
static __constant__ TetNodeConst c_tn;

constexpr int TPX_CUBES = 64, TPX_THREADS = 256;
// grid: ((cj*l + ck) * nxb + xb) * 5 + t — one CTA per (grid line, block of 64 cubes, tetrahedron of the cube): load + sum, one barrier, transposed
// store.  No loop over t inside the CTA: the loads of one CTA overlap the stores of the others resident on the SM (8 CTAs of 18.7 KB).
static __global__ void __launch_bounds__(TPX_THREADS) k_tet_presum_x(int n, int l, int z0, int nz, QuadRule rule, int g, const double* __restrict__ coef,
                                                                     double* __restrict__ hx) {
  __shared__ double tile[36 * (TPX_CUBES + 1)];
  const int nxb = (n + TPX_CUBES - 1) / TPX_CUBES;
  const int t = blockIdx.x % 5, rest = blockIdx.x / 5, xb = rest % nxb, lz = rest / nxb, ck = z0 + lz % nz, cj = lz / nz, line = cj * l + ck;
  const int ci0 = xb * TPX_CUBES, ncx = min(TPX_CUBES, n - ci0), half = tn_half(n), nblk = tn_nblk(n);
  const int tid = threadIdx.x;
  for (int idx = tid; idx < ncx * 36; idx += TPX_THREADS) {
    const int cil = idx / 36, c = idx - 36 * cil;
    const size_t e = (size_t)5 * (((size_t)(ci0 + cil) * n + cj) * l + ck) + t;
    const double* p = coef + e * g * 36 + c;
    double s = 0.0;
    for (int k = 0; k < g; k++) s += __ldg(p + 36 * k) * rule.w[k];
    tile[c * (TPX_CUBES + 1) + cil] = s;
  }
  __syncthreads();
  for (int idx = tid; idx < 36 * TPX_CUBES; idx += TPX_THREADS) {
    const int c = idx / TPX_CUBES, r = idx - TPX_CUBES * c, hh = r >> 5, xx = r & 31, cil = 2 * xx + hh;     // ci0 is even: px = hh*half + ci0/2 + xx
    const int px = hh * half + (ci0 >> 1) + xx;
    if (cil < ncx) hx[((((size_t)line * 5 + t) * nblk + (px >> 5)) * 36 + c) * 32 + (px & 31)] = tile[c * (TPX_CUBES + 1) + cil];
  }
}

// The same pass for a compile-time number of Gauss points: 16-byte loads (two adjacent tangent entries, 18 lanes per tetrahedron), all loads of a
// batch of three items in flight before the first is used (12 x 16 B per thread), no per-item 64-bit index arithmetic.  The generic kernel above
// needed 125 warp instructions per tetrahedron (profiles/ncu_r02_cfg5_tet_node_v2.md: issue slots 35 % busy, DRAM at 54 %).
template <int G>
static __global__ void __launch_bounds__(TPX_THREADS) k_tet_presum_xg(int n, int l, int z0, int nz, QuadRule rule, const double* __restrict__ coef,
                                                                      double* __restrict__ hx) {
  __shared__ double tile[36 * (TPX_CUBES + 1)];
  const int nxb = (n + TPX_CUBES - 1) / TPX_CUBES;
  const int t = blockIdx.x % 5, rest = blockIdx.x / 5, xb = rest % nxb, lz = rest / nxb, ck = z0 + lz % nz, cj = lz / nz;
  const int line = cj * l + ck;
  const int ci0 = xb * TPX_CUBES, ncx = min(TPX_CUBES, n - ci0), half = tn_half(n), nblk = tn_nblk(n);
  const int tid = threadIdx.x;
  double w[G];
#pragma unroll
  for (int k = 0; k < G; k++) w[k] = rule.w[k];
  const size_t cube_stride = (size_t)5 * n * l * G * 36;                      // doubles between the tetrahedra t of cubes ci and ci + 1
  const double2* base = reinterpret_cast<const double2*>(coef + ((size_t)5 * (((size_t)ci0 * n + cj) * l + ck) + t) * G * 36);
  const int nitem = ncx * 18;
  constexpr int NIT = (TPX_CUBES * 18 + TPX_THREADS - 1) / TPX_THREADS, NB = 3;
#pragma unroll
  for (int it0 = 0; it0 < NIT; it0 += NB) {
    double2 v[NB][G];
#pragma unroll
    for (int u = 0; u < NB; u++) {
      const int idx = tid + (it0 + u) * TPX_THREADS;
      if (it0 + u < NIT && idx < nitem) {
        const int cil = idx / 18, c2 = idx - 18 * cil;
        const double2* p = base + (cil * cube_stride) / 2 + c2;
#pragma unroll
        for (int k = 0; k < G; k++) v[u][k] = __ldg(p + 18 * k);
      }
    }
#pragma unroll
    for (int u = 0; u < NB; u++) {
      const int idx = tid + (it0 + u) * TPX_THREADS;
      if (it0 + u < NIT && idx < nitem) {
        const int cil = idx / 18, c2 = idx - 18 * cil;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < G; k++) { s0 += v[u][k].x * w[k]; s1 += v[u][k].y * w[k]; }
        tile[(2 * c2) * (TPX_CUBES + 1) + cil] = s0;
        tile[(2 * c2 + 1) * (TPX_CUBES + 1) + cil] = s1;
      }
    }
  }
  __syncthreads();
  double* out = hx + ((size_t)line * 5 + t) * nblk * (36 * 32);
#pragma unroll 3
  for (int idx = tid; idx < 36 * TPX_CUBES; idx += TPX_THREADS) {
    const int c = idx / TPX_CUBES, r = idx - TPX_CUBES * c, hh = r >> 5, xx = r & 31, cil = 2 * xx + hh;     // ci0 is even: px = hh*half + ci0/2 + xx
    const int px = hh * half + (ci0 >> 1) + xx;
    if (cil < ncx) out[((size_t)(px >> 5) * 36 + c) * 32 + (px & 31)] = tile[c * (TPX_CUBES + 1) + cil];
  }
}

constexpr int TN_NODES = 64, TN_THREADS = 192;
constexpr int TN_SLOT = 97;                                                  // 3 x 32 accumulators per slot + 1: the store phase walks the slots (odd stride: no bank conflicts)
constexpr int TN_ACC1 = 19 * TN_SLOT, TN_ACC0 = 7 * TN_SLOT;                  // doubles per warp of the 19-slot / 7-slot parity
constexpr int TN_SMEM_BYTES = (3 * TN_ACC1 + 3 * TN_ACC0) * 8, TN_SMEM_BYTES2 = (6 * TN_ACC1 + 3 * TN_ACC0) * 8;

template <int B>
__device__ __forceinline__ void tn_accumulate(const GridTet& gt, const TetSpacing& sp, int par, bool valid, int i, int j, int k, int lane,
                                              const double* __restrict__ hx, double* __restrict__ acc, int t0 = 0, int t1 = 64) {
  // columns of the tangent that component B of a strain-displacement column touches (device_fem.cuh bdot<3>)
  constexpr int COL[3][3] = {{0, 4, 5}, {1, 3, 5}, {2, 3, 4}};
  const int n = gt.n, l = gt.l, half = tn_half(n), nblk = tn_nblk(n);
  const int ninc = min(c_tn.ninc[par], t1);
  for (int t = t0; t < ninc; t++) {
    const int ci = i + c_tn.inc[par][t][0], cj = j + c_tn.inc[par][t][1], ck = k + c_tn.inc[par][t][2];
    if (!valid || ci < 0 || ci >= n || cj < 0 || cj >= n || ck < 0 || ck >= l) continue;
    const double ihx = __ldg(sp.ihx + ci), ihy = __ldg(sp.ihy + cj), ihz = __ldg(sp.ihz + ck);
    const double ws = c_tn.detfac[par][t] * (__ldg(sp.hx + ci) * __ldg(sp.hy + cj) * __ldg(sp.hz + ck));
    const int px = (ci & 1) * half + (ci >> 1);
    const double* he = hx + ((((size_t)cj * l + ck) * 5 + c_tn.inc[par][t][3]) * nblk + (px >> 5)) * (36 * 32) + (px & 31);
    if (c_tn.kind[par][t]) {
      // corner tetrahedron: single-axis gradients.  hb_q[r] = g_q H[r][SEL[B][axis of q]]; the apex blocks follow from sum_q grad lambda_q = 0
      constexpr int SEL[3][3] = {{0, 5, 4}, {5, 1, 3}, {4, 3, 2}};       // bdot<3>(c, g e_ax, v) = g v[SEL[c][ax]]
      const double gX = c_tn.sgn[par][t][0] * ihx, gY = c_tn.sgn[par][t][1] * ihy, gZ = c_tn.sgn[par][t][2] * ihz;
      const double gXw = gX * ws, gYw = gY * ws, gZw = gZ * ws;
      double hbX[6], hbY[6], hbZ[6];
#pragma unroll
      for (int r = 0; r < 6; r++) {
        hbX[r] = gXw * __ldg(he + (6 * r + SEL[B][0]) * 32);
        hbY[r] = gYw * __ldg(he + (6 * r + SEL[B][1]) * 32);
        hbZ[r] = gZw * __ldg(he + (6 * r + SEL[B][2]) * 32);
      }
      double oX[3], oY[3], oZ[3];
      const int axp = c_tn.axp[par][t];
      if (axp < 0) {                      // the node is the apex: grad lambda_p = -(gX, gY, gZ)
#pragma unroll
        for (int a = 0; a < 3; a++) {
          oX[a] = -(gX * hbX[SEL[a][0]] + gY * hbX[SEL[a][1]] + gZ * hbX[SEL[a][2]]);
          oY[a] = -(gX * hbY[SEL[a][0]] + gY * hbY[SEL[a][1]] + gZ * hbY[SEL[a][2]]);
          oZ[a] = -(gX * hbZ[SEL[a][0]] + gY * hbZ[SEL[a][1]] + gZ * hbZ[SEL[a][2]]);
        }
      } else if (axp == 0) {
#pragma unroll
        for (int a = 0; a < 3; a++) { oX[a] = gX * hbX[SEL[a][0]]; oY[a] = gX * hbY[SEL[a][0]]; oZ[a] = gX * hbZ[SEL[a][0]]; }
      } else if (axp == 1) {
#pragma unroll
        for (int a = 0; a < 3; a++) { oX[a] = gY * hbX[SEL[a][1]]; oY[a] = gY * hbY[SEL[a][1]]; oZ[a] = gY * hbZ[SEL[a][1]]; }
      } else {
#pragma unroll
        for (int a = 0; a < 3; a++) { oX[a] = gZ * hbX[SEL[a][2]]; oY[a] = gZ * hbY[SEL[a][2]]; oZ[a] = gZ * hbZ[SEL[a][2]]; }
      }
      double* dX = acc + (int)c_tn.cslot[par][t][c_tn.qax[par][t][0]] * TN_SLOT + lane;
      double* dY = acc + (int)c_tn.cslot[par][t][c_tn.qax[par][t][1]] * TN_SLOT + lane;
      double* dZ = acc + (int)c_tn.cslot[par][t][c_tn.qax[par][t][2]] * TN_SLOT + lane;
      double* dA = acc + (int)c_tn.cslot[par][t][c_tn.qax[par][t][3]] * TN_SLOT + lane;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        dX[a * 32] += oX[a]; dY[a * 32] += oY[a]; dZ[a * 32] += oZ[a];
        dA[a * 32] -= (oX[a] + oY[a]) + oZ[a];
      }
      continue;
    }
    double H[6][6];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int cc = 0; cc < 3; cc++) H[r][COL[B][cc]] = __ldg(he + (6 * r + COL[B][cc]) * 32) * ws;
    const int p = c_tn.inc[par][t][4];
    const double gp[3] = {c_tn.gcoef[par][t][p][0] * ihx, c_tn.gcoef[par][t][p][1] * ihy, c_tn.gcoef[par][t][p][2] * ihz};
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const double gq[3] = {c_tn.gcoef[par][t][q][0] * ihx, c_tn.gcoef[par][t][q][1] * ihy, c_tn.gcoef[par][t][q][2] * ihz};
      double hb[6];
#pragma unroll
      for (int r = 0; r < 6; r++) hb[r] = bdot<3>(B, gq, H[r]);
      double* dst = acc + (int)c_tn.cslot[par][t][q] * TN_SLOT + lane;
#pragma unroll
      for (int a = 0; a < 3; a++) dst[a * 32] += bdot<3>(a, gp, hb);
    }
  }
}

// nodes [node0, node1) (flat ids; whole node planes when the forward runs in z-chunks on two streams).
// HS = warps per column component of the 32-tetrahedron parity.  HS = 1 (first version, 6 warps): its three warps walk all 32 incident tetrahedra while
// the three warps of the 8-tetrahedron parity wait at the barrier three quarters of the time (profiles/ncu_r02_cfg5_v3.md: 32 % of the warp
// samples).  HS = 2 (9 warps): warps 0-2 take tetrahedra [0, 16), warps 3-5 tetrahedra [16, 32) of the same nodes into accumulators of their own,
// summed when the rows are stored; the longest warp then runs 16 iterations instead of 32 (105 KB of accumulators, two CTAs per SM).
template <int HS>
static __global__ void __launch_bounds__((3 * HS + 3) * 32, HS == 1 ? 3 : 2) k_tet_node_fwd(GridTet gt, TetSpacing sp, long long nnz, long long node0, long long node1,
                                                                       const long long* __restrict__ rowptr, const double* __restrict__ hx, double* __restrict__ vals) {
  constexpr int NTH = (3 * HS + 3) * 32, HW = 3 * HS;              // threads, warps of the heavy parity
  extern __shared__ __align__(16) double tn_acc[];
  __shared__ long long s_rs[TN_NODES];
  __shared__ int s_mask[TN_NODES], s_src[TN_NODES];
  __shared__ unsigned char s_slot27[2][19];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, heavy = c_tn.heavy, par = warp < HW ? heavy : 1 - heavy;
  const int b = warp < HW ? warp % 3 : warp - HW, part = warp < HW ? warp / 3 : 0;
  const long long n1 = gt.n + 1, nn = node1, f0 = node0 + (long long)blockIdx.x * TN_NODES;
  // the lane's node: the one of the flat pair (f0 + 2*lane, f0 + 2*lane + 1) whose index parity (i + j + k) & 1 is `par`
  long long f = f0 + 2 * lane;
  {
    const int i0 = (int)(f % n1), j0 = (int)((f / n1) % n1), k0 = (int)(f / (n1 * n1));
    f += ((i0 + j0 + k0) & 1) ^ par;
  }
  const bool valid = f < nn;
  const int i = (int)(f % n1), j = (int)((f / n1) % n1), k = (int)(f / (n1 * n1));
  double* acc = tn_acc + (warp < HW ? warp * TN_ACC1 : HW * TN_ACC1 + b * TN_ACC0);
  const int nslot = warp < HW ? 19 : 7;
  if (tid < 38) s_slot27[tid / 19][tid % 19] = c_tn.slot27[tid / 19][tid % 19];
  for (int s = 0; s < nslot; s++) { acc[s * TN_SLOT + lane] = 0.0; acc[s * TN_SLOT + 32 + lane] = 0.0; acc[s * TN_SLOT + 64 + lane] = 0.0; }
  if (b == 0 && part == 0) {
    const int fpos = (int)(f - f0);
    int mask = 0;
    if (valid) {      // structurally present for the parity and inside the grid (27-neighbourhood id = (dk+1)*9 + (dj+1)*3 + (di+1))
      const int mx = (i > 0 ? 1 : 0) | 2 | (i < gt.n ? 4 : 0), my = (j > 0 ? 1 : 0) | 2 | (j < gt.n ? 4 : 0), mz = (k > 0 ? 1 : 0) | 2 | (k < gt.l ? 4 : 0);
      const int row9 = ((my & 1) ? mx : 0) | (mx << 3) | ((my & 4) ? mx << 6 : 0);
      mask = (((mz & 1) ? row9 : 0) | (row9 << 9) | ((mz & 4) ? row9 << 18 : 0)) & c_tn.present[par];
    }
    s_mask[fpos] = mask; s_src[fpos] = par * 32 + lane; s_rs[fpos] = valid ? rowptr[f] : 0;
  }
  // tetrahedra of this warp: the heavy parity's incidence list is cut into HS ascending pieces
  const int per = (c_tn.ninc[heavy] + HS - 1) / HS, t0 = warp < HW ? part * per : 0, t1 = warp < HW ? t0 + per : 64;
  if (b == 0) tn_accumulate<0>(gt, sp, par, valid, i, j, k, lane, hx, acc, t0, t1);
  else if (b == 1) tn_accumulate<1>(gt, sp, par, valid, i, j, k, lane, hx, acc, t0, t1);
  else tn_accumulate<2>(gt, sp, par, valid, i, j, k, lane, hx, acc, t0, t1);
  __syncthreads();
  // store: item (node position nd, column component bb, compact slot cs); the three row components a share the lookups
  for (int idx = tid; idx < TN_NODES * 57; idx += NTH) {
    const int nd = idx / 57, r = idx - 57 * nd, bb = r / 19, cs = r - 19 * bb;
    const int mask = s_mask[nd], src = s_src[nd], pn = src >> 5, ln = src & 31;
    if (cs >= (pn == heavy ? 19 : 7)) continue;
    const int s27 = s_slot27[pn][cs];
    if (!((mask >> s27) & 1)) continue;
    const int len = __popc((unsigned)mask), jpos = __popc((unsigned)(mask & ((1 << s27) - 1)));
    const double* a0 = tn_acc + (pn == heavy ? bb * TN_ACC1 : HW * TN_ACC1 + bb * TN_ACC0) + cs * TN_SLOT + ln;
    double* out = vals + 3 * s_rs[nd] + bb * len + jpos;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      double v = a0[a * 32];
      if (HS == 2 && pn == heavy) v += a0[3 * TN_ACC1 + a * 32];      // second piece of the incidence list (ascending element order inside each piece)
      out[3 * (long long)a * nnz] = v;
    }
  }
}
#endif

}  // namespace adfem
