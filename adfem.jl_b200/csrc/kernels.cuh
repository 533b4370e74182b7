// Hand-written fp64 sm_100a kernels of the assembly path.
//
// Two output layouts:
//  * "COO-compatible": exactly the reference ops' outputs — one block of D*D values per Gauss point in
//    slot order ((e*g+k)*D+p)*D+q (deps/MFEM/FemLaplace1/FemLaplaceScalar.h:16-22) — kept because the
//    TF-op boundary and src/pcl.jl observe it.
//  * "CSR": values of the canonical CSR matrix the reference's callers build from that COO
//    (src/MFEM/MCore.jl:118-119).  This is the fast path: k_tile_fwd / k_tile_adj.
#pragma once
#include "device_fem.cuh"

namespace adfem {

constexpr int TILE_THREADS = 256;

struct DevPattern {
  int n; long long nnz;
  const long long* rowptr;        // n+1
  const int* colind;              // nnz
  const uint32_t* slot_nnz;       // [d*d][ne] struct-of-arrays
};
struct DevTilePlan {
  int ntiles, max_rows, max_elems, max_nnz;
  const int *row_ptr, *rows, *elem_ptr, *elems;
  const long long* soff_ptr; const uint16_t* src_off;
  const long long* src_ptr;  const uint16_t* src;
};
struct DevAdjPlan {
  int ntiles, max_rows, max_elems, max_nnz;
  const int *elem_ptr, *elems, *row_ptr, *rows;
  const long long* gidx_ptr; const uint16_t* gidx;
};

// ==================================================================================================
// COO-compatible kernels: one thread per Gauss point
// ==================================================================================================
// FemLaplaceScalar_forward / ComputeFemMassMatrix1_forward (+ 3-D FemLaplaceScalarT_forward)
template <int DIM, int DEG, int OP>
__global__ void k_coo_scalar_fwd(DevMesh m, const double* __restrict__ coef, double* __restrict__ vv) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)m.ne * m.g) return;
  const int e = (int)(t / m.g), k = (int)(t % m.g);
  Geom<DIM> G; load_geom<DIM>(m, e, G);
  double L[DIM + 1]; bary<DIM>(m.rule, k, L);
  const double c = coef[t], w = m.rule.w[k] * G.wscale;
  double* out = vv + t * (D * D);
  if (OP == OP_LAPLACE) {
    double gp[D][DIM]; basis_grad<DIM, DEG>(G, L, gp);
#pragma unroll
    for (int p = 0; p < D; p++)
#pragma unroll
      for (int q = 0; q < D; q++) out[p * D + q] = dotg<DIM>(gp[p], gp[q]) * c * w;
  } else {
    double phi[D]; basis_val<DIM, DEG>(L, phi);
#pragma unroll
    for (int p = 0; p < D; p++)
#pragma unroll
      for (int q = 0; q < D; q++) out[p * D + q] = phi[p] * phi[q] * c * w;
  }
}
// FemLaplaceScalar_backward / ComputeFemMassMatrix1_backward
template <int DIM, int DEG, int OP>
__global__ void k_coo_scalar_bwd(DevMesh m, const double* __restrict__ grad_vv, double* __restrict__ grad_coef) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)m.ne * m.g) return;
  const int e = (int)(t / m.g), k = (int)(t % m.g);
  Geom<DIM> G; load_geom<DIM>(m, e, G);
  double L[DIM + 1]; bary<DIM>(m.rule, k, L);
  const double w = m.rule.w[k] * G.wscale;
  const double* gin = grad_vv + t * (D * D);
  double v = 0.0;
  if (OP == OP_LAPLACE) {
    double gp[D][DIM]; basis_grad<DIM, DEG>(G, L, gp);
#pragma unroll
    for (int p = 0; p < D; p++)
#pragma unroll
      for (int q = 0; q < D; q++) v += gin[p * D + q] * (dotg<DIM>(gp[p], gp[q]) * w);
  } else {
    double phi[D]; basis_val<DIM, DEG>(L, phi);
#pragma unroll
    for (int p = 0; p < D; p++)
#pragma unroll
      for (int q = 0; q < D; q++) v += gin[p * D + q] * (phi[p] * phi[q] * w);
  }
  grad_coef[t] = v;
}
// 3-D mass in the reference's own layout: ONE slot per (e,p,q), summed over Gauss points
// (deps/MFEM3/ComputeFemMassMatrixMfem3/ComputeFemMassMatrixMfemT.h:4-27). One thread per (e,p).
template <int DEG>
__global__ void k_coo_mass3_fwd(DevMesh m, const double* __restrict__ rho, double* __restrict__ vv) {
  constexpr int D = ElemTraits<3, DEG>::D;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)m.ne * D) return;
  const int e = (int)(t / D), p = (int)(t % D);
  Geom<3> G; load_geom<3>(m, e, G);
  double acc[D];
#pragma unroll
  for (int q = 0; q < D; q++) acc[q] = 0.0;
  for (int k = 0; k < m.g; k++) {
    double L[4]; bary<3>(m.rule, k, L);
    double phi[D]; basis_val<3, DEG>(L, phi);
    double php = 0.0;
#pragma unroll
    for (int q = 0; q < D; q++) php = (q == p) ? phi[q] : php;
    const double w = m.rule.w[k] * G.wscale, r = rho[(size_t)e * m.g + k];
#pragma unroll
    for (int q = 0; q < D; q++) acc[q] += php * phi[q] * w * r;
  }
#pragma unroll
  for (int q = 0; q < D; q++) vv[t * D + q] = acc[q];
}
// adjoint of the above (the reference's Grad op body is empty — extension Q5)
template <int DEG>
__global__ void k_coo_mass3_bwd(DevMesh m, const double* __restrict__ grad_vv, double* __restrict__ grad_rho) {
  constexpr int D = ElemTraits<3, DEG>::D;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)m.ne * m.g) return;
  const int e = (int)(t / m.g), k = (int)(t % m.g);
  Geom<3> G; load_geom<3>(m, e, G);
  double L[4]; bary<3>(m.rule, k, L);
  double phi[D]; basis_val<3, DEG>(L, phi);
  const double w = m.rule.w[k] * G.wscale;
  const double* gin = grad_vv + (size_t)e * D * D;
  double v = 0.0;
#pragma unroll
  for (int p = 0; p < D; p++)
#pragma unroll
    for (int q = 0; q < D; q++) v += phi[p] * phi[q] * w * gin[p * D + q];
  grad_rho[t] = v;
}
// ComputeFemStiffnessMatrixMfem_forward (2-D) and its 3-D extension: NN = B^T H B w
template <int DIM, int DEG>
__global__ void k_coo_stiff_fwd(DevMesh m, const double* __restrict__ hmat, double* __restrict__ vv) {
  constexpr int D = ElemTraits<DIM, DEG>::D, NS = Voigt<DIM>::NS, Dt = DIM * D;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)m.ne * m.g) return;
  const int e = (int)(t / m.g), k = (int)(t % m.g);
  Geom<DIM> G; load_geom<DIM>(m, e, G);
  double L[DIM + 1]; bary<DIM>(m.rule, k, L);
  double gp[D][DIM]; basis_grad<DIM, DEG>(G, L, gp);
  const double w = m.rule.w[k] * G.wscale;
  double H[NS * NS];
#pragma unroll
  for (int i = 0; i < NS * NS; i++) H[i] = hmat[t * (NS * NS) + i];
  double* out = vv + t * (Dt * Dt);
#pragma unroll
  for (int cs = 0; cs < DIM; cs++)
#pragma unroll
    for (int ps = 0; ps < D; ps++) {
      double hb[NS];
#pragma unroll
      for (int i = 0; i < NS; i++) hb[i] = bdot<DIM>(cs, gp[ps], &H[i * NS]);
#pragma unroll
      for (int cl = 0; cl < DIM; cl++)
#pragma unroll
        for (int pl = 0; pl < D; pl++) out[(cl * D + pl) * Dt + cs * D + ps] = bdot<DIM>(cl, gp[pl], hb) * w;
    }
}
// ComputeFemStiffnessMatrixMfem_backward: grad_H = B dK B^T w
template <int DIM, int DEG>
__global__ void k_coo_stiff_bwd(DevMesh m, const double* __restrict__ grad_vv, double* __restrict__ grad_hmat) {
  constexpr int D = ElemTraits<DIM, DEG>::D, NS = Voigt<DIM>::NS, Dt = DIM * D;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)m.ne * m.g) return;
  const int e = (int)(t / m.g), k = (int)(t % m.g);
  Geom<DIM> G; load_geom<DIM>(m, e, G);
  double L[DIM + 1]; bary<DIM>(m.rule, k, L);
  double gp[D][DIM]; basis_grad<DIM, DEG>(G, L, gp);
  const double w = m.rule.w[k] * G.wscale;
  const double* gin = grad_vv + t * (Dt * Dt);
  double gH[NS * NS];
#pragma unroll
  for (int i = 0; i < NS * NS; i++) gH[i] = 0.0;
#pragma unroll
  for (int cl = 0; cl < DIM; cl++)
#pragma unroll
    for (int pl = 0; pl < D; pl++) {
      double tl[NS];
#pragma unroll
      for (int i = 0; i < NS; i++) tl[i] = 0.0;
#pragma unroll
      for (int cs = 0; cs < DIM; cs++)
#pragma unroll
        for (int ps = 0; ps < D; ps++) badd<DIM>(cs, gp[ps], gin[(cl * D + pl) * Dt + cs * D + ps], tl);
      double bl[NS];
#pragma unroll
      for (int i = 0; i < NS; i++) bl[i] = 0.0;
      badd<DIM>(cl, gp[pl], 1.0, bl);
#pragma unroll
      for (int i = 0; i < NS; i++)
#pragma unroll
        for (int j = 0; j < NS; j++) gH[i * NS + j] += bl[i] * tl[j];
    }
#pragma unroll
  for (int i = 0; i < NS * NS; i++) grad_hmat[t * (NS * NS) + i] = gH[i] * w;
}
// mesh-static COO indices, 0-based interleaved (row, col) int64 pairs; NC = 1 (scalar) or DIM (elasticity)
__global__ void k_coo_indices(DevMesh m, int nc, int per_gauss, long long* __restrict__ indices) {
  const int d = m.d, Dt = nc * d;
  const long long nblk = per_gauss ? (long long)m.ne * m.g : (long long)m.ne;
  const long long total = nblk * Dt * Dt;
  for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < total; s += (long long)gridDim.x * blockDim.x) {
    const long long blk = s / (Dt * Dt);
    const int ls = (int)(s % (Dt * Dt)), l = ls / Dt, c = ls % Dt;
    const int e = (int)(per_gauss ? blk / m.g : blk);
    indices[2 * s] = ldg(m.conn + (size_t)(l % d) * m.ne + e) + (long long)(l / d) * m.ndof;
    indices[2 * s + 1] = ldg(m.conn + (size_t)(c % d) * m.ne + e) + (long long)(c / d) * m.ndof;
  }
}

// ==================================================================================================
// Source term
// ==================================================================================================
// FemSourceScalar_forward: rhs[dof] = sum over incident (e,p) of sum_k f[e,k] phi_p(k) w_k.  One thread
// per dof row walking the dof -> (element, local) adjacency in element order: no atomics, and rhs needs
// no pre-zeroing (the reference requires a zeroed rhs, deps/MFEM/FemSource1/FemSourceScalar.cpp:73).
template <int DIM, int DEG>
__global__ void k_source_fwd(DevMesh m, const long long* __restrict__ adj_ptr, const int* __restrict__ adj_elem,
                             const uint8_t* __restrict__ adj_loc, const double* __restrict__ f, double* __restrict__ rhs) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m.ndof) return;
  double acc = 0.0;
  for (long long a = adj_ptr[r]; a < adj_ptr[r + 1]; a++) {
    const int e = adj_elem[a], p = adj_loc[a];
    Geom<DIM> G; load_geom<DIM>(m, e, G);
    for (int k = 0; k < m.g; k++) {
      double L[DIM + 1]; bary<DIM>(m.rule, k, L);
      double phi[D]; basis_val<DIM, DEG>(L, phi);
      double php = 0.0;
#pragma unroll
      for (int q = 0; q < D; q++) php = (q == p) ? phi[q] : php;
      acc += f[(size_t)e * m.g + k] * php * (m.rule.w[k] * G.wscale);
    }
  }
  rhs[r] = acc;
}
// FemSourceScalar_backward: pure gather
template <int DIM, int DEG>
__global__ void k_source_bwd(DevMesh m, const double* __restrict__ grad_rhs, double* __restrict__ grad_f) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)m.ne * m.g) return;
  const int e = (int)(t / m.g), k = (int)(t % m.g);
  Geom<DIM> G; load_geom<DIM>(m, e, G);
  double L[DIM + 1]; bary<DIM>(m.rule, k, L);
  double phi[D]; basis_val<DIM, DEG>(L, phi);
  const double w = m.rule.w[k] * G.wscale;
  double v = 0.0;
#pragma unroll
  for (int r = 0; r < D; r++) v += phi[r] * w * grad_rhs[ldg(m.conn + (size_t)r * m.ne + e)];
  grad_f[t] = v;
}

// ==================================================================================================
// CSR fast path
// ==================================================================================================
// Exclusive scan of n ints in shared memory by the whole CTA; out[n] = total. tmp = 32 ints of smem.
__device__ __forceinline__ void block_exclusive_scan(const int* in, int* out, int n, int* tmp) {
  const int nth = blockDim.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int ipt = (n + nth - 1) / nth, b = min(n, tid * ipt), e = min(n, b + ipt);
  int s = 0;
  for (int i = b; i < e; i++) s += in[i];
  int x = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  if (lane == 31) tmp[w] = x;
  __syncthreads();
  if (w == 0) {
    int v = lane < (nth >> 5) ? tmp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += y; }
    tmp[lane] = v;
  }
  __syncthreads();
  int base = (w > 0 ? tmp[w - 1] : 0) + x - s;
  for (int i = b; i < e; i++) { out[i] = base; base += in[i]; }
  if (tid == 0) out[n] = tmp[(nth >> 5) - 1];
  __syncthreads();
}

// Local element matrix summed over Gauss points, handed to `put(slot, value)` with slot = l*Dt + s.
template <int DIM, int DEG, int OP, typename Put>
__device__ __forceinline__ void local_matrix(const DevMesh& m, int e, const double* __restrict__ coef, Put put) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  Geom<DIM> G; load_geom<DIM>(m, e, G);
  if (OP == OP_LAPLACE && DEG == 1) {
    double c = 0.0;                                   // gradients are constant: sum the coefficients first
    for (int k = 0; k < m.g; k++) c += coef[(size_t)e * m.g + k] * (m.rule.w[k] * G.wscale);
#pragma unroll
    for (int p = 0; p < D; p++)
#pragma unroll
      for (int q = p; q < D; q++) { const double v = dotg<DIM>(G.gL[p], G.gL[q]) * c; put(p * D + q, v); if (q != p) put(q * D + p, v); }
  } else if (OP == OP_LAPLACE || OP == OP_MASS) {
    constexpr int NA = D * (D + 1) / 2;                // symmetric accumulators
    double acc[NA];
#pragma unroll
    for (int i = 0; i < NA; i++) acc[i] = 0.0;
    for (int k = 0; k < m.g; k++) {
      double L[DIM + 1]; bary<DIM>(m.rule, k, L);
      const double c = coef[(size_t)e * m.g + k] * (m.rule.w[k] * G.wscale);
      if (OP == OP_LAPLACE) {
        double gp[D][DIM]; basis_grad<DIM, DEG>(G, L, gp);
        int i = 0;
#pragma unroll
        for (int p = 0; p < D; p++)
#pragma unroll
          for (int q = p; q < D; q++) acc[i++] += dotg<DIM>(gp[p], gp[q]) * c;
      } else {
        double phi[D]; basis_val<DIM, DEG>(L, phi);
        int i = 0;
#pragma unroll
        for (int p = 0; p < D; p++)
#pragma unroll
          for (int q = p; q < D; q++) acc[i++] += phi[p] * phi[q] * c;
      }
    }
    int i = 0;
#pragma unroll
    for (int p = 0; p < D; p++)
#pragma unroll
      for (int q = p; q < D; q++) { put(p * D + q, acc[i]); if (q != p) put(q * D + p, acc[i]); i++; }
  } else {   // OP_STIFFNESS (H may be unsymmetric: keep the full block)
    constexpr int NS = Voigt<DIM>::NS, Dt = DIM * D;
    if (DEG == 1) {
      double H[NS * NS];                               // constant B: sum H_k w_k first
#pragma unroll
      for (int i = 0; i < NS * NS; i++) H[i] = 0.0;
      for (int k = 0; k < m.g; k++) {
        const double w = m.rule.w[k] * G.wscale;
        const double* hk = coef + ((size_t)e * m.g + k) * (NS * NS);
#pragma unroll
        for (int i = 0; i < NS * NS; i++) H[i] += hk[i] * w;
      }
#pragma unroll
      for (int cs = 0; cs < DIM; cs++)
#pragma unroll
        for (int ps = 0; ps < D; ps++) {
          double hb[NS];
#pragma unroll
          for (int i = 0; i < NS; i++) hb[i] = bdot<DIM>(cs, G.gL[ps], &H[i * NS]);
#pragma unroll
          for (int cl = 0; cl < DIM; cl++)
#pragma unroll
            for (int pl = 0; pl < D; pl++) put((cl * D + pl) * Dt + cs * D + ps, bdot<DIM>(cl, G.gL[pl], hb));
        }
    } else {
      // P2: Dt*Dt accumulators do not fit in registers; the caller's put() must accumulate
      // (first Gauss point stores, later ones add) — signalled through negative slot offset.
      for (int k = 0; k < m.g; k++) {
        double L[DIM + 1]; bary<DIM>(m.rule, k, L);
        double gp[D][DIM]; basis_grad<DIM, DEG>(G, L, gp);
        const double w = m.rule.w[k] * G.wscale;
        const double* hk = coef + ((size_t)e * m.g + k) * (NS * NS);
        double H[NS * NS];
#pragma unroll
        for (int i = 0; i < NS * NS; i++) H[i] = hk[i] * w;
        for (int cs = 0; cs < DIM; cs++)
          for (int ps = 0; ps < D; ps++) {
            double hb[NS];
#pragma unroll
            for (int i = 0; i < NS; i++) hb[i] = bdot<DIM>(cs, gp[ps], &H[i * NS]);
            for (int cl = 0; cl < DIM; cl++)
              for (int pl = 0; pl < D; pl++) {
                const int slot = (cl * D + pl) * Dt + cs * D + ps;
                put(k == 0 ? slot : -slot - 1, bdot<DIM>(cl, gp[pl], hb));
              }
          }
      }
    }
  }
}

// Forward: one CTA per row tile.  Phase A evaluates the local matrices of every element touching the
// tile's rows into shared memory; phase B lets each CSR entry of those rows sum its contributions in a
// fixed (column, element) order and writes it once.
template <int DIM, int DEG, int OP>
__global__ void __launch_bounds__(TILE_THREADS) k_tile_fwd(DevMesh m, DevPattern pat, DevTilePlan tp, const double* __restrict__ coef,
                                                           double* __restrict__ vals) {
  constexpr int D = ElemTraits<DIM, DEG>::D, NC = OP == OP_STIFFNESS ? DIM : 1, Dt = NC * D, S = Dt * Dt, dd = D * D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int scan_tmp[32];
  const int t = blockIdx.x, tid = threadIdx.x;
  const int r0 = tp.row_ptr[t], nrows = tp.row_ptr[t + 1] - r0;
  const int e0 = tp.elem_ptr[t], nel = tp.elem_ptr[t + 1] - e0;
  const long long so0 = tp.soff_ptr[t];
  const int nnz_t = (int)(tp.soff_ptr[t + 1] - so0) - 1;
  const uint16_t* __restrict__ soff = tp.src_off + so0;
  const uint16_t* __restrict__ src = tp.src + tp.src_ptr[t];
  double* loc = reinterpret_cast<double*>(smem_raw);                                   // S x nel
  long long* rstart = reinterpret_cast<long long*>(loc + (size_t)S * tp.max_elems);    // max_rows
  int* roff = reinterpret_cast<int*>(rstart + tp.max_rows);                            // max_rows + 1
  int* rlen = roff + tp.max_rows + 1;                                                  // max_rows
  unsigned short* lrow = reinterpret_cast<unsigned short*>(rlen + tp.max_rows);        // max_nnz

  for (int lr = tid; lr < nrows; lr += TILE_THREADS) {
    const int r = tp.rows[r0 + lr];
    const long long a = pat.rowptr[r], b = pat.rowptr[r + 1];
    rstart[lr] = a; rlen[lr] = (int)(b - a);
  }
  for (int le = tid; le < nel; le += TILE_THREADS) {
    const int e = tp.elems[e0 + le];
    local_matrix<DIM, DEG, OP>(m, e, coef, [&](int slot, double v) {
      if (slot >= 0) loc[(size_t)slot * nel + le] = v; else loc[(size_t)(-slot - 1) * nel + le] += v;
    });
  }
  __syncthreads();
  block_exclusive_scan(rlen, roff, nrows, scan_tmp);
  for (int lr = tid; lr < nrows; lr += TILE_THREADS) {
    const int o = roff[lr], n = rlen[lr];
    for (int j = 0; j < n; j++) lrow[o + j] = (unsigned short)lr;
  }
  __syncthreads();
  for (int i = tid; i < nnz_t; i += TILE_THREADS) {
    const int lr = lrow[i], j = i - roff[lr];
    const int sb = soff[i], se = soff[i + 1];
    if (NC == 1) {
      double v = 0.0;
      for (int s = sb; s < se; s++) { const int c = src[s]; v += loc[(size_t)(c % dd) * nel + c / dd]; }
      vals[rstart[lr] + j] = v;
    } else {
      double v[NC * NC];
#pragma unroll
      for (int ab = 0; ab < NC * NC; ab++) v[ab] = 0.0;
      for (int s = sb; s < se; s++) {
        const int c = src[s], le = c / dd, pq = c % dd, p = pq / D, q = pq % D;
#pragma unroll
        for (int a = 0; a < NC; a++)
#pragma unroll
          for (int b = 0; b < NC; b++) v[a * NC + b] += loc[(size_t)((a * D + p) * Dt + b * D + q) * nel + le];
      }
      const long long len = rlen[lr];
#pragma unroll
      for (int a = 0; a < NC; a++)
#pragma unroll
        for (int b = 0; b < NC; b++) vals[NC * (a * pat.nnz + rstart[lr]) + b * len + j] = v[a * NC + b];
    }
  }
}

// Contract the D*D (or Dt*Dt) upstream gradients `g(l, s)` of one element with its shape tables:
// writes grad_coef for every Gauss point of the element.
template <int DIM, int DEG, int OP, typename Get>
__device__ __forceinline__ void local_adjoint(const DevMesh& m, int e, Get g, double* __restrict__ grad_coef) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  Geom<DIM> G; load_geom<DIM>(m, e, G);
  if (OP == OP_LAPLACE && DEG == 1) {
    double s = 0.0;
#pragma unroll
    for (int p = 0; p < D; p++)
#pragma unroll
      for (int q = 0; q < D; q++) s += g(p, q) * dotg<DIM>(G.gL[p], G.gL[q]);
    for (int k = 0; k < m.g; k++) grad_coef[(size_t)e * m.g + k] = s * (m.rule.w[k] * G.wscale);
  } else if (OP == OP_LAPLACE || OP == OP_MASS) {
    constexpr int NA = D * (D + 1) / 2;     // only the symmetric part of g matters
    double gs[NA];
    { int i = 0;
#pragma unroll
      for (int p = 0; p < D; p++)
#pragma unroll
        for (int q = p; q < D; q++) gs[i++] = (q == p) ? g(p, p) : g(p, q) + g(q, p); }
    for (int k = 0; k < m.g; k++) {
      double L[DIM + 1]; bary<DIM>(m.rule, k, L);
      double v = 0.0;
      if (OP == OP_LAPLACE) {
        double gp[D][DIM]; basis_grad<DIM, DEG>(G, L, gp);
        int i = 0;
#pragma unroll
        for (int p = 0; p < D; p++)
#pragma unroll
          for (int q = p; q < D; q++) v += gs[i++] * dotg<DIM>(gp[p], gp[q]);
      } else {
        double phi[D]; basis_val<DIM, DEG>(L, phi);
        int i = 0;
#pragma unroll
        for (int p = 0; p < D; p++)
#pragma unroll
          for (int q = p; q < D; q++) v += gs[i++] * phi[p] * phi[q];
      }
      grad_coef[(size_t)e * m.g + k] = v * (m.rule.w[k] * G.wscale);
    }
  } else {
    constexpr int NS = Voigt<DIM>::NS;
    for (int k = 0; k < m.g; k++) {
      double L[DIM + 1]; bary<DIM>(m.rule, k, L);
      double gp[D][DIM]; basis_grad<DIM, DEG>(G, L, gp);
      double gH[NS * NS];
#pragma unroll
      for (int i = 0; i < NS * NS; i++) gH[i] = 0.0;
      for (int cl = 0; cl < DIM; cl++)
        for (int pl = 0; pl < D; pl++) {
          double tl[NS];
#pragma unroll
          for (int i = 0; i < NS; i++) tl[i] = 0.0;
          for (int cs = 0; cs < DIM; cs++)
            for (int ps = 0; ps < D; ps++) badd<DIM>(cs, gp[ps], g(cl * D + pl, cs * D + ps), tl);
          double bl[NS];
#pragma unroll
          for (int i = 0; i < NS; i++) bl[i] = 0.0;
          badd<DIM>(cl, gp[pl], 1.0, bl);
#pragma unroll
          for (int i = 0; i < NS; i++)
#pragma unroll
            for (int j = 0; j < NS; j++) gH[i * NS + j] += bl[i] * tl[j];
        }
      const double w = m.rule.w[k] * G.wscale;
      double* out = grad_coef + ((size_t)e * m.g + k) * (NS * NS);
#pragma unroll
      for (int i = 0; i < NS * NS; i++) out[i] = gH[i] * w;
      if (DEG == 1) {   // constant B: every Gauss point gets the same matrix up to its weight
        for (int k2 = 1; k2 < m.g; k2++) {
          const double w2 = m.rule.w[k2] * G.wscale;
          double* o2 = grad_coef + ((size_t)e * m.g + k2) * (NS * NS);
#pragma unroll
          for (int i = 0; i < NS * NS; i++) o2[i] = gH[i] * w2;
        }
        break;
      }
    }
  }
}

// Adjoint, direct version: one thread per element gathers dK through the slot -> nnz map.
template <int DIM, int DEG, int OP>
__global__ void k_csr_adj_gather(DevMesh m, DevPattern pat, const double* __restrict__ dvals, double* __restrict__ grad_coef) {
  constexpr int D = ElemTraits<DIM, DEG>::D, NC = OP == OP_STIFFNESS ? DIM : 1;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.ne) return;
  if (NC == 1) {
    local_adjoint<DIM, DEG, OP>(m, e, [&](int p, int q) { return dvals[pat.slot_nnz[(size_t)(p * D + q) * m.ne + e]]; }, grad_coef);
  } else {
    local_adjoint<DIM, DEG, OP>(m, e, [&](int l, int s) {
      const int a = l / D, p = l % D, b = s / D, q = s % D;
      const int r = ldg(m.conn + (size_t)p * m.ne + e);
      const long long rs = pat.rowptr[r], len = pat.rowptr[r + 1] - rs;
      const long long j = (long long)pat.slot_nnz[(size_t)(p * D + q) * m.ne + e] - rs;
      return dvals[NC * (a * pat.nnz + rs) + b * len + j];
    }, grad_coef);
  }
}

// Adjoint, tiled version: one CTA per element tile stages the CSR rows its elements touch into shared
// memory with coalesced loads, then every element gathers its upstream gradients from there.
template <int DIM, int DEG, int OP>
__global__ void __launch_bounds__(TILE_THREADS) k_tile_adj(DevMesh m, DevPattern pat, DevAdjPlan ap, const double* __restrict__ dvals,
                                                           double* __restrict__ grad_coef) {
  constexpr int D = ElemTraits<DIM, DEG>::D, NC = OP == OP_STIFFNESS ? DIM : 1, dd = D * D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int scan_tmp[32];
  const int t = blockIdx.x, tid = threadIdx.x;
  const int r0 = ap.row_ptr[t], nrows = ap.row_ptr[t + 1] - r0;
  const int e0 = ap.elem_ptr[t], nel = ap.elem_ptr[t + 1] - e0;
  const uint16_t* __restrict__ gidx = ap.gidx + ap.gidx_ptr[t];
  double* sd = reinterpret_cast<double*>(smem_raw);                                             // NC*NC x max_nnz
  long long* rstart = reinterpret_cast<long long*>(sd + (size_t)NC * NC * ap.max_nnz);          // max_rows
  int* roff = reinterpret_cast<int*>(rstart + ap.max_rows);                                     // max_rows + 1
  int* rlen = roff + ap.max_rows + 1;                                                           // max_rows
  unsigned short* lrow = reinterpret_cast<unsigned short*>(rlen + ap.max_rows);                 // max_nnz
  for (int lr = tid; lr < nrows; lr += TILE_THREADS) {
    const int r = ap.rows[r0 + lr];
    const long long a = pat.rowptr[r], b = pat.rowptr[r + 1];
    rstart[lr] = a; rlen[lr] = (int)(b - a);
  }
  __syncthreads();
  block_exclusive_scan(rlen, roff, nrows, scan_tmp);
  const int nnz_t = roff[nrows];
  for (int lr = tid; lr < nrows; lr += TILE_THREADS) {
    const int o = roff[lr], n = rlen[lr];
    for (int j = 0; j < n; j++) lrow[o + j] = (unsigned short)lr;
  }
  __syncthreads();
  for (int i = tid; i < nnz_t; i += TILE_THREADS) {
    const int lr = lrow[i], j = i - roff[lr];
    if (NC == 1) sd[i] = dvals[rstart[lr] + j];
    else {
      const long long len = rlen[lr];
#pragma unroll
      for (int a = 0; a < NC; a++)
#pragma unroll
        for (int b = 0; b < NC; b++) sd[(size_t)(a * NC + b) * nnz_t + i] = dvals[NC * (a * pat.nnz + rstart[lr]) + b * len + j];
    }
  }
  __syncthreads();
  for (int le = tid; le < nel; le += TILE_THREADS) {
    const int e = ap.elems[e0 + le];
    const uint16_t* __restrict__ gi = gidx + (size_t)le * dd;
    if (NC == 1) local_adjoint<DIM, DEG, OP>(m, e, [&](int p, int q) { return sd[gi[p * D + q]]; }, grad_coef);
    else local_adjoint<DIM, DEG, OP>(m, e, [&](int l, int s) {
      const int a = l / D, p = l % D, b = s / D, q = s % D;
      return sd[(size_t)(a * NC + b) * nnz_t + gi[p * D + q]];
    }, grad_coef);
  }
}

}  // namespace adfem
