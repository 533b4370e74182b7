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

constexpr int TILE_MAX_THREADS = 512;

struct DevPattern {
  int n; long long nnz;
  const long long* rowptr;        // n+1
  const int* colind;              // nnz
  const uint32_t* slot_nnz;       // [d*d][ne] struct-of-arrays
};
struct DevTiles {                  // forward (FwdTiles) or adjoint (AdjTiles) blobs on the device
  int ntiles, sym;
  unsigned max_blob;               // bytes, multiple of 16
  int max_elems, max_nnz;
  const long long* blob_ptr;
  const unsigned char* blob;
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
  Geom<DIM> G; load_geom(m, e, G);
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
  Geom<DIM> G; load_geom(m, e, G);
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
  Geom<3> G; load_geom(m, e, G);
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
  Geom<3> G; load_geom(m, e, G);
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
  Geom<DIM> G; load_geom(m, e, G);
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
  Geom<DIM> G; load_geom(m, e, G);
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
    Geom<DIM> G; load_geom(m, e, G);
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
  Geom<DIM> G; load_geom(m, e, G);
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
// ---- TMA bulk copy + mbarrier (sm_90+/sm_100a PTX) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one contiguous global -> shared bulk copy (SASS: UBLKCP), completion signalled on the mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ unsigned a16(unsigned x) { return (x + 15u) & ~15u; }
// asynchronous 8-byte global -> shared copy (SASS: LDGSTS)
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Local element matrix summed over Gauss points, handed to `put(slot, value)`.
//   scalar ops (symmetric): slot = index in the packed upper triangle (p <= q, row-major)
//   stiffness (H may be unsymmetric): slot = l*Dt + s; P2 accumulates over Gauss points through put(-slot-1, v)
template <int DIM, int DEG, int OP, typename Put>
__device__ __forceinline__ void local_matrix(const DevMesh& m, const Geom<DIM>& G, int e, const double* __restrict__ coef, Put put) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  if (OP == OP_LAPLACE && DEG == 1) {
    double c = 0.0;                                   // gradients are constant: sum the coefficients first
    for (int k = 0; k < m.g; k++) c += coef[(size_t)e * m.g + k] * (m.rule.w[k] * G.wscale);
    int i = 0;
#pragma unroll
    for (int p = 0; p < D; p++)
#pragma unroll
      for (int q = p; q < D; q++) put(i++, dotg<DIM>(G.gL[p], G.gL[q]) * c);
  } else if (OP == OP_LAPLACE || OP == OP_MASS) {
    constexpr int NA = D * (D + 1) / 2;
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
#pragma unroll
    for (int i = 0; i < NA; i++) put(i, acc[i]);
  } else {
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

// non-blocking probe of an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Forward: persistent CTAs walk the row tiles (tile = blockIdx.x + k*gridDim.x).  A tile's mesh-static blob (index
// lists + vertex coordinates) arrives by ONE TMA bulk copy; with nbuf = 2 the copy of the NEXT tile is issued before
// the current one is processed and, once it has landed, the coefficient lines of the next tile are prefetched into
// L2, so neither latency sits on the critical path.  Phase A evaluates the local matrices of every element touching
// the tile's rows into shared memory; phase B lets each CSR entry sum its contributions in a fixed (column, element)
// order and writes it once.
template <int DIM, int DEG, int OP>
__global__ void __launch_bounds__(TILE_MAX_THREADS) k_tile_fwd(DevMesh m, long long nnz_s, DevTiles tp, int nbuf, const double* __restrict__ coef,
                                                               double* __restrict__ vals) {
  constexpr int D = ElemTraits<DIM, DEG>::D, NC = OP == OP_STIFFNESS ? DIM : 1, Dt = NC * D, dd = D * D, NVL = DIM + 1;
  extern __shared__ __align__(128) unsigned char smem_all[];
  __shared__ __align__(8) uint64_t mbar[2];
  const int tid = threadIdx.x, nth = blockDim.x;
  double* loc = reinterpret_cast<double*>(smem_all + (size_t)nbuf * tp.max_blob);
  const int cpe = m.g * (OP == OP_STIFFNESS ? Voigt<DIM>::NS * Voigt<DIM>::NS : 1);      // coefficients per element
  if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); }
  __syncthreads();
  auto issue = [&](int t, int buf) {
    const long long b0 = tp.blob_ptr[t];
    const uint32_t bytes = (uint32_t)(tp.blob_ptr[t + 1] - b0);
    mbar_expect_tx(&mbar[buf], bytes);
    tma_bulk_g2s(smem_all + (size_t)buf * tp.max_blob, tp.blob + b0, bytes, &mbar[buf]);
  };
  if (nbuf == 2 && tid == 0 && (int)blockIdx.x < tp.ntiles) issue(blockIdx.x, 0);
  int it = 0;
  for (int t = blockIdx.x; t < tp.ntiles; t += gridDim.x, it++) {
  const int sbuf = nbuf == 2 ? (it & 1) : 0, tn = t + gridDim.x;
  const uint32_t parity = nbuf == 2 ? ((it >> 1) & 1) : (it & 1);
  if (tid == 0) { if (nbuf == 1) issue(t, 0); else if (tn < tp.ntiles) issue(tn, sbuf ^ 1); }
  unsigned char* smem = smem_all + (size_t)sbuf * tp.max_blob;
  mbar_wait(&mbar[sbuf], parity);
  const int* hdr = reinterpret_cast<const int*>(smem);
  const int nrows = hdr[0], nel = hdr[1], nvt = hdr[2], nnz_t = hdr[3], nsrc = hdr[4], ncls = hdr[5], ent32 = hdr[6];
  unsigned o = 32;
  const unsigned* rstart = reinterpret_cast<const unsigned*>(smem + o); o += a16(4u * nrows);
  const unsigned short* rlen = reinterpret_cast<const unsigned short*>(smem + o); o += a16(2u * nrows);
  const int* elems = reinterpret_cast<const int*>(smem + o); o += a16(4u * nel);
  const unsigned short* tv = reinterpret_cast<const unsigned short*>(smem + o); o += a16(2u * NVL * nel);
  const double* xy = reinterpret_cast<const double*>(smem + o); o += a16(8u * DIM * nvt);
  const int* cls = reinterpret_cast<const int*>(smem + o); o += a16(16u * ncls);
  const unsigned char* ent = smem + o; o += a16((ent32 ? 4u : 2u) * nnz_t);
  const unsigned short* src = reinterpret_cast<const unsigned short*>(smem + o);
  (void)nsrc; (void)rlen;

  if (nbuf == 1 || it == 0) {  // warm the coefficient lines of this thread's elements before the first use (non-blocking)
    for (int le = tid; le < nel; le += nth) {
      const double* p = coef + (size_t)elems[le] * cpe;
      for (int b = 0; b < cpe; b += 16) prefetch_l1(p + b);
      prefetch_l1(p + cpe - 1);
    }
  }
  for (int le = tid; le < nel; le += nth) {
    Geom<DIM> G; tile_geom(tv, xy, nel, le, m.heron, G);
    local_matrix<DIM, DEG, OP>(m, G, elems[le], coef, [&](int slot, double v) {
      if (slot >= 0) loc[slot * nel + le] = v; else loc[(-slot - 1) * nel + le] += v;
    });
  }
  if (nbuf == 2 && tn < tp.ntiles && mbar_test(&mbar[sbuf ^ 1], ((it + 1) >> 1) & 1)) {
    // the next tile's blob has landed: pull its coefficient lines into L2 while this tile finishes
    const unsigned char* nb = smem_all + (size_t)(sbuf ^ 1) * tp.max_blob;
    const int* nh = reinterpret_cast<const int*>(nb);
    const int* nel_ids = reinterpret_cast<const int*>(nb + 32 + a16(4u * nh[0]) + a16(2u * nh[0]));
    for (int le = tid; le < nh[1]; le += nth) {
      const double* p = coef + (size_t)nel_ids[le] * cpe;
      for (int b = 0; b < cpe; b += 4) prefetch_l2(p + b);
      prefetch_l2(p + cpe - 1);
    }
  }
  __syncthreads();
  for (int c = 0; c < ncls; c++) {
    const int cnt = cls[4 * c], n = cls[4 * c + 1];
    const unsigned short* sc = src + cls[4 * c + 2];
    const int e0 = cls[4 * c + 3];
    auto dest = [&](int i, int& lr, int& j) {
      if (ent32) { const unsigned v = reinterpret_cast<const unsigned*>(ent)[e0 + i]; lr = v & 0xffffu; j = v >> 16; }
      else { const unsigned v = reinterpret_cast<const unsigned short*>(ent)[e0 + i]; lr = v & 0xffu; j = v >> 8; }
    };
    if (NC == 1) {
      // every entry of the class has `cnt` sources: the same unrolled gather in every lane
#define ADFEM_GATHER_CLASS(C)                                                              \
      for (int i = tid; i < n; i += nth) {                                                 \
        double v = 0.0;                                                                    \
        _Pragma("unroll") for (int k = 0; k < C; k++) v += loc[sc[k * n + i]];           \
        int lr, j; dest(i, lr, j);                                                         \
        vals[(size_t)rstart[lr] + j] = v;                                                  \
      }
      switch (cnt) {
        case 1: ADFEM_GATHER_CLASS(1) break;
        case 2: ADFEM_GATHER_CLASS(2) break;
        case 3: ADFEM_GATHER_CLASS(3) break;
        case 4: ADFEM_GATHER_CLASS(4) break;
        case 5: ADFEM_GATHER_CLASS(5) break;
        case 6: ADFEM_GATHER_CLASS(6) break;
        case 7: ADFEM_GATHER_CLASS(7) break;
        case 8: ADFEM_GATHER_CLASS(8) break;
        default:
          for (int i = tid; i < n; i += nth) {
            double v = 0.0;
            for (int k = 0; k < cnt; k++) v += loc[sc[k * n + i]];
            int lr, j; dest(i, lr, j);
            vals[(size_t)rstart[lr] + j] = v;
          }
      }
#undef ADFEM_GATHER_CLASS
    } else {
      for (int i = tid; i < n; i += nth) {
        double v[NC * NC];
#pragma unroll
        for (int ab = 0; ab < NC * NC; ab++) v[ab] = 0.0;
        for (int k = 0; k < cnt; k++) {
          const int cc = sc[k * n + i], le = cc / dd, pq = cc - le * dd, p = pq / D, q = pq - p * D;
#pragma unroll
          for (int a = 0; a < NC; a++)
#pragma unroll
            for (int b = 0; b < NC; b++) v[a * NC + b] += loc[((a * D + p) * Dt + b * D + q) * nel + le];
        }
        int lr, j; dest(i, lr, j);
        const long long len = rlen[lr], rs = rstart[lr];
#pragma unroll
        for (int a = 0; a < NC; a++)
#pragma unroll
          for (int b = 0; b < NC; b++) vals[NC * (a * nnz_s + rs) + b * len + j] = v[a * NC + b];
      }
    }
  }
  __syncthreads();   // loc and this blob buffer are free again
  }
}

// Contract the upstream gradients `g(l, s)` of one element's local matrix with its shape tables: writes
// grad_coef for every Gauss point of the element.
template <int DIM, int DEG, int OP, typename Get>
__device__ __forceinline__ void local_adjoint(const DevMesh& m, const Geom<DIM>& G, int e, Get g, double* __restrict__ grad_coef) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
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
  Geom<DIM> G; load_geom(m, e, G);
  if (NC == 1) {
    local_adjoint<DIM, DEG, OP>(m, G, e, [&](int p, int q) { return dvals[pat.slot_nnz[(size_t)(p * D + q) * m.ne + e]]; }, grad_coef);
  } else {
    local_adjoint<DIM, DEG, OP>(m, G, e, [&](int l, int s) {
      const int a = l / D, p = l % D, b = s / D, q = s % D;
      const int r = ldg(m.conn + (size_t)p * m.ne + e);
      const long long rs = pat.rowptr[r], len = pat.rowptr[r + 1] - rs;
      const long long j = (long long)pat.slot_nnz[(size_t)(p * D + q) * m.ne + e] - rs;
      return dvals[NC * (a * pat.nnz + rs) + b * len + j];
    }, grad_coef);
  }
}

// Adjoint, tiled version: persistent CTAs walk the element tiles.  One TMA bulk copy brings a tile blob (double
// buffered: the next tile's copy is in flight while this one is processed, and once it has landed the CSR rows it
// will stage are prefetched into L2); the CSR rows the tile's elements touch are staged into shared memory with
// asynchronous copies; every element then gathers its upstream gradients from shared memory and contracts them
// with its shape tables.
template <int DIM, int DEG, int OP>
__global__ void __launch_bounds__(TILE_MAX_THREADS) k_tile_adj(DevMesh m, long long nnz_s, DevTiles ap, int nbuf, const double* __restrict__ dvals,
                                                               double* __restrict__ grad_coef) {
  constexpr int D = ElemTraits<DIM, DEG>::D, NC = OP == OP_STIFFNESS ? DIM : 1, dd = D * D, NVL = DIM + 1;
  extern __shared__ __align__(128) unsigned char smem_all[];
  __shared__ __align__(8) uint64_t mbar[2];
  const int tid = threadIdx.x, nth = blockDim.x;
  double* sd = reinterpret_cast<double*>(smem_all + (size_t)nbuf * ap.max_blob);       // NC*NC x nnz_t staged upstream gradients
  if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); }
  __syncthreads();
  auto issue = [&](int t, int buf) {
    const long long b0 = ap.blob_ptr[t];
    const uint32_t bytes = (uint32_t)(ap.blob_ptr[t + 1] - b0);
    mbar_expect_tx(&mbar[buf], bytes);
    tma_bulk_g2s(smem_all + (size_t)buf * ap.max_blob, ap.blob + b0, bytes, &mbar[buf]);
  };
  if (nbuf == 2 && tid == 0 && (int)blockIdx.x < ap.ntiles) issue(blockIdx.x, 0);
  int it = 0;
  for (int t = blockIdx.x; t < ap.ntiles; t += gridDim.x, it++) {
  const int sbuf = nbuf == 2 ? (it & 1) : 0, tn = t + gridDim.x;
  const uint32_t parity = nbuf == 2 ? ((it >> 1) & 1) : (it & 1);
  if (tid == 0) { if (nbuf == 1) issue(t, 0); else if (tn < ap.ntiles) issue(tn, sbuf ^ 1); }
  unsigned char* smem = smem_all + (size_t)sbuf * ap.max_blob;
  mbar_wait(&mbar[sbuf], parity);
  const int* hdr = reinterpret_cast<const int*>(smem);
  const int nrows = hdr[0], nel = hdr[1], nvt = hdr[2], nnz_t = hdr[3];
  unsigned o = 32;
  const unsigned* rstart = reinterpret_cast<const unsigned*>(smem + o); o += a16(4u * nrows);
  const unsigned short* roff = reinterpret_cast<const unsigned short*>(smem + o); o += a16(2u * (nrows + 1));
  const int* elems = reinterpret_cast<const int*>(smem + o); o += a16(4u * nel);
  const unsigned short* tv = reinterpret_cast<const unsigned short*>(smem + o); o += a16(2u * NVL * nel);
  const double* xy = reinterpret_cast<const double*>(smem + o); o += a16(8u * DIM * nvt);
  const unsigned short* lrow = reinterpret_cast<const unsigned short*>(smem + o); o += a16(2u * nnz_t);
  const unsigned short* gidx = reinterpret_cast<const unsigned short*>(smem + o);

  // stage the upstream gradients with asynchronous 8-byte copies (LDGSTS): every thread fires all of its
  // copies back to back, nothing waits on a register
  for (int i = tid; i < nnz_t; i += nth) {
    const int lr = lrow[i], j = i - roff[lr];
    if (NC == 1) cp_async8(sd + i, dvals + ((size_t)rstart[lr] + j));
    else {
      const long long len = roff[lr + 1] - roff[lr], rs = rstart[lr];
#pragma unroll
      for (int a = 0; a < NC; a++)
#pragma unroll
        for (int b = 0; b < NC; b++) cp_async8(sd + (a * NC + b) * nnz_t + i, dvals + (NC * (a * nnz_s + rs) + b * len + j));
    }
  }
  if (nbuf == 2 && tn < ap.ntiles && mbar_test(&mbar[sbuf ^ 1], ((it + 1) >> 1) & 1)) {
    // the next tile's blob has landed: pull the CSR rows it will stage into L2 (scalar layout; rows are <= a few sectors)
    const unsigned char* nb = smem_all + (size_t)(sbuf ^ 1) * ap.max_blob;
    const int* nh = reinterpret_cast<const int*>(nb);
    const unsigned* nrs = reinterpret_cast<const unsigned*>(nb + 32);
    const unsigned short* nro = reinterpret_cast<const unsigned short*>(nb + 32 + a16(4u * nh[0]));
    if (NC == 1)
      for (int lr = tid; lr < nh[0]; lr += nth) {
        const double* p = dvals + nrs[lr];
        const int len = nro[lr + 1] - nro[lr];
        for (int b = 0; b < len; b += 4) prefetch_l2(p + b);
        prefetch_l2(p + len - 1);
      }
  }
  cp_async_wait_all();
  __syncthreads();
  for (int le = tid; le < nel; le += nth) {
    Geom<DIM> G; tile_geom(tv, xy, nel, le, m.heron, G);
    const int e = elems[le];
    if (NC == 1) local_adjoint<DIM, DEG, OP>(m, G, e, [&](int p, int q) { return sd[gidx[(p * D + q) * nel + le]]; }, grad_coef);
    else local_adjoint<DIM, DEG, OP>(m, G, e, [&](int l, int s) {
      const int a = l / D, p = l % D, b = s / D, q = s % D;
      return sd[(a * NC + b) * nnz_t + gidx[(p * D + q) * nel + le]];
    }, grad_coef);
  }
  __syncthreads();   // sd and this blob buffer are free again
  }
}

}  // namespace adfem
