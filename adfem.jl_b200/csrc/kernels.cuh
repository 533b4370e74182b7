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
  unsigned max_head, max_body;     // bytes, multiples of 16
  int max_elems, max_nnz;
  const long long* blob_ptr;       // 2*ntiles+1: head offset, body offset per tile, then the end
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
// pcl_FemLaplaceScalar_Jacobian (deps/MFEM/FemLaplace1/FemLaplaceScalar.h:65-92): d vv[slot] / d kappa[t], a G x N column-major
// dense matrix whose only structural non-zeros are H[t + slot*G] = (grad phi_p . grad phi_q) w for the d*d slots of Gauss point t.
// One thread per Gauss point; the caller zero-fills H (as the reference's Julia wrapper does, src/pcl.jl:35-39).
template <int DIM, int DEG>
__global__ void k_pcl_laplace(DevMesh m, double* __restrict__ H) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  const long long G = (long long)m.ne * m.g, t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= G) return;
  const int e = (int)(t / m.g), k = (int)(t % m.g);
  Geom<DIM> Gm; load_geom(m, e, Gm);
  double L[DIM + 1]; bary<DIM>(m.rule, k, L);
  double gp[D][DIM]; basis_grad<DIM, DEG>(Gm, L, gp);
  const double w = m.rule.w[k] * Gm.wscale;
#pragma unroll
  for (int p = 0; p < D; p++)
#pragma unroll
    for (int q = 0; q < D; q++) H[t + (t * (D * D) + p * D + q) * G] = dotg<DIM>(gp[p], gp[q]) * w;
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
// FemSourceScalar_backward: pure gather.  One thread per ELEMENT: geometry, connectivity and the D upstream values are loaded once
// and reused for all Gauss points.
template <int DIM, int DEG>
__global__ void k_source_bwd(DevMesh m, const double* __restrict__ grad_rhs, double* __restrict__ grad_f) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m.ne) return;
  Geom<DIM> G; load_geom(m, e, G);
  double gr[D];
#pragma unroll
  for (int r = 0; r < D; r++) gr[r] = grad_rhs[ldg(m.conn + (size_t)r * m.ne + e)];
  for (int k = 0; k < m.g; k++) {
    double L[DIM + 1]; bary<DIM>(m.rule, k, L);
    double phi[D]; basis_val<DIM, DEG>(L, phi);
    const double w = m.rule.w[k] * G.wscale;
    double v = 0.0;
#pragma unroll
    for (int r = 0; r < D; r++) v += phi[r] * w * gr[r];
    grad_f[(size_t)e * m.g + k] = v;
  }
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

// loop over the Gauss points of an element; GMAX > 0 unrolls (k < g <= GMAX) so that per-point data can live in registers
template <int GMAX, class F> __device__ __forceinline__ void for_gauss(int g, F f) {
  if constexpr (GMAX > 0) {
#pragma unroll
    for (int k = 0; k < GMAX; k++) if (k < g) f(k);
  } else {
    for (int k = 0; k < g; k++) f(k);
  }
}

// Scalar operators (Laplace / mass): local matrix summed over Gauss points, packed upper triangle (p <= q, row-major)
// handed to `put(index, value)`.  `cf(k)` is the coefficient at Gauss point k of this element.
template <int DIM, int DEG, int OP, int GMAX, typename Coef, typename Put>
__device__ __forceinline__ void local_matrix_scalar(const DevMesh& m, const Geom<DIM>& G, Coef cf, Put put) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  if (OP == OP_LAPLACE && DEG == 1) {
    double c = 0.0;                                   // gradients are constant: sum the coefficients first
    for_gauss<GMAX>(m.g, [&](int k) { c += cf(k) * (m.rule.w[k] * G.wscale); });
    int i = 0;
#pragma unroll
    for (int p = 0; p < D; p++)
#pragma unroll
      for (int q = p; q < D; q++) put(i++, dotg<DIM>(G.gL[p], G.gL[q]) * c);
  } else {
    constexpr int NA = D * (D + 1) / 2;
    double acc[NA];
#pragma unroll
    for (int i = 0; i < NA; i++) acc[i] = 0.0;
    for_gauss<GMAX>(m.g, [&](int k) {
      double L[DIM + 1]; bary<DIM>(m.rule, k, L);
      const double c = cf(k) * (m.rule.w[k] * G.wscale);
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
    });
#pragma unroll
    for (int i = 0; i < NA; i++) put(i, acc[i]);
  }
}

// P1 elasticity: the (DIM*D)^2 local matrix B^T H B from the Gauss-summed, weighted coefficient matrix H (NS x NS, row-major);
// slot = l*Dt + s with l = c*D + p (component-blocked, deps/MFEM/ComputeFemStiffnessMatrixMfem/ComputeFemStiffnessMatrixMfem.h:18-34)
template <int DIM, typename Put>
__device__ __forceinline__ void stiffness_p1_blocks(const Geom<DIM>& G, const double* H, Put put) {
  constexpr int D = DIM + 1, NS = Voigt<DIM>::NS, Dt = DIM * D;
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
}

// Adjoint of the P1 elasticity local matrix B^T H B (constant B): gH = sum_{l,s} dK(l,s) col(l) (x) col(s), l = a*D+p, s = b*D+q,
// col(c, g) = column of B for component c (device_fem.cuh).  Unweighted: grad H_k = w_k gH.
template <int DIM, typename Get>
__device__ __forceinline__ void stiffness_p1_adjoint(const Geom<DIM>& G, Get g, double* gH) {
  constexpr int D = DIM + 1, NS = Voigt<DIM>::NS;
#pragma unroll
  for (int i = 0; i < NS * NS; i++) gH[i] = 0.0;
  for (int cl = 0; cl < DIM; cl++)
    for (int pl = 0; pl < D; pl++) {
      double tl[NS];
#pragma unroll
      for (int i = 0; i < NS; i++) tl[i] = 0.0;
      for (int cs = 0; cs < DIM; cs++)
        for (int ps = 0; ps < D; ps++) badd<DIM>(cs, G.gL[ps], g(cl * D + pl, cs * D + ps), tl);
      double bl[NS];
#pragma unroll
      for (int i = 0; i < NS; i++) bl[i] = 0.0;
      badd<DIM>(cl, G.gL[pl], 1.0, bl);
#pragma unroll
      for (int i = 0; i < NS; i++)
#pragma unroll
        for (int j = 0; j < NS; j++) gH[i * NS + j] += bl[i] * tl[j];
    }
}

// Local element matrix summed over Gauss points, handed to `put(slot, value)`.
//   scalar ops (symmetric): slot = index in the packed upper triangle (p <= q, row-major)
//   stiffness (H may be unsymmetric): slot = l*Dt + s; P2 accumulates over Gauss points through put(-slot-1, v)
template <int DIM, int DEG, int OP, typename Put>
__device__ __forceinline__ void local_matrix(const DevMesh& m, const Geom<DIM>& G, int e, const double* __restrict__ coef, Put put) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  if (OP == OP_LAPLACE || OP == OP_MASS) {
    const double* ce = coef + (size_t)e * m.g;
    local_matrix_scalar<DIM, DEG, OP, 0>(m, G, [&](int k) { return ce[k]; }, put);
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

// --------------------------------------------------------------------------------------------------
// P2 scalar operators in MOMENT FORM (round 2).  On an affine simplex grad phi_p = sum_a (d phi_p / d lambda_a) grad lambda_a, so with
// S_ab = grad lambda_a . grad lambda_b |det| the P2 Laplace matrix needs only NV(2NV+1) coefficient moments (21 on triangles, 36 on
// tetrahedra) against mesh-independent tables instead of D(D+1)/2 gradient products per Gauss point:
//   vertex p, vertex q        : S_pq MV(p,q)                                MV(p,q) = sum_k c_k w_k alpha_p alpha_q,  alpha_p = 4 L_p - 1
//   vertex p, edge (a,b)      : 4 (S_pb AV[p][a] + S_pa AV[p][b])           AV[p][i] = sum_k c_k w_k alpha_p L_i
//   edge (a,b), edge (c,d)    : 16 (S_bd BL(a,c) + S_bc BL(a,d) + S_ad BL(b,c) + S_ac BL(b,d)),   BL(i,m) = sum_k c_k w_k L_i L_m
// (552 -> ~200 fp64 operations per P2 triangle, 3300 -> ~600 per P2 tetrahedron; same sums, different association: agreement with the
// per-point form is ~1e-15 relative).  The mass matrix uses the table phi_p(k) phi_q(k) w_k the same way.  Tables are built by every CTA
// in its prologue from the quadrature rule into shared memory, layout tab[k * NM + m].
template <int DIM> struct P2M {
  static constexpr int NV = DIM + 1, NE = NV * (NV - 1) / 2, D = NV + NE, NS = NV * (NV + 1) / 2;
  static constexpr int NM = 2 * NS + NV * NV;        // Laplace moments: MV (NS) | AV (NV*NV) | BL (NS)
  static constexpr int NA = D * (D + 1) / 2;         // packed upper triangle of the local matrix
  __host__ __device__ static constexpr int sym(int i, int j) { return i <= j ? i * NV - i * (i - 1) / 2 + (j - i) : j * NV - j * (j - 1) / 2 + (i - j); }
};
template <int DIM, int OP> __host__ __device__ constexpr int p2_table_rows() { return OP == OP_LAPLACE ? P2M<DIM>::NM : P2M<DIM>::NA; }

template <int DIM, int OP> ADFEM_HD void p2_build_table(const QuadRule& rule, int g, int tid, int nth, double* tab) {
  using M = P2M<DIM>;
  constexpr int R = p2_table_rows<DIM, OP>();
  for (int idx = tid; idx < R * g; idx += nth) {
    const int k = idx / R, mm = idx - R * k;
    double L[DIM + 1]; bary<DIM>(rule, k, L);
    double v;
    if (OP == OP_LAPLACE) {
      int i = 0, j = 0, kind = 0, c = 0;                 // decode mm -> (kind, i, j)
      for (int a = 0; a < M::NV; a++) for (int b = a; b < M::NV; b++) { if (c == mm) { kind = 0; i = a; j = b; } c++; }
      for (int a = 0; a < M::NV; a++) for (int b = 0; b < M::NV; b++) { if (c == mm) { kind = 1; i = a; j = b; } c++; }
      for (int a = 0; a < M::NV; a++) for (int b = a; b < M::NV; b++) { if (c == mm) { kind = 2; i = a; j = b; } c++; }
      double Li = 0, Lj = 0;
      for (int a = 0; a <= DIM; a++) { Li = a == i ? L[a] : Li; Lj = a == j ? L[a] : Lj; }
      v = kind == 0 ? (4.0 * Li - 1.0) * (4.0 * Lj - 1.0) : (kind == 1 ? (4.0 * Li - 1.0) * Lj : Li * Lj);
    } else {
      double phi[M::D]; basis_val<DIM, 2>(L, phi);
      int c = 0; v = 0.0;
      for (int a = 0; a < M::D; a++) for (int b = a; b < M::D; b++) { if (c == mm) v = phi[a] * phi[b]; c++; }
    }
    tab[idx] = v * rule.w[k];
  }
}

// S_ab = grad lambda_a . grad lambda_b * wscale, packed symmetric (P2M::sym)
template <int DIM> __device__ __forceinline__ void p2_metric(const Geom<DIM>& G, double* S) {
  int i = 0;
#pragma unroll
  for (int a = 0; a <= DIM; a++)
#pragma unroll
    for (int b = a; b <= DIM; b++) S[i++] = dotg<DIM>(G.gL[a], G.gL[b]) * G.wscale;
}

// forward: packed upper triangle of the P2 local matrix (same slot order as local_matrix_scalar) from the staged coefficients cf(k)
template <int DIM, int OP, typename Coef, typename Put>
__device__ __forceinline__ void p2_local_matrix(int g, const Geom<DIM>& G, const double* __restrict__ tab, Coef cf, Put put) {
  using M = P2M<DIM>;
  constexpr int R = p2_table_rows<DIM, OP>(), NV = M::NV, NE = M::NE;
  double mom[R];
#pragma unroll
  for (int i = 0; i < R; i++) mom[i] = 0.0;
  for (int k = 0; k < g; k++) {
    const double c = cf(k);
    const double* t = tab + k * R;
#pragma unroll
    for (int i = 0; i < R; i++) mom[i] += c * t[i];
  }
  if (OP == OP_MASS) {
#pragma unroll
    for (int i = 0; i < R; i++) put(i, mom[i] * G.wscale);
    return;
  }
  double S[M::NS]; p2_metric<DIM>(G, S);
  const double* MV = mom; const double* AV = mom + M::NS; const double* BL = mom + M::NS + NV * NV;
  int i = 0;
#pragma unroll
  for (int p = 0; p < NV; p++) {
#pragma unroll
    for (int q = p; q < NV; q++) put(i++, S[M::sym(p, q)] * MV[M::sym(p, q)]);
#pragma unroll
    for (int j = 0; j < NE; j++) {
      int a, b; edge_ends<DIM>(j, a, b);
      put(i++, 4.0 * (S[M::sym(p, b)] * AV[p * NV + a] + S[M::sym(p, a)] * AV[p * NV + b]));
    }
  }
#pragma unroll
  for (int j = 0; j < NE; j++) {
    int a, b; edge_ends<DIM>(j, a, b);
#pragma unroll
    for (int j2 = j; j2 < NE; j2++) {
      int c, d; edge_ends<DIM>(j2, c, d);
      put(i++, 16.0 * (S[M::sym(b, d)] * BL[M::sym(a, c)] + S[M::sym(b, c)] * BL[M::sym(a, d)] + S[M::sym(a, d)] * BL[M::sym(b, c)] +
                       S[M::sym(a, c)] * BL[M::sym(b, d)]));
    }
  }
}

// adjoint: gs = packed, symmetrised upstream gradient of the local matrix; out(k, d loss / d c_k)
template <int DIM, int OP, typename Out>
__device__ __forceinline__ void p2_local_adjoint(int g, const Geom<DIM>& G, const double* __restrict__ tab, const double* gs, Out out) {
  using M = P2M<DIM>;
  constexpr int R = p2_table_rows<DIM, OP>(), NV = M::NV, NE = M::NE;
  double cm[R];
  if (OP == OP_MASS) {
#pragma unroll
    for (int i = 0; i < R; i++) cm[i] = gs[i] * G.wscale;
  } else {
#pragma unroll
    for (int i = 0; i < R; i++) cm[i] = 0.0;
    double S[M::NS]; p2_metric<DIM>(G, S);
    double* MV = cm; double* AV = cm + M::NS; double* BL = cm + M::NS + NV * NV;
    int i = 0;
#pragma unroll
    for (int p = 0; p < NV; p++) {
#pragma unroll
      for (int q = p; q < NV; q++) MV[M::sym(p, q)] += gs[i++] * S[M::sym(p, q)];
#pragma unroll
      for (int j = 0; j < NE; j++) {
        int a, b; edge_ends<DIM>(j, a, b);
        const double v = 4.0 * gs[i++];
        AV[p * NV + a] += v * S[M::sym(p, b)]; AV[p * NV + b] += v * S[M::sym(p, a)];
      }
    }
#pragma unroll
    for (int j = 0; j < NE; j++) {
      int a, b; edge_ends<DIM>(j, a, b);
#pragma unroll
      for (int j2 = j; j2 < NE; j2++) {
        int c, d; edge_ends<DIM>(j2, c, d);
        const double v = 16.0 * gs[i++];
        BL[M::sym(a, c)] += v * S[M::sym(b, d)]; BL[M::sym(a, d)] += v * S[M::sym(b, c)];
        BL[M::sym(b, c)] += v * S[M::sym(a, d)]; BL[M::sym(b, d)] += v * S[M::sym(a, c)];
      }
    }
  }
  for (int k = 0; k < g; k++) {
    const double* t = tab + k * R;
    double v = 0.0;
#pragma unroll
    for (int i = 0; i < R; i++) v += cm[i] * t[i];
    out(k, v);
  }
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Scalar operators: contract the upstream gradients `g(p, q)` of one element's local matrix with its shape tables and
// hand the gradient w.r.t. the coefficient at Gauss point k to `out(k, value)`.  GMAX as in for_gauss.
template <int DIM, int DEG, int OP, int GMAX, typename Get, typename Out>
__device__ __forceinline__ void local_adjoint_scalar(const DevMesh& m, const Geom<DIM>& G, Get g, Out out) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  if (OP == OP_LAPLACE && DEG == 1) {
    double s = 0.0;
#pragma unroll
    for (int p = 0; p < D; p++)
#pragma unroll
      for (int q = 0; q < D; q++) s += g(p, q) * dotg<DIM>(G.gL[p], G.gL[q]);
    for_gauss<GMAX>(m.g, [&](int k) { out(k, s * (m.rule.w[k] * G.wscale)); });
  } else {
    constexpr int NA = D * (D + 1) / 2;     // only the symmetric part of g matters
    double gs[NA];
    { int i = 0;
#pragma unroll
      for (int p = 0; p < D; p++)
#pragma unroll
        for (int q = p; q < D; q++) gs[i++] = (q == p) ? g(p, p) : g(p, q) + g(q, p); }
    for_gauss<GMAX>(m.g, [&](int k) {
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
      out(k, v * (m.rule.w[k] * G.wscale));
    });
  }
}

// Contract the upstream gradients `g(l, s)` of one element's local matrix with its shape tables: writes
// grad_coef for every Gauss point of the element.
template <int DIM, int DEG, int OP, typename Get>
__device__ __forceinline__ void local_adjoint(const DevMesh& m, const Geom<DIM>& G, int e, Get g, double* __restrict__ grad_coef) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  if (OP == OP_LAPLACE || OP == OP_MASS) {
    double* ge = grad_coef + (size_t)e * m.g;
    local_adjoint_scalar<DIM, DEG, OP, 0>(m, G, g, [&](int k, double v) { ge[k] = v; });
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


// --------------------------------------------------------------------------------------------------
// Tile kernels.  Persistent CTAs walk tiles t_i = blockIdx.x + i*gridDim.x (gridDim.x == ntiles gives one CTA per tile).
// Software pipeline, per CTA:  heads in a ring of 3 buffers, bodies in a ring of 2, each filled by ONE TMA bulk copy
// (UBLKCP) signalled on its own mbarrier.  While tile i is processed
//   * body(i+1) and head(i+2) are already in flight (requested when tile i-1 finished), and
//   * as soon as head(i+1) has landed the data-dependent loads of tile i+1 are started — coefficient loads into
//     registers (forward) or asynchronous copies of the CSR rows into shared memory (adjoint) — so they are in flight
//     during the whole second half / all of tile i.
// In steady state no thread waits on a global load.
// --------------------------------------------------------------------------------------------------
struct TileRing {
  unsigned char *heads, *bodies;
  uint64_t* mbar;                  // [0..2] heads, [3..4] bodies
  unsigned max_head, max_body;
  const long long* blob_ptr; const unsigned char* blob;
  int first, stride, count;        // tiles of this CTA: first + i*stride, i < count
  __device__ __forceinline__ unsigned char* head(int i) const { return heads + (size_t)(i % 3) * max_head; }
  __device__ __forceinline__ unsigned char* body(int i) const { return bodies + (size_t)(i & 1) * max_body; }
  __device__ __forceinline__ void issue_head(int i) const {
    const long long t = first + (long long)i * stride, b0 = blob_ptr[2 * t];
    const uint32_t bytes = (uint32_t)(blob_ptr[2 * t + 1] - b0);
    mbar_expect_tx(&mbar[i % 3], bytes);
    tma_bulk_g2s(head(i), blob + b0, bytes, &mbar[i % 3]);
  }
  __device__ __forceinline__ void issue_body(int i) const {
    const long long t = first + (long long)i * stride, b0 = blob_ptr[2 * t + 1];
    const uint32_t bytes = (uint32_t)(blob_ptr[2 * t + 2] - b0);
    mbar_expect_tx(&mbar[3 + (i & 1)], bytes);
    tma_bulk_g2s(body(i), blob + b0, bytes, &mbar[3 + (i & 1)]);
  }
  __device__ __forceinline__ void wait_head(int i) const { mbar_wait(&mbar[i % 3], (i / 3) & 1); }
  __device__ __forceinline__ void wait_body(int i) const { mbar_wait(&mbar[3 + (i & 1)], (i >> 1) & 1); }
  // called by ONE thread after the CTA-wide barrier that ends tile i: its body buffer and head buffer are free
  __device__ __forceinline__ void refill_after(int i) const {
    if (i + 2 < count) issue_body(i + 2);
    if (i + 3 < count) issue_head(i + 3);
  }
  __device__ __forceinline__ void prologue() const {
    for (int i = 0; i < 3 && i < count; i++) issue_head(i);
    for (int i = 0; i < 2 && i < count; i++) issue_body(i);
  }
};

// decoded view of a forward tile (layout: plan.h)
struct FwdView {
  int nrows, nel, nvt, ncls, ent32;
  const int* elems;
  const unsigned* rstart; const unsigned short* rlen; const unsigned short* tv; const double* xy;
  const int* cls; const unsigned char* dst; const unsigned short* src;
  __device__ __forceinline__ FwdView(const unsigned char* head, const unsigned char* body, int nvl, int dim, bool has_rlen) {
    const int* hdr = reinterpret_cast<const int*>(head);
    nrows = hdr[0]; nel = hdr[1]; nvt = hdr[2]; ncls = hdr[5]; ent32 = hdr[6] & 1;
    const int ndst = hdr[3];
    elems = reinterpret_cast<const int*>(head + 32);
    unsigned o = 0;
    rstart = reinterpret_cast<const unsigned*>(body + o); o += a16(4u * nrows);
    rlen = reinterpret_cast<const unsigned short*>(body + o); if (has_rlen) o += a16(2u * nrows);
    tv = reinterpret_cast<const unsigned short*>(body + o); o += a16(2u * nvl * nel);
    xy = reinterpret_cast<const double*>(body + o); o += a16(8u * dim * nvt);
    cls = reinterpret_cast<const int*>(body + o); o += a16(16u * ncls);
    dst = body + o; o += a16((ent32 ? 4u : 2u) * ndst);
    src = reinterpret_cast<const unsigned short*>(body + o);
  }
  __device__ __forceinline__ void dest(int i, int& lr, int& j) const {
    if (ent32) { const unsigned v = reinterpret_cast<const unsigned*>(dst)[i]; lr = v & 0xffffu; j = v >> 16; }
    else { const unsigned v = reinterpret_cast<const unsigned short*>(dst)[i]; lr = v & 0xffu; j = v >> 8; }
  }
};

// phase B of the scalar forward: every gather item sums its sources in a fixed order and is written once (twice for a
// paired item: the (r,c) and (c,r) entries of a symmetric local-matrix sum)
__device__ __forceinline__ void fwd_gather_scalar(const FwdView& V, const double* __restrict__ loc, double* __restrict__ vals, int tid, int nth) {
  for (int c = 0; c < V.ncls; c++) {
    const int key = V.cls[4 * c], cnt = key & 0xffff, paired = key >> 16, n = V.cls[4 * c + 1];
    const unsigned short* sc = V.src + V.cls[4 * c + 2];
    const int d0 = V.cls[4 * c + 3];
    auto put = [&](int i, double v) {
      int lr, j; V.dest(d0 + i, lr, j);
      vals[(size_t)V.rstart[lr] + j] = v;
      if (paired) { V.dest(d0 + n + i, lr, j); vals[(size_t)V.rstart[lr] + j] = v; }
    };
#define ADFEM_GATHER_CLASS(C)                                                              \
    for (int i = tid; i < n; i += nth) {                                                 \
      double v = 0.0;                                                                      \
      _Pragma("unroll") for (int k = 0; k < C; k++) v += loc[sc[k * n + i]];             \
      put(i, v);                                                                           \
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
          put(i, v);
        }
    }
#undef ADFEM_GATHER_CLASS
  }
}

// Forward.  Phase A evaluates the local matrices of every element touching the tile's rows into shared memory; phase B
// lets each CSR entry sum its contributions in a fixed (column, element) order and writes it once.  KPRE (P1 scalar
// operators, g <= PIPE_GMAX Gauss points, at most PIPE_EPT tile elements per thread): the coefficients of the NEXT tile
// are loaded into registers before phase B of the current one.  CST (other scalar operators: P2, or more Gauss points): they are
// copied asynchronously (LDGSTS) into a shared-memory staging buffer instead.  Elasticity: prefetched into L2 at that point.
constexpr int PIPE_GMAX = 4, PIPE_EPT = 2;
template <int DIM, int DEG, int OP, bool KPRE, bool CST>
__global__ void __launch_bounds__(TILE_MAX_THREADS) k_tile_fwd(DevMesh m, long long nnz_s, DevTiles tp, const double* __restrict__ coef,
                                                               double* __restrict__ vals) {
  constexpr int D = ElemTraits<DIM, DEG>::D, NC = OP == OP_STIFFNESS ? DIM : 1, Dt = NC * D, dd = D * D, NVL = DIM + 1;
  static_assert(!KPRE || (DEG == 1 && OP != OP_STIFFNESS), "register prefetch is for P1 scalar operators");
  static_assert(!CST || (!KPRE && (OP != OP_STIFFNESS || DEG == 1)), "coefficient staging: scalar operators without register prefetch, P1 elasticity");
  extern __shared__ __align__(128) unsigned char smem_all[];
  __shared__ __align__(8) uint64_t mbar[5];
  const int tid = threadIdx.x, nth = blockDim.x, g = m.g;
  const int cpe = g * (OP == OP_STIFFNESS ? Voigt<DIM>::NS * Voigt<DIM>::NS : 1);      // coefficients per element
  TileRing R{smem_all, smem_all + (size_t)3 * tp.max_head, mbar, tp.max_head, tp.max_body, tp.blob_ptr, tp.blob,
             (int)blockIdx.x, (int)gridDim.x, ((int)blockIdx.x < tp.ntiles) ? (tp.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0};
  double* loc = reinterpret_cast<double*>(R.bodies + (size_t)2 * tp.max_body);
  // CST: a staging buffer [g][nel] behind the local matrices; a thread copies (cp.async) and later reads only ITS elements' coefficients,
  // and the copies for tile i+1 are issued after the barrier that ends phase A of tile i, so one buffer suffices
  // P1 elasticity (CST): the raw coefficient blocks [nel][cpe] of the tile, copied with a flat (coalesced) index by all threads.
  double* cst_all = loc + (size_t)(OP == OP_STIFFNESS ? Dt * Dt : D * (D + 1) / 2) * tp.max_elems;
  // P2 scalar operators: moment tables of the quadrature rule (p2_local_matrix), built once per CTA
  constexpr bool P2TAB = DEG == 2 && OP != OP_STIFFNESS;
  __shared__ double p2tab[P2TAB ? p2_table_rows<DIM, P2TAB ? OP : OP_LAPLACE>() * MAX_QP : 1];
  if constexpr (P2TAB) p2_build_table<DIM, OP>(m.rule, g, tid, nth, p2tab);
  if (tid == 0) { for (int i = 0; i < 5; i++) mbar_init(&mbar[i], 1); }
  __syncthreads();
  if (R.count == 0) return;
  if (tid == 0) R.prologue();
  double kr[KPRE ? PIPE_EPT : 1][KPRE ? PIPE_GMAX : 1];
  auto fetch_coef = [&](const unsigned char* head, int buf) {      // coefficients of this thread's elements of the tile with this head
    const int* hdr = reinterpret_cast<const int*>(head);
    const int nel = hdr[1];
    const int* elems = hdr + 8;
    if constexpr (KPRE) {
#pragma unroll
      for (int s = 0; s < PIPE_EPT; s++) {
        const int le = tid + s * nth;
        if (le < nel) {
          const double* p = coef + (size_t)elems[le] * g;
#pragma unroll
          for (int k = 0; k < PIPE_GMAX; k++) if (k < g) kr[s][k] = __ldg(p + k);
        }
      }
    } else if constexpr (CST && OP == OP_STIFFNESS) {
      // flat index over (element, coefficient): consecutive threads copy consecutive doubles of consecutive elements
      for (int idx = tid; idx < nel * cpe; idx += nth) {
        const int le = idx / cpe, c = idx - le * cpe;
        cp_async8(cst_all + idx, coef + (size_t)elems[le] * cpe + c);
      }
    } else if constexpr (CST) {
      for (int le = tid; le < nel; le += nth) {
        const double* p = coef + (size_t)elems[le] * g;
        for (int k = 0; k < g; k++) cp_async8(cst_all + k * nel + le, p + k);
      }
    } else {
      for (int le = tid; le < nel; le += nth) {
        const double* p = coef + (size_t)elems[le] * cpe;
        for (int b = 0; b < cpe; b += 4) prefetch_l2(p + b);
        prefetch_l2(p + cpe - 1);
      }
    }
    (void)buf;
  };
  R.wait_head(0);
  fetch_coef(R.head(0), 0);
  for (int i = 0; i < R.count; i++) {
    R.wait_body(i);
    const FwdView V(R.head(i), R.body(i), NVL, DIM, NC > 1);
    // ---- phase A
    if constexpr (KPRE) {
#pragma unroll
      for (int s = 0; s < PIPE_EPT; s++) {
        const int le = tid + s * nth;
        if (le < V.nel) {
          Geom<DIM> G; tile_geom(V.tv, V.xy, V.nel, le, m.heron, G);
          local_matrix_scalar<DIM, DEG, OP, PIPE_GMAX>(m, G, [&](int k) { return kr[s][k]; }, [&](int slot, double v) { loc[slot * V.nel + le] = v; });
        }
      }
    } else if constexpr (CST && OP == OP_STIFFNESS) {
      constexpr int NS2 = Voigt<DIM>::NS * Voigt<DIM>::NS;
      cp_async_wait_all();
      __syncthreads();                                     // the copies of every thread have landed
      for (int le = tid; le < V.nel; le += nth) {
        Geom<DIM> G; tile_geom(V.tv, V.xy, V.nel, le, m.heron, G);
        const double* he = cst_all + (size_t)le * cpe;
        double H[NS2];
#pragma unroll
        for (int c = 0; c < NS2; c++) H[c] = 0.0;
        for (int k = 0; k < g; k++) {
          const double w = m.rule.w[k] * G.wscale;
#pragma unroll
          for (int c = 0; c < NS2; c++) H[c] += he[k * NS2 + c] * w;
        }
        stiffness_p1_blocks<DIM>(G, H, [&](int slot, double v) { loc[slot * V.nel + le] = v; });
      }
    } else if constexpr (CST) {
      cp_async_wait_all();                                 // this thread's copies of tile i (requested one tile ago) have landed
      const double* cs = cst_all;
      for (int le = tid; le < V.nel; le += nth) {
        Geom<DIM> G; tile_geom(V.tv, V.xy, V.nel, le, m.heron, G);
        if constexpr (P2TAB) p2_local_matrix<DIM, OP>(g, G, p2tab, [&](int k) { return cs[k * V.nel + le]; }, [&](int slot, double v) { loc[slot * V.nel + le] = v; });
        else local_matrix_scalar<DIM, DEG, OP, 0>(m, G, [&](int k) { return cs[k * V.nel + le]; }, [&](int slot, double v) { loc[slot * V.nel + le] = v; });
      }
    } else {
      for (int le = tid; le < V.nel; le += nth) {
        Geom<DIM> G; tile_geom(V.tv, V.xy, V.nel, le, m.heron, G);
        local_matrix<DIM, DEG, OP>(m, G, V.elems[le], coef, [&](int slot, double v) {
          if (slot >= 0) loc[slot * V.nel + le] = v; else loc[(-slot - 1) * V.nel + le] += v;
        });
      }
    }
    __syncthreads();
    if (i + 1 < R.count) { R.wait_head(i + 1); fetch_coef(R.head(i + 1), (i + 1) & 1); }
    // ---- phase B
    if constexpr (NC == 1) {
      fwd_gather_scalar(V, loc, vals, tid, nth);
    } else {
      for (int c = 0; c < V.ncls; c++) {
        const int cnt = V.cls[4 * c] & 0xffff, n = V.cls[4 * c + 1];
        const unsigned short* sc = V.src + V.cls[4 * c + 2];
        const int d0 = V.cls[4 * c + 3];
        for (int it = tid; it < n; it += nth) {
          double v[NC * NC];
#pragma unroll
          for (int ab = 0; ab < NC * NC; ab++) v[ab] = 0.0;
          for (int k = 0; k < cnt; k++) {
            const int cc = sc[k * n + it], le = cc / dd, pq = cc - le * dd, p = pq / D, q = pq - p * D;
#pragma unroll
            for (int a = 0; a < NC; a++)
#pragma unroll
              for (int b = 0; b < NC; b++) v[a * NC + b] += loc[((a * D + p) * Dt + b * D + q) * V.nel + le];
          }
          int lr, j; V.dest(d0 + it, lr, j);
          const long long len = V.rlen[lr], rs = V.rstart[lr];
#pragma unroll
          for (int a = 0; a < NC; a++)
#pragma unroll
            for (int b = 0; b < NC; b++) vals[NC * (a * nnz_s + rs) + b * len + j] = v[a * NC + b];
        }
      }
    }
    __syncthreads();                                       // loc and the buffers of tile i are free again
    if (tid == 0) R.refill_after(i);
  }
}

// Forward, scalar operators, ONE CTA-wide barrier per tile ("tile_overlap"): the local matrices are double-buffered, and an iteration runs
// phase B of tile i (gather from loc[i & 1]) followed by phase A of tile i + 1 (local matrices into loc[(i + 1) & 1]) with no barrier in
// between, so warps drift apart — some still gather while others already evaluate elements — instead of all meeting twice per tile.  Phase B
// comes first so that the body of tile i + 2 (requested when tile i is released) has the whole of phase B of tile i + 1 to arrive: the ring of
// two bodies suffices.  Costs D (D + 1) / 2 more doubles of shared memory per tile element (smaller tiles); same blobs, same summation order,
// bit-identical values.
template <int DIM, int DEG, int OP, bool KPRE>
__global__ void __launch_bounds__(TILE_MAX_THREADS) k_tile_fwd_ov(DevMesh m, long long nnz_s, DevTiles tp, const double* __restrict__ coef,
                                                                  double* __restrict__ vals) {
  static_assert(OP != OP_STIFFNESS, "scalar operators only");
  static_assert(!KPRE || DEG == 1, "register prefetch is for P1 scalar operators");
  constexpr int D = ElemTraits<DIM, DEG>::D, NVL = DIM + 1, LS = D * (D + 1) / 2;
  extern __shared__ __align__(128) unsigned char smem_all[];
  __shared__ __align__(8) uint64_t mbar[5];
  const int tid = threadIdx.x, nth = blockDim.x, g = m.g;
  TileRing R{smem_all, smem_all + (size_t)3 * tp.max_head, mbar, tp.max_head, tp.max_body, tp.blob_ptr, tp.blob,
             (int)blockIdx.x, (int)gridDim.x, ((int)blockIdx.x < tp.ntiles) ? (tp.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0};
  double* loc_all = reinterpret_cast<double*>(R.bodies + (size_t)2 * tp.max_body);      // two buffers of LS * max_elems doubles
  const size_t loc_stride = (size_t)LS * tp.max_elems;
  double* cst_all = loc_all + 2 * loc_stride;                                           // staging [g][nel] of the tile whose phase A comes next (not KPRE)
  constexpr bool P2TAB = DEG == 2;
  __shared__ double p2tab[P2TAB ? p2_table_rows<DIM, OP>() * MAX_QP : 1];
  if constexpr (P2TAB) p2_build_table<DIM, OP>(m.rule, g, tid, nth, p2tab);
  if (tid == 0) { for (int i = 0; i < 5; i++) mbar_init(&mbar[i], 1); }
  __syncthreads();
  if (R.count == 0) return;
  if (tid == 0) R.prologue();
  double kr[KPRE ? PIPE_EPT : 1][KPRE ? PIPE_GMAX : 1];
  auto fetch_coef = [&](const unsigned char* head) {
    const int* hdr = reinterpret_cast<const int*>(head);
    const int nel = hdr[1];
    const int* elems = hdr + 8;
    if constexpr (KPRE) {
#pragma unroll
      for (int s = 0; s < PIPE_EPT; s++) {
        const int le = tid + s * nth;
        if (le < nel) {
          const double* p = coef + (size_t)elems[le] * g;
#pragma unroll
          for (int k = 0; k < PIPE_GMAX; k++) if (k < g) kr[s][k] = __ldg(p + k);
        }
      }
    } else {
      for (int le = tid; le < nel; le += nth) {
        const double* p = coef + (size_t)elems[le] * g;
        for (int k = 0; k < g; k++) cp_async8(cst_all + k * nel + le, p + k);
      }
    }
  };
  auto phase_a = [&](const FwdView& V, double* loc) {
    if constexpr (KPRE) {
#pragma unroll
      for (int s = 0; s < PIPE_EPT; s++) {
        const int le = tid + s * nth;
        if (le < V.nel) {
          Geom<DIM> G; tile_geom(V.tv, V.xy, V.nel, le, m.heron, G);
          local_matrix_scalar<DIM, DEG, OP, PIPE_GMAX>(m, G, [&](int k) { return kr[s][k]; }, [&](int slot, double v) { loc[slot * V.nel + le] = v; });
        }
      }
    } else {
      cp_async_wait_all();                                 // this thread's copies (it reads only what it copied)
      const double* cs = cst_all;
      for (int le = tid; le < V.nel; le += nth) {
        Geom<DIM> G; tile_geom(V.tv, V.xy, V.nel, le, m.heron, G);
        if constexpr (P2TAB) p2_local_matrix<DIM, OP>(g, G, p2tab, [&](int k) { return cs[k * V.nel + le]; }, [&](int slot, double v) { loc[slot * V.nel + le] = v; });
        else local_matrix_scalar<DIM, DEG, OP, 0>(m, G, [&](int k) { return cs[k * V.nel + le]; }, [&](int slot, double v) { loc[slot * V.nel + le] = v; });
      }
    }
  };
  // prologue: local matrices of the first tile
  R.wait_head(0);
  fetch_coef(R.head(0));
  R.wait_body(0);
  { const FwdView V0(R.head(0), R.body(0), NVL, DIM, false); phase_a(V0, loc_all); }
  __syncthreads();
  for (int i = 0; i < R.count; i++) {
    const FwdView V(R.head(i), R.body(i), NVL, DIM, false);
    const bool more = i + 1 < R.count;
    if (more) { R.wait_head(i + 1); fetch_coef(R.head(i + 1)); }      // in flight during phase B
    fwd_gather_scalar(V, loc_all + (size_t)(i & 1) * loc_stride, vals, tid, nth);
    if (more) {
      R.wait_body(i + 1);
      const FwdView Vn(R.head(i + 1), R.body(i + 1), NVL, DIM, false);
      phase_a(Vn, loc_all + (size_t)((i + 1) & 1) * loc_stride);
    }
    __syncthreads();                                       // loc[(i+1)&1] complete; loc[i&1], body(i), head(i) free
    if (tid == 0) R.refill_after(i);
  }
}

// decoded view of an adjoint tile
struct AdjView {
  int nrows, nel, nvt, nnz_t, lrow16;
  const unsigned* delta; const unsigned short* roff; const unsigned char* lrow;
  const int* elems; const unsigned short* tv; const double* xy; const unsigned short* td; const unsigned* gpk;
  __device__ __forceinline__ void set_head(const unsigned char* head) {
    const int* hdr = reinterpret_cast<const int*>(head);
    nrows = hdr[0]; nel = hdr[1]; nvt = hdr[2]; nnz_t = hdr[3]; lrow16 = hdr[4] & 1;
    unsigned o = 32;
    delta = reinterpret_cast<const unsigned*>(head + o); o += a16(4u * nrows);
    roff = reinterpret_cast<const unsigned short*>(head + o); o += a16(2u * (nrows + 1));
    lrow = head + o;
  }
  __device__ __forceinline__ void set_body(const unsigned char* head, const unsigned char* body, int nvl, int dim, int d) {
    const int has_td = (reinterpret_cast<const int*>(head)[4] >> 1) & 1;
    unsigned o = 0;
    elems = reinterpret_cast<const int*>(body + o); o += a16(4u * nel);
    tv = reinterpret_cast<const unsigned short*>(body + o); o += a16(2u * nvl * nel);
    xy = reinterpret_cast<const double*>(body + o); o += a16(8u * dim * nvt);
    td = has_td ? reinterpret_cast<const unsigned short*>(body + o) : tv; if (has_td) o += a16(2u * d * nel);
    gpk = reinterpret_cast<const unsigned*>(body + o);
  }
};

// Adjoint.  The CSR rows the tile's elements touch are staged into shared memory with asynchronous copies (LDGSTS) one
// tile ahead; every element then gathers its d*d upstream gradients from shared memory (row base roff[td_p] + 8-bit
// position, four positions per 32-bit word) and contracts them with its shape tables.  One CTA-wide barrier per tile.
template <int DIM, int DEG, int OP>
__global__ void __launch_bounds__(TILE_MAX_THREADS) k_tile_adj(DevMesh m, long long nnz_s, DevTiles ap, const double* __restrict__ dvals,
                                                               double* __restrict__ grad_coef) {
  constexpr int D = ElemTraits<DIM, DEG>::D, NC = OP == OP_STIFFNESS ? DIM : 1, NVL = DIM + 1, W = (D + 3) / 4;
  extern __shared__ __align__(128) unsigned char smem_all[];
  __shared__ __align__(8) uint64_t mbar[5];
  const int tid = threadIdx.x, nth = blockDim.x;
  TileRing R{smem_all, smem_all + (size_t)3 * ap.max_head, mbar, ap.max_head, ap.max_body, ap.blob_ptr, ap.blob,
             (int)blockIdx.x, (int)gridDim.x, ((int)blockIdx.x < ap.ntiles) ? (ap.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0};
  double* sd_all = reinterpret_cast<double*>(R.bodies + (size_t)2 * ap.max_body);
  const size_t sd_stride = (size_t)NC * NC * ap.max_nnz;
  double* gacc = sd_all + 2 * sd_stride;                   // P1 elasticity only: [NS*NS][nel] gradient matrices of the tile's elements
  (void)gacc;
  constexpr bool P2TAB = DEG == 2 && OP != OP_STIFFNESS;
  __shared__ double p2tab[P2TAB ? p2_table_rows<DIM, P2TAB ? OP : OP_LAPLACE>() * MAX_QP : 1];
  if constexpr (P2TAB) p2_build_table<DIM, OP>(m.rule, m.g, tid, nth, p2tab);
  if (tid == 0) { for (int i = 0; i < 5; i++) mbar_init(&mbar[i], 1); }
  __syncthreads();
  if (R.count == 0) return;
  if (tid == 0) R.prologue();
  auto stage = [&](const unsigned char* head, double* sd) {   // request the CSR row segments of the tile with this head
    AdjView H; H.set_head(head);
    for (int i = tid; i < H.nnz_t; i += nth) {
      const int lr = H.lrow16 ? (int)reinterpret_cast<const unsigned short*>(H.lrow)[i] : (int)H.lrow[i];
      if (NC == 1) cp_async8(sd + i, dvals + (unsigned)(i + H.delta[lr]));
      else {
        const long long len = H.roff[lr + 1] - H.roff[lr], rs = (unsigned)(H.roff[lr] + H.delta[lr]), j = i - H.roff[lr];
#pragma unroll
        for (int a = 0; a < NC; a++)
#pragma unroll
          for (int b = 0; b < NC; b++) cp_async8(sd + (a * NC + b) * H.nnz_t + i, dvals + (NC * (a * nnz_s + rs) + b * len + j));
      }
    }
  };
  R.wait_head(0);
  stage(R.head(0), sd_all);
  cp_async_wait_all();
  __syncthreads();
  for (int i = 0; i < R.count; i++) {
    // here: sd[i&1] is complete and visible; every thread is done with tile i-1
    const double* sd = sd_all + (i & 1) * sd_stride;
    if (i + 1 < R.count) { R.wait_head(i + 1); stage(R.head(i + 1), sd_all + ((i + 1) & 1) * sd_stride); }
    R.wait_body(i);
    AdjView V; V.set_head(R.head(i)); V.set_body(R.head(i), R.body(i), NVL, DIM, D);
    for (int le = tid; le < V.nel; le += nth) {
      Geom<DIM> G; tile_geom(V.tv, V.xy, V.nel, le, m.heron, G);
      const int e = V.elems[le];
      if constexpr (NC == 1) {
        unsigned pk[D][W]; int rb[D];
#pragma unroll
        for (int p = 0; p < D; p++) {
          rb[p] = V.roff[V.td[p * V.nel + le]];
#pragma unroll
          for (int w = 0; w < W; w++) pk[p][w] = V.gpk[(p * W + w) * V.nel + le];
        }
        auto dK = [&](int p, int q) { return sd[rb[p] + ((pk[p][q >> 2] >> (8 * (q & 3))) & 0xffu)]; };
        if constexpr (P2TAB) {
          double gs[D * (D + 1) / 2];
          { int i = 0;
#pragma unroll
            for (int p = 0; p < D; p++)
#pragma unroll
              for (int q = p; q < D; q++) gs[i++] = (q == p) ? dK(p, p) : dK(p, q) + dK(q, p); }
          double* ge = grad_coef + (size_t)e * m.g;
          p2_local_adjoint<DIM, OP>(m.g, G, p2tab, gs, [&](int k, double v) { ge[k] = v; });
        } else {
          local_adjoint<DIM, DEG, OP>(m, G, e, dK, grad_coef);
        }
      } else {
        auto dK = [&](int l, int s) {
          const int a = l / D, p = l % D, b = s / D, q = s % D;
          const unsigned pos = (V.gpk[(p * W + (q >> 2)) * V.nel + le] >> (8 * (q & 3))) & 0xffu;
          return sd[(a * NC + b) * V.nnz_t + V.roff[V.td[p * V.nel + le]] + pos];
        };
        if constexpr (DEG == 1) {
          // P1: one gradient matrix per element (times w_k).  A thread writing the NS*NS*g gradients of ITS element touches 32
          // different sectors per warp store, so the matrix is parked in shared memory and written by consecutive lanes below.
          constexpr int NS2 = Voigt<DIM>::NS * Voigt<DIM>::NS;
          double gH[NS2];
          stiffness_p1_adjoint<DIM>(G, dK, gH);
#pragma unroll
          for (int c = 0; c < NS2; c++) gacc[c * V.nel + le] = gH[c] * G.wscale;
        } else {
          local_adjoint<DIM, DEG, OP>(m, G, e, dK, grad_coef);
        }
      }
    }
    if constexpr (NC > 1 && DEG == 1) {
      // flat index (element, entry): consecutive threads write consecutive entries of consecutive elements
      constexpr int NS2 = Voigt<DIM>::NS * Voigt<DIM>::NS;
      __syncthreads();
      for (int idx = tid; idx < V.nel * NS2; idx += nth) {
        const int le = idx / NS2, c = idx - le * NS2;
        double* ge = grad_coef + (size_t)V.elems[le] * m.g * NS2 + c;
        const double v = gacc[c * V.nel + le];
        for (int k = 0; k < m.g; k++) ge[k * NS2] = v * m.rule.w[k];
      }
    }
    cp_async_wait_all();
    __syncthreads();                                       // tile i is finished everywhere AND sd[(i+1)&1] is complete and visible
    if (tid == 0) R.refill_after(i);
  }
}

}  // namespace adfem
