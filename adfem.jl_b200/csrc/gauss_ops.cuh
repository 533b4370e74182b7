// Gauss-point operators on the element tables (SURVEY §8 f, rank 2): matrix-free actions and dof <-> Gauss-point
// transfers that every nonlinear AdFem example interleaves with the assembly ops.
//
//   reference op (deps/MFEM/...)                                  | as a linear map                          | here
//   FemToGaussPoints/FemToGaussPointsMfem.h:6-35                  | u[node]  -> P1 shapes at Gauss points    | gather<P1SHAPE>, adjoint scatter<P1SHAPE>
//   DofToGaussPoints/DofToGaussPointsMfem.h:6-34                  | u[dof]   -> all shapes at Gauss points   | gather<SHAPE>,   adjoint scatter<SHAPE>
//   FemGrad/FemGradMfem.h:8-42                                    | u[dof]   -> (du/dx, du/dy) per point     | gather<GRAD>,    adjoint scatter<GRAD>
//   EvalStrainOnGaussPtsMfem/EvalStrainOnGaussPts.h:4-29          | u[2 ndof]-> (exx, eyy, gxy) per point    | gather<STRAIN>,  adjoint scatter<STRAIN>
//   ComputeStrainEnergyTermMfem/ComputeStrainEnergyTermMfem.h:4-35| sigma[3G]-> int sigma : eps(v) [2 ndof]  | scatter<STRAIN, weighted>, adjoint gather<STRAIN, weighted>
//   ComputeLaplaceTermMfem/ComputeLaplaceTermMfem.h:4-39 (+MFEM3) | (nu, u)  -> int nu grad u . grad v       | laplace_term_row, laplace_term_grad_nu_elem
//
// Every map is either a GATHER (one thread per element: its dof values -> its Gauss points) or the transposed SCATTER,
// which runs as a row gather over the dof -> (element, local dof) adjacency (one thread per dof row, contributions added in
// ascending element order, no atomics, no pre-zeroed output) exactly like the source term (kernels.cuh, k_source_fwd).
// The shape tables h/hx/hy/w the reference streams from its per-element heap objects are recomputed in registers.
//
// All bodies are __host__ __device__ so that tests/host_emul/ can run them in plain host loops against the oracle where
// there is no GPU; the product only ever launches them as kernels (gauss_ops.cu).
#pragma once
#include "device_fem.cuh"

namespace adfem {

enum GpBasis : int { GB_P1SHAPE = 0, GB_SHAPE = 1, GB_GRAD = 2, GB_STRAIN = 3 };

template <int DIM, int DEG, int B> struct GpShape {
  static constexpr int D = ElemTraits<DIM, DEG>::D;
  static constexpr int ND = (B == GB_P1SHAPE) ? DIM + 1 : D;                                  // local functions that take part
  static constexpr int NQ = (B == GB_GRAD) ? DIM : (B == GB_STRAIN ? Voigt<DIM>::NS : 1);     // values per Gauss point (interleaved)
  static constexpr int NC = (B == GB_STRAIN) ? DIM : 1;                                       // components per dof (blocked: dof + c*ndof)
};

// entry p of a register array without dynamic indexing
template <int N> ADFEM_HD double pick(const double* a, int p) {
  double v = 0.0;
#pragma unroll
  for (int q = 0; q < N; q++) v = (q == p) ? a[q] : v;
  return v;
}

// ---- gather: the dof values of ONE element -> its g Gauss points ------------------------------------------------------
// uloc[c * ND + p] = value of component c at local dof p; out[k * NQ + i]
template <int DIM, int DEG, int B, bool W>
ADFEM_HD void gp_gather_elem(const Geom<DIM>& G, const QuadRule& rule, int g, const double* uloc, double* out) {
  using S = GpShape<DIM, DEG, B>;
  constexpr int D = S::D, ND = S::ND, NQ = S::NQ, NC = S::NC;
  for (int k = 0; k < g; k++) {
    double L[DIM + 1]; bary<DIM>(rule, k, L);
    double q[NQ];
#pragma unroll
    for (int i = 0; i < NQ; i++) q[i] = 0.0;
    if (B == GB_P1SHAPE) {
#pragma unroll
      for (int p = 0; p < ND; p++) q[0] += uloc[p] * L[p];
    } else if (B == GB_SHAPE) {
      double phi[D]; basis_val<DIM, DEG>(L, phi);
#pragma unroll
      for (int p = 0; p < ND; p++) q[0] += uloc[p] * phi[p];
    } else {
      double gp[D][DIM]; basis_grad<DIM, DEG>(G, L, gp);
      if (B == GB_GRAD) {
#pragma unroll
        for (int p = 0; p < ND; p++)
#pragma unroll
          for (int i = 0; i < DIM; i++) q[i] += uloc[p] * gp[p][i];
      } else {
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
          for (int p = 0; p < ND; p++) badd<DIM>(c, gp[p], uloc[c * ND + p], q);
      }
    }
    const double s = W ? rule.w[k] * G.wscale : 1.0;
#pragma unroll
    for (int i = 0; i < NQ; i++) out[k * NQ + i] = q[i] * s;
  }
}

// the same for element e of a mesh in device (or, under emulation, host) memory
template <int DIM, int DEG, int B, bool W>
ADFEM_HD void gp_gather_body(const DevMesh& m, int e, const double* in, double* out) {
  using S = GpShape<DIM, DEG, B>;
  Geom<DIM> G; load_geom(m, e, G);
  double uloc[S::NC * S::ND];
#pragma unroll
  for (int p = 0; p < S::ND; p++) {
    const int dof = ldg(m.conn + (size_t)p * m.ne + e);       // conn[0..DIM] are the vertices for both degrees
#pragma unroll
    for (int c = 0; c < S::NC; c++) uloc[c * S::ND + p] = ldg(in + dof + (size_t)c * m.ndof);
  }
  gp_gather_elem<DIM, DEG, B, W>(G, m.rule, m.g, uloc, out + (size_t)e * m.g * S::NQ);
}

// ---- scatter: Gauss-point values -> ONE dof row (all NC components) ------------------------------------------------------
// contribution of element (geometry G, Gauss-point values se[k*NQ + i]) to the row of its local dof p:
// acc[c] += sum over k, i of C_ic(p, k) * se[k*NQ + i] (* w_k)
template <int DIM, int DEG, int B, bool W>
ADFEM_HD void gp_scatter_elem(const Geom<DIM>& G, const QuadRule& rule, int g, int p, const double* se, double* acc) {
  using S = GpShape<DIM, DEG, B>;
  constexpr int D = S::D, NQ = S::NQ, NC = S::NC;
  for (int k = 0; k < g; k++) {
    double L[DIM + 1]; bary<DIM>(rule, k, L);
    const double* sk = se + (size_t)k * NQ;
    const double w = W ? rule.w[k] * G.wscale : 1.0;
    if (B == GB_P1SHAPE) {
      acc[0] += pick<DIM + 1>(L, p) * ldg(sk) * w;
    } else if (B == GB_SHAPE) {
      double phi[D]; basis_val<DIM, DEG>(L, phi);
      acc[0] += pick<D>(phi, p) * ldg(sk) * w;
    } else {
      double gp[D][DIM]; basis_grad<DIM, DEG>(G, L, gp);
      double gs[DIM];
#pragma unroll
      for (int i = 0; i < DIM; i++) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < D; q++) v = (q == p) ? gp[q][i] : v;
        gs[i] = v;
      }
      double sv[NQ];
#pragma unroll
      for (int i = 0; i < NQ; i++) sv[i] = ldg(sk + i);
      if (B == GB_GRAD) {
        acc[0] += dotg<DIM>(gs, sv) * w;
      } else {
#pragma unroll
        for (int c = 0; c < NC; c++) acc[c] += bdot<DIM>(c, gs, sv) * w;
      }
    }
  }
}

// acc[c] = sum over the incident (e, p) of row r, in ascending element order
template <int DIM, int DEG, int B, bool W>
ADFEM_HD void gp_scatter_row(const DevMesh& m, const long long* adj_ptr, const int* adj_elem, const uint8_t* adj_loc, int r,
                             const double* s, double* acc) {
  using S = GpShape<DIM, DEG, B>;
#pragma unroll
  for (int c = 0; c < S::NC; c++) acc[c] = 0.0;
  for (long long a = adj_ptr[r]; a < adj_ptr[r + 1]; a++) {
    const int e = adj_elem[a], p = adj_loc[a];
    Geom<DIM> G; load_geom(m, e, G);
    gp_scatter_elem<DIM, DEG, B, W>(G, m.rule, m.g, p, s + (size_t)e * m.g * S::NQ, acc);
  }
}

// ---- Laplace term: out[r] = sum over incident (e, p), k of nu[e,k] w_k grad phi_p . (sum_q grad phi_q u[dof_q]) -------------
// contribution of one element (geometry G, coefficients nue[k], dof values ul[q]) to the row of its local dof p
template <int DIM, int DEG>
ADFEM_HD double laplace_term_elem(const Geom<DIM>& G, const QuadRule& rule, int g, int p, const double* nue, const double* ul) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  double acc = 0.0;
  for (int k = 0; k < g; k++) {
    double L[DIM + 1]; bary<DIM>(rule, k, L);
    double gp[D][DIM]; basis_grad<DIM, DEG>(G, L, gp);
    double gu[DIM], gs[DIM];
#pragma unroll
    for (int i = 0; i < DIM; i++) {
      double su = 0.0, v = 0.0;
#pragma unroll
      for (int q = 0; q < D; q++) { su += gp[q][i] * ul[q]; v = (q == p) ? gp[q][i] : v; }
      gu[i] = su; gs[i] = v;
    }
    acc += ldg(nue + k) * (rule.w[k] * G.wscale) * dotg<DIM>(gs, gu);
  }
  return acc;
}
template <int DIM, int DEG>
ADFEM_HD double laplace_term_row(const DevMesh& m, const long long* adj_ptr, const int* adj_elem, const uint8_t* adj_loc, int r,
                                 const double* nu, const double* u) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  double acc = 0.0;
  for (long long a = adj_ptr[r]; a < adj_ptr[r + 1]; a++) {
    const int e = adj_elem[a], p = adj_loc[a];
    Geom<DIM> G; load_geom(m, e, G);
    double ul[D];
#pragma unroll
    for (int q = 0; q < D; q++) ul[q] = ldg(u + ldg(m.conn + (size_t)q * m.ne + e));
    acc += laplace_term_elem<DIM, DEG>(G, m.rule, m.g, p, nu + (size_t)e * m.g, ul);
  }
  return acc;
}

// grad_nu[e,k] = w_k (grad go . grad u) at Gauss point k, go = upstream gradient of the term (ComputeLaplaceTermMfem.h:27-30)
template <int DIM, int DEG>
ADFEM_HD void laplace_term_grad_nu_body(const DevMesh& m, int e, const double* u, const double* go, double* grad_nu) {
  constexpr int D = ElemTraits<DIM, DEG>::D;
  Geom<DIM> G; load_geom(m, e, G);
  double ul[D], gl[D];
#pragma unroll
  for (int q = 0; q < D; q++) {
    const int dof = ldg(m.conn + (size_t)q * m.ne + e);
    ul[q] = ldg(u + dof); gl[q] = ldg(go + dof);
  }
  for (int k = 0; k < m.g; k++) {
    double L[DIM + 1]; bary<DIM>(m.rule, k, L);
    double gp[D][DIM]; basis_grad<DIM, DEG>(G, L, gp);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; i++) {
      double su = 0.0, sg = 0.0;
#pragma unroll
      for (int q = 0; q < D; q++) { su += gp[q][i] * ul[q]; sg += gp[q][i] * gl[q]; }
      s += su * sg;
    }
    grad_nu[(size_t)e * m.g + k] = s * (m.rule.w[k] * G.wscale);
  }
}

// ---- Gauss-summed coefficients for P1 elasticity (option "coef_presum") ------------------------------------------------------
// For P1 elements B is constant per element, so the stiffness block only needs Hbar_e = sum_k w_k H_k (reference weights w_k of the
// quadrature rule; the geometric scale is applied by the tile kernel).  A streaming pre-pass reduces the g blocks of an element to one,
// so that the row-tile kernel — which re-evaluates halo elements — gathers NS*NS instead of g*NS*NS doubles per element evaluation.
// The adjoint runs the other way: one block per element from the tile kernel, expanded to the g Gauss points here.
// idx = e*ns2 + c (flat over the reduced array)
ADFEM_HD double presum_coef_body(const QuadRule& rule, int g, int ns2, long long idx, const double* coef) {
  const long long e = idx / ns2;
  const int c = (int)(idx - e * ns2);
  const double* p = coef + (e * g) * ns2 + c;
  double s = 0.0;
  for (int k = 0; k < g; k++) s += ldg(p + (size_t)k * ns2) * rule.w[k];
  return s;
}
// idx = (e*g + k)*ns2 + c (flat over the per-Gauss-point gradient)
ADFEM_HD double expand_grad_body(const QuadRule& rule, int g, int ns2, long long idx, const double* gbar) {
  const long long t = idx / ns2;                 // e*g + k
  const int c = (int)(idx - t * ns2), k = (int)(t % g);
  return ldg(gbar + (t / g) * ns2 + c) * rule.w[k];
}

// ---- constitutive pre-step (SURVEY §8 f, rank 3): per-point 3x3 tangent from (E, nu) -----------------------------------------
// deps/MFEM/PlaneStrainAndStress/PlaneStrainAndStress.h:5-19 (mode 0, `PlaneStrainMatrix`: s on the diagonal, s nu/(1-nu) in
// ALL six off-diagonal entries) and :46-60 (mode 1, `PlaneStressMatrix`: E/((1+nu)(1-2nu)) [[1-nu,nu,0],[nu,1-nu,0],[0,0,(1-2nu)/2]]).
// The reference's names and formulas are kept as they are (src/Core.jl:742-768 documents the same two matrices).
ADFEM_HD void plane_matrix_body(int mode, double E, double nu, double* H) {
  if (mode == 0) {
    const double s = E * (1 - nu) / (1 + nu) / (1 - 2 * nu), t = s * nu / (1 - nu);
    H[0] = s; H[1] = t; H[2] = t; H[3] = t; H[4] = s; H[5] = t; H[6] = t; H[7] = t; H[8] = s;
  } else {
    const double s = E / (1 + nu) / (1 - 2 * nu);
    H[0] = s * (1 - nu); H[1] = s * nu; H[2] = 0.0; H[3] = s * nu; H[4] = s * (1 - nu); H[5] = 0.0;
    H[6] = 0.0; H[7] = 0.0; H[8] = s * (1 - 2 * nu) / 2.0;
  }
}
// Adjoint (PlaneStrainAndStress.h:21-44, 62-78; the reference differentiates the same expressions with a tape): with
// D = (1+nu)(1-2nu), f = (1-nu)/D, h = nu/D:  f' = 2nu(2-nu)/D^2,  h' = (1+2nu^2)/D^2.
ADFEM_HD void plane_matrix_grad_body(int mode, double E, double nu, const double* g, double* grad_E, double* grad_nu) {
  const double D = (1 + nu) * (1 - 2 * nu), f = (1 - nu) / D, h = nu / D;
  const double df = 2 * nu * (2 - nu) / (D * D), dh = (1 + 2 * nu * nu) / (D * D);
  if (mode == 0) {
    const double gd = g[0] + g[4] + g[8], go = g[1] + g[2] + g[3] + g[5] + g[6] + g[7];
    *grad_E = gd * f + go * h;
    *grad_nu = E * (gd * df + go * dh);
  } else {
    const double ga = g[0] + g[4], gb = g[1] + g[3], gc = g[8];
    *grad_E = ga * f + gb * h + gc / (2 * (1 + nu));
    *grad_nu = E * (ga * df + gb * dh - gc / (2 * (1 + nu) * (1 + nu)));
  }
}

// ---- constitutive pre-step FUSED with the Gauss sum (SURVEY §8 f rank 3: H is never materialised) ----------------------------
// P1 triangles: hbar[e*9 + c] = sum_k w_k H(E[e*g+k], nu[e*g+k])[c] straight from the moduli (16 B per Gauss point in instead of 72),
// and on the way back (dE, dnu)[e*g+k] = dH/d(E,nu)^T (w_k gbar[e]).  idx = e*9 + c.
ADFEM_HD double presum_plane_body(const QuadRule& rule, int g, int mode, long long idx, const double* E, const double* nu) {
  const long long e = idx / 9;
  const int c = (int)(idx - e * 9);
  double s = 0.0;
  for (int k = 0; k < g; k++) {
    double H[9];
    plane_matrix_body(mode, ldg(E + e * g + k), ldg(nu + e * g + k), H);
    double hc = 0.0;
#pragma unroll
    for (int i = 0; i < 9; i++) hc = (i == c) ? H[i] : hc;
    s += hc * rule.w[k];
  }
  return s;
}
// t = e*g + k
ADFEM_HD void expand_plane_grad_body(const QuadRule& rule, int g, int mode, long long t, const double* E, const double* nu, const double* gbar,
                                     double* grad_E, double* grad_nu) {
  const long long e = t / g;
  const double w = rule.w[(int)(t - e * g)];
  double gk[9];
#pragma unroll
  for (int i = 0; i < 9; i++) gk[i] = ldg(gbar + e * 9 + i) * w;
  plane_matrix_grad_body(mode, ldg(E + t), ldg(nu + t), gk, grad_E + t, grad_nu + t);
}

}  // namespace adfem
