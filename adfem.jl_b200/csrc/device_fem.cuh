// Per-element finite-element arithmetic on the device (fp64, sm_100a).
//
// The reference precomputes h/hx/hy/hz/w per element and Gauss point on the host
// (deps/MFEM/Common.cpp:59-131, deps/MFEM3/Common.cpp:62-134) and streams those heap tables in every
// assembly loop.  Here the only mesh data read by a kernel are the connectivity and the vertex
// coordinates; barycentric gradients, areas/volumes, weights and the P1/P2 shape values are
// recomputed in registers.
#pragma once
#include <cstdint>

#include "quadrature.h"

namespace adfem {

struct DevMesh {            // passed by value to kernels
  int dim, ne, nv, d, g, ndof;
  int heron;                // 2-D weight scale: 1 = Heron area like the reference, 0 = det/2
  const double* coords;     // nv x dim packed
  const int* verts;         // [(dim+1)][ne]  struct-of-arrays, post orientation fix
  const int* conn;          // [d][ne]        struct-of-arrays dof ids
  QuadRule rule;
};

enum Op : int { OP_LAPLACE = 0, OP_MASS = 1, OP_STIFFNESS = 2 };

template <int DIM, int DEG> struct ElemTraits;
template <> struct ElemTraits<2, 1> { static constexpr int D = 3; };
template <> struct ElemTraits<2, 2> { static constexpr int D = 6; };
template <> struct ElemTraits<3, 1> { static constexpr int D = 4; };
template <> struct ElemTraits<3, 2> { static constexpr int D = 10; };

template <int DIM> struct Geom {
  double gL[DIM + 1][DIM];   // physical gradients of the barycentric coordinates
  double wscale;             // w_k = rule.w[k] * wscale
};

// Host + device: the per-element arithmetic below is also compiled for the host by the test-only emulation harness
// (tests/host_emul/), which runs the kernel BODIES in plain loops against the oracle on machines without a GPU.
#define ADFEM_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
ADFEM_HD double ldg(const double* p) { return __ldg(p); }
ADFEM_HD int ldg(const int* p) { return __ldg(p); }
ADFEM_HD double2 ldg(const double2* p) { return __ldg(p); }
#else
ADFEM_HD double ldg(const double* p) { return *p; }
ADFEM_HD int ldg(const int* p) { return *p; }
ADFEM_HD double2 ldg(const double2* p) { return *p; }
#endif

// vertex ids of element e
template <int DIM> ADFEM_HD void load_verts(const DevMesh& m, int e, int v[DIM + 1]) {
#pragma unroll
  for (int k = 0; k <= DIM; k++) v[k] = ldg(m.verts + (size_t)k * m.ne + e);
}

// Geometry of a triangle from its (orientation-fixed) vertices: gradients of the barycentric coordinates
// as MFEM's CalcPhysDShape (adj(J)/det).  Weight scale: the reference uses w = ip.weight * area / 0.5 with the
// HERON area (deps/MFEM/Common.cpp:9-15,83,116); `heron` = 0 uses area = det/2 instead (identical up to the
// rounding of Heron's formula, 4 sqrt cheaper).
ADFEM_HD void geom_tri(const double2 p1, const double2 p2, const double2 p3, int heron, Geom<2>& G) {
  const double det = (p2.x - p1.x) * (p3.y - p1.y) - (p3.x - p1.x) * (p2.y - p1.y);
  const double inv = 1.0 / det;
  G.gL[1][0] = (p3.y - p1.y) * inv;  G.gL[1][1] = -(p3.x - p1.x) * inv;
  G.gL[2][0] = -(p2.y - p1.y) * inv; G.gL[2][1] = (p2.x - p1.x) * inv;
  G.gL[0][0] = -G.gL[1][0] - G.gL[2][0]; G.gL[0][1] = -G.gL[1][1] - G.gL[2][1];
  if (heron) {
    const double a = sqrt((p1.x - p2.x) * (p1.x - p2.x) + (p1.y - p2.y) * (p1.y - p2.y));
    const double b = sqrt((p3.x - p2.x) * (p3.x - p2.x) + (p3.y - p2.y) * (p3.y - p2.y));
    const double c = sqrt((p1.x - p3.x) * (p1.x - p3.x) + (p1.y - p3.y) * (p1.y - p3.y));
    const double s = (a + b + c) / 2.0;
    G.wscale = sqrt(s * (s - a) * (s - b) * (s - c)) / 0.5;
  } else {
    G.wscale = det;
  }
}

// Geometry of a tetrahedron: volume = det/6 (Mesh::GetElementVolume), w = ip.weight * volume * 6
// (deps/MFEM3/Common.cpp:88,117).
ADFEM_HD void geom_tet(const double X[4][3], Geom<3>& G) {
  double J[3][3];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) J[r][c] = X[c + 1][r] - X[0][r];
  const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                     J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  const double inv = 1.0 / det;
  G.gL[1][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * inv; G.gL[1][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * inv; G.gL[1][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * inv;
  G.gL[2][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * inv; G.gL[2][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * inv; G.gL[2][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * inv;
  G.gL[3][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * inv; G.gL[3][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * inv; G.gL[3][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * inv;
#pragma unroll
  for (int c = 0; c < 3; c++) G.gL[0][c] = -G.gL[1][c] - G.gL[2][c] - G.gL[3][c];
  G.wscale = det * (1. / 6.) * 6.0;
}

// geometry of element e from the global vertex / coordinate arrays
ADFEM_HD void load_geom(const DevMesh& m, int e, Geom<2>& G) {
  int v[3]; load_verts<2>(m, e, v);
  const double2* X = reinterpret_cast<const double2*>(m.coords);
  geom_tri(ldg(X + v[0]), ldg(X + v[1]), ldg(X + v[2]), m.heron, G);
}
ADFEM_HD void load_geom(const DevMesh& m, int e, Geom<3>& G) {
  int v[4]; load_verts<3>(m, e, v);
  double X[4][3];
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int c = 0; c < 3; c++) X[k][c] = ldg(m.coords + (size_t)v[k] * 3 + c);
  geom_tet(X, G);
}
// geometry of tile element `le` from a tile blob staged in shared memory: tv = k-major tile-local vertex ids,
// xy = coordinates of the tile-local vertices
ADFEM_HD void tile_geom(const unsigned short* tv, const double* xy, int nel, int le, int heron, Geom<2>& G) {
  const double2* X = reinterpret_cast<const double2*>(xy);
  geom_tri(X[tv[le]], X[tv[nel + le]], X[tv[2 * nel + le]], heron, G);
}
ADFEM_HD void tile_geom(const unsigned short* tv, const double* xy, int nel, int le, int, Geom<3>& G) {
  double X[4][3];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const double* p = xy + 3 * (int)tv[k * nel + le];
#pragma unroll
    for (int c = 0; c < 3; c++) X[k][c] = p[c];
  }
  geom_tet(X, G);
}

template <int DIM> ADFEM_HD void bary(const QuadRule& r, int k, double L[DIM + 1]) {
  if (DIM == 2) { L[0] = 1 - r.x[k] - r.y[k]; L[1] = r.x[k]; L[2] = r.y[k]; }
  else { L[0] = 1 - r.x[k] - r.y[k] - r.z[k]; L[1] = r.x[k]; L[2] = r.y[k]; L[DIM] = r.z[k]; }
}

// local edge (a,b) of edge function j, MFEM geometry order
template <int DIM> ADFEM_HD void edge_ends(int j, int& a, int& b) {
  if (DIM == 2) { a = j; b = j == 2 ? 0 : j + 1; }                        // (0,1) (1,2) (2,0)
  else { a = j < 3 ? 0 : (j < 5 ? 1 : 2); b = j < 3 ? j + 1 : (j < 5 ? j - 1 : 3); }   // (0,1)(0,2)(0,3)(1,2)(1,3)(2,3)
}

// nodal H1 basis values at barycentric point L (H1_TriangleElement / H1_TetrahedronElement, p = DEG)
template <int DIM, int DEG> ADFEM_HD void basis_val(const double L[DIM + 1], double phi[]) {
  constexpr int NV = DIM + 1;
  if (DEG == 1) {
#pragma unroll
    for (int i = 0; i < NV; i++) phi[i] = L[i];
  } else {
#pragma unroll
    for (int i = 0; i < NV; i++) phi[i] = L[i] * (2.0 * L[i] - 1.0);
#pragma unroll
    for (int j = 0; j < ElemTraits<DIM, 2>::D - NV; j++) { int a, b; edge_ends<DIM>(j, a, b); phi[NV + j] = 4.0 * L[a] * L[b]; }
  }
}

// physical gradients of the basis at barycentric point L
template <int DIM, int DEG> ADFEM_HD void basis_grad(const Geom<DIM>& G, const double L[DIM + 1], double gphi[][DIM]) {
  constexpr int NV = DIM + 1;
  if (DEG == 1) {
#pragma unroll
    for (int i = 0; i < NV; i++)
#pragma unroll
      for (int c = 0; c < DIM; c++) gphi[i][c] = G.gL[i][c];
  } else {
#pragma unroll
    for (int i = 0; i < NV; i++)
#pragma unroll
      for (int c = 0; c < DIM; c++) gphi[i][c] = (4.0 * L[i] - 1.0) * G.gL[i][c];
#pragma unroll
    for (int j = 0; j < ElemTraits<DIM, 2>::D - NV; j++) {
      int a, b; edge_ends<DIM>(j, a, b);
#pragma unroll
      for (int c = 0; c < DIM; c++) gphi[NV + j][c] = 4.0 * (L[a] * G.gL[b][c] + L[b] * G.gL[a][c]);
    }
  }
}

template <int DIM> ADFEM_HD double dotg(const double* a, const double* b) {
  double s = a[0] * b[0] + a[1] * b[1];
  if (DIM == 3) s += a[2] * b[2];
  return s;
}

// ---- strain-displacement columns -----------------------------------------------------------------
// Column (component c, node gradient g) of B.  2-D rows [exx, eyy, gxy] as in
// deps/MFEM/ComputeFemStiffnessMatrixMfem/ComputeFemStiffnessMatrixMfem.h:18-23; 3-D extension uses
// Voigt order [xx, yy, zz, yz, xz, xy].
template <int DIM> struct Voigt { static constexpr int NS = DIM == 2 ? 3 : 6; };

// b(c,g) . v
template <int DIM> ADFEM_HD double bdot(int c, const double* g, const double* v) {
  if (DIM == 2) return c == 0 ? g[0] * v[0] + g[1] * v[2] : g[1] * v[1] + g[0] * v[2];
  return c == 0 ? g[0] * v[0] + g[2] * v[4] + g[1] * v[5] : (c == 1 ? g[1] * v[1] + g[2] * v[3] + g[0] * v[5] : g[2] * v[2] + g[1] * v[3] + g[0] * v[4]);
}
// strided variant: b(c,g) . v[0], v[stride], ...
template <int DIM> ADFEM_HD double bdot_s(int c, const double* g, const double* v, int st) {
  if (DIM == 2) return c == 0 ? g[0] * v[0] + g[1] * v[2 * st] : g[1] * v[st] + g[0] * v[2 * st];
  return c == 0 ? g[0] * v[0] + g[2] * v[4 * st] + g[1] * v[5 * st]
                : (c == 1 ? g[1] * v[st] + g[2] * v[3 * st] + g[0] * v[5 * st] : g[2] * v[2 * st] + g[1] * v[3 * st] + g[0] * v[4 * st]);
}
// v += s * b(c,g)
template <int DIM> ADFEM_HD void badd(int c, const double* g, double s, double* v) {
  if (DIM == 2) {
    if (c == 0) { v[0] += s * g[0]; v[2] += s * g[1]; } else { v[1] += s * g[1]; v[2] += s * g[0]; }
  } else {
    if (c == 0) { v[0] += s * g[0]; v[4] += s * g[2]; v[5] += s * g[1]; }
    else if (c == 1) { v[1] += s * g[1]; v[3] += s * g[2]; v[5] += s * g[0]; }
    else { v[2] += s * g[2]; v[3] += s * g[1]; v[4] += s * g[0]; }
  }
}

}  // namespace adfem
