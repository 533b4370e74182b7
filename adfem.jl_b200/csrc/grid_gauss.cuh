// Scatter-type Gauss-point operators on the reference's structured triangulation `Mesh(m, n, h)` (P1): the transposed transfers, the
// strain-energy term and the matrix-free Laplace term as ONE THREAD PER NODE that visits its (up to) six incident triangles in ascending
// element order — the order of the dof -> element adjacency walked by the general kernels (gauss_ops.cuh), so both paths produce the same
// bits — with index arithmetic instead of the adjacency, connectivity and coordinate arrays (≈ 250 B of mesh-static data per node in the
// general kernels; here the only streams are the Gauss-point values in and the dof vector out).
// Host + device bodies (tests/host_emul/).
#pragma once
#include "gauss_ops.cuh"
#include "grid_index.cuh"

namespace adfem {

// the six triangles around node (i, j), ascending element id: which cell (di, dj relative to the node's cell (i, j)), which triangle of the
// cell, and the node's local index in it.  T0 = [BL BR TL], T1 = [TL BR TR] (grid_index.cuh).
struct GridIncident { int di, dj, tri, p; };
ADFEM_HD GridIncident grid_incident(int s) {
  switch (s) {
    case 0: return {-1, -1, 1, 2};      // T1(i-1, j-1): node = TR
    case 1: return {-1, 0, 0, 2};       // T0(i-1, j)  : node = TL
    case 2: return {-1, 0, 1, 0};       // T1(i-1, j)  : node = TL
    case 3: return {0, -1, 0, 1};       // T0(i, j-1)  : node = BR
    case 4: return {0, -1, 1, 1};       // T1(i, j-1)  : node = BR
    default: return {0, 0, 0, 0};       // T0(i, j)    : node = BL
  }
}
// geometry and element id of triangle `tri` of cell (ci, cj); false when the cell is outside the mesh
// xy != nullptr: structured connectivity on mapped / jittered node positions — the corner positions come from the coordinate array ([node][2])
ADFEM_HD bool grid_triangle(const GridTri& gt, int heron, int ci, int cj, int tri, Geom<2>& G, long long& e, const double* xy = nullptr) {
  if (ci < 0 || ci >= gt.n || cj < 0 || cj >= gt.m) return false;
  double2 BL, BR, TL, TR;
  if (xy) {
    const double* p0 = xy + 2 * ((size_t)ci * (gt.m + 1) + cj);
    const double* p1 = xy + 2 * ((size_t)(ci + 1) * (gt.m + 1) + cj);
    BL = make_double2(ldg(p0), ldg(p0 + 1)); BR = make_double2(ldg(p0 + 2), ldg(p0 + 3));
    TL = make_double2(ldg(p1), ldg(p1 + 1)); TR = make_double2(ldg(p1 + 2), ldg(p1 + 3));
  } else {
    const double x0 = ldg(gt.xs + cj), x1 = ldg(gt.xs + cj + 1), y0 = ldg(gt.ys + ci), y1 = ldg(gt.ys + ci + 1);
    BL = make_double2(x0, y0); BR = make_double2(x1, y0); TL = make_double2(x0, y1); TR = make_double2(x1, y1);
  }
  if (tri == 0) geom_tri(BL, BR, TL, heron, G);
  else geom_tri(TL, BR, TR, heron, G);
  e = 2 * ((long long)ci * gt.m + cj) + tri;
  return true;
}

// transposed transfer / strain-energy term for node (i, j): acc[c], c < NC
template <int B, bool W>
ADFEM_HD void grid_scatter_node(const GridTri& gt, int heron, const QuadRule& rule, int g, int i, int j, const double* s, double* acc, const double* xy = nullptr) {
  using S = GpShape<2, 1, B>;
#pragma unroll
  for (int c = 0; c < S::NC; c++) acc[c] = 0.0;
#pragma unroll
  for (int t = 0; t < 6; t++) {
    const GridIncident inc = grid_incident(t);
    Geom<2> G; long long e;
    if (grid_triangle(gt, heron, i + inc.di, j + inc.dj, inc.tri, G, e, xy)) gp_scatter_elem<2, 1, B, W>(G, rule, g, inc.p, s + (size_t)e * g * S::NQ, acc);
  }
}

// matrix-free Laplace term for node (i, j)
ADFEM_HD double grid_laplace_term_node(const GridTri& gt, int heron, const QuadRule& rule, int g, int i, int j, const double* nu, const double* u,
                                       const double* xy = nullptr) {
  double acc = 0.0;
#pragma unroll
  for (int t = 0; t < 6; t++) {
    const GridIncident inc = grid_incident(t);
    const int ci = i + inc.di, cj = j + inc.dj;
    Geom<2> G; long long e;
    if (!grid_triangle(gt, heron, ci, cj, inc.tri, G, e, xy)) continue;
    const long long bl = (long long)ci * (gt.m + 1) + cj;        // node ids of the cell: BL, BR = BL + 1, TL = BL + m + 1, TR = TL + 1
    double ul[3];
    if (inc.tri == 0) { ul[0] = ldg(u + bl); ul[1] = ldg(u + bl + 1); ul[2] = ldg(u + bl + gt.m + 1); }
    else { ul[0] = ldg(u + bl + gt.m + 1); ul[1] = ldg(u + bl + 1); ul[2] = ldg(u + bl + gt.m + 2); }
    acc += laplace_term_elem<2, 1>(G, rule, g, inc.p, nu + (size_t)e * g, ul);
  }
  return acc;
}

}  // namespace adfem
