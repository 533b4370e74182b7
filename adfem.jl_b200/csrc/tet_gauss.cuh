// Scatter-type Gauss-point operators and the matrix-free Laplace term (ComputeLaplaceTermMfemT, deps/MFEM3/ComputeLaplaceTermMfem) on the structured
// tetrahedral grid `Mesh3(n, n, l, h)`: one thread per node visits its 8 or 32 incident tetrahedra (tet_grid_tables.h) in ascending element
// order with index arithmetic instead of the dof -> element adjacency, connectivity and coordinate arrays of the general kernels (gauss_ops.cuh)
// — ≈ 500 B of mesh-static data per node against ≈ 160 B of Gauss-point values.
// The per-Gauss-point arrays are indexed in the element's OWN vertex order, i.e. after MFEM's orientation fix (negative volume => swap local
// vertices 0 and 1), so the fix is replayed here on the generator's order.  Host + device bodies (tests/host_emul/).
#pragma once
#include "gauss_ops.cuh"
#include "tet_grid.cuh"

namespace adfem {

// incident tetrahedron t of node (i, j, k): geometry in the mesh's vertex order, the node's local index p, the element id and the node ids of the
// four vertices; false when the tetrahedron lies outside the grid
ADFEM_HD bool tg_incident(const GridTet& gt, int par, int t, int i, int j, int k, Geom<3>& G, int& p, long long& e, long long v[4]) {
  const TetGridTables& T = *gt.tab;
  const int ci = i + T.inc[par][t][0], cj = j + T.inc[par][t][1], ck = k + T.inc[par][t][2];
  if (ci < 0 || ci >= gt.n || cj < 0 || cj >= gt.n || ck < 0 || ck >= gt.l) return false;
  const long long n1 = gt.n + 1;
  double X[4][3];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int vi = i + T.voff[par][t][q][0], vj = j + T.voff[par][t][q][1], vk = k + T.voff[par][t][q][2];
    X[q][0] = ldg(gt.xs + vi); X[q][1] = ldg(gt.ys + vj); X[q][2] = ldg(gt.zs + vk);
    v[q] = ((long long)vk * n1 + vj) * n1 + vi;
  }
  p = T.inc[par][t][4];
  geom_tet(X, G);
  if (G.wscale < 0) {                                        // the orientation fix of the mesh tables: swap local vertices 0 and 1
#pragma unroll
    for (int c = 0; c < 3; c++) { const double x = X[0][c]; X[0][c] = X[1][c]; X[1][c] = x; }
    const long long w = v[0]; v[0] = v[1]; v[1] = w;
    p = p == 0 ? 1 : (p == 1 ? 0 : p);
    geom_tet(X, G);
  }
  e = 5 * (((long long)ci * gt.n + cj) * gt.l + ck) + T.inc[par][t][3];
  return true;
}

template <int B, bool W>
ADFEM_HD void tg_scatter_node(const GridTet& gt, const QuadRule& rule, int g, int i, int j, int k, const double* s, double* acc) {
  using S = GpShape<3, 1, B>;
#pragma unroll
  for (int c = 0; c < S::NC; c++) acc[c] = 0.0;
  const int par = (i + j + k) & 1;
  for (int t = 0; t < gt.tab->ninc[par]; t++) {
    Geom<3> G; int p; long long e, v[4];
    if (tg_incident(gt, par, t, i, j, k, G, p, e, v)) gp_scatter_elem<3, 1, B, W>(G, rule, g, p, s + (size_t)e * g * S::NQ, acc);
  }
}

ADFEM_HD double tg_laplace_term_node(const GridTet& gt, const QuadRule& rule, int g, int i, int j, int k, const double* nu, const double* u) {
  double acc = 0.0;
  const int par = (i + j + k) & 1;
  for (int t = 0; t < gt.tab->ninc[par]; t++) {
    Geom<3> G; int p; long long e, v[4];
    if (!tg_incident(gt, par, t, i, j, k, G, p, e, v)) continue;
    double ul[4];
#pragma unroll
    for (int q = 0; q < 4; q++) ul[q] = ldg(u + v[q]);
    acc += laplace_term_elem<3, 1>(G, rule, g, p, nu + (size_t)e * g, ul);
  }
  return acc;
}

}  // namespace adfem
