// =====================================================================================
// oracle/adfem_oracle.cpp — TEST INFRASTRUCTURE ONLY.
//
// A dependency-free CPU restatement (plain C++ loops, no MFEM / Eigen / TensorFlow) of the
// differentiable FEM-assembly path of kailaix/AdFem.jl.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library; it is the checker
// and the reported CPU baseline, never part of the product path (libadfem_cuda.so).
//
// PARITY STATUS: pinned by INDEPENDENT EXACT GOLDENS, not by reference-made files.  The reference cannot be built in
// this environment (needs TensorFlow-1.x headers, MFEM, Eigen, Julia) and the golden files its own
// tests read (fenics/A.txt, A2.txt, edges.txt of deps/MFEM/FemLaplace1/ftest.jl:6-33) are not shipped, so
// tests/golden/make_exact_goldens.py recomputes those matrices — and their mass / source / elasticity / tetrahedral
// siblings — in rational arithmetic (sympy) from the mesh arrays alone, on the reference's own test mesh Mesh(8,8,1/8)
// (P1, P2), a distorted renumbered triangulation, Mesh3(2,2,2,1/2) and a distorted tetrahedral cube; the oracle and the
// CUDA library both match them to 1e-12 (tests/test_exact_goldens.py), edge dofs matched through the `edges` table as
// ftest.jl:28-31 does.  Also pinned against the reference's tests (tests/test_oracle_known_answers.py):
//   * test/MFEM2.jl:7-27         Mesh(2,2,0.5): ngauss==24, area==0.125
//   * test/MFEM/MCore.jl:1-14    4-point segment rule values (lorder=6)
//   * deps/MFEM/FemSource1/ftest.jl:4-16   source(c) == mass(c)*1
//   * quadrature exactness degrees (triangle order 2/4, tetrahedron order 2/4 incl. the negative weight)
//   * polynomial exactness / K*1=0 / sum(M)=area / finite-difference gradient convergence
// What stays UNVERIFIABLE without MFEM itself: (1) the ORDER of the Gauss points inside an element (coefficient arrays are
// indexed e*g+k) — every operator is invariant to it as long as coefficients are sampled at `gauss_nodes`, which is the
// only way the reference's API produces them, and the exact goldens test exactly that; (2) the NUMBERS MFEM gives to
// edges (DSTable order) — observable only through `mesh.edges`, which is returned consistently with the P2 dofs.
// The COO slot ORDER (which duplicate comes first) follows the reference's loops by reading, not by execution.
//
// Third-party arithmetic that is NOT under /root/reference and is restated from its published
// algorithm (MFEM, version unpinned by the reference — src/ToolChain.jl:8 calls install_mfem()):
//   * orientation fix of Mesh::CheckElementOrientation (det<0 => swap local vertices 0,1)
//   * edge numbering of Mesh::GetElementToEdgeTable (DSTable first-appearance order)
//   * quadrature tables of IntegrationRules (intrules.cpp) for TRIANGLE / TETRAHEDRON / SEGMENT
//   * nodal H1 P1/P2 bases of H1_TriangleElement / H1_TetrahedronElement (closed form)
//
// Every function cites the reference file:line it follows.  Loop order, COO slot layout and
// data layout (one heap object per element, one heap temporary per Gauss point) follow the
// reference so that the timed CPU baseline is representative ("kind": "port").
// =====================================================================================
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <thread>
#include <vector>

typedef long long int64;

namespace {

// ------------------------------------------------------------------------------------
// Quadrature (MFEM fem/intrules.cpp — un-vendored; values restated from the published tables)
// ------------------------------------------------------------------------------------
struct IP { double x, y, z, w; };

void tri3(std::vector<IP>& r, double a, double w) {      // IntegrationRule::AddTriPoints3
  double b = 1. - 2. * a;
  r.push_back({a, a, 0, w}); r.push_back({a, b, 0, w}); r.push_back({b, a, 0, w});
}
void tri6(std::vector<IP>& r, double a, double b, double w) {   // AddTriPoints6
  double c = 1. - a - b;
  r.push_back({a, b, 0, w}); r.push_back({b, a, 0, w}); r.push_back({a, c, 0, w});
  r.push_back({c, a, 0, w}); r.push_back({b, c, 0, w}); r.push_back({c, b, 0, w});
}
std::vector<IP> tri_rule(int order) {     // IntegrationRules::TriangleIntegrationRule
  std::vector<IP> r;
  switch (order) {
    case 0: case 1: r.push_back({1. / 3., 1. / 3., 0, 0.5}); break;
    case 2: tri3(r, 1. / 6., 1. / 6.); break;
    case 3: r.push_back({1. / 3., 1. / 3., 0, -0.28125}); tri3(r, 0.2, 25. / 96.); break;
    case 4:
      tri3(r, 0.091576213509770743460, 0.054975871827660933819);
      tri3(r, 0.44594849091596488632, 0.11169079483900573285);
      break;
    case 5:
      r.push_back({1. / 3., 1. / 3., 0, 0.1125});
      tri3(r, 0.10128650732345633880, 0.062969590272413576298);
      tri3(r, 0.47014206410511508977, 0.066197076394253090369);
      break;
    case 6:
      tri3(r, 0.063089014491502228340, 0.025422453185103408460);
      tri3(r, 0.24928674517091042129, 0.058393137863189683013);
      tri6(r, 0.053145049844816947353, 0.31035245103378440542, 0.041425537809186787597);
      break;
    default: break;
  }
  return r;
}
void tet4(std::vector<IP>& r, double a, double w) {       // AddTetPoints4
  double b = 1. - 3. * a;
  r.push_back({a, a, a, w}); r.push_back({a, a, b, w}); r.push_back({a, b, a, w}); r.push_back({b, a, a, w});
}
void tet4b(std::vector<IP>& r, double b, double w) {      // AddTetPoints4b
  double a = (1. - b) / 3.;
  r.push_back({a, a, a, w}); r.push_back({a, a, b, w}); r.push_back({a, b, a, w}); r.push_back({b, a, a, w});
}
void tet6(std::vector<IP>& r, double a, double w) {       // AddTetPoints6
  double b = 0.5 - a;
  r.push_back({a, a, b, w}); r.push_back({a, b, a, w}); r.push_back({b, a, a, w});
  r.push_back({a, b, b, w}); r.push_back({b, a, b, w}); r.push_back({b, b, a, w});
}
std::vector<IP> tet_rule(int order) {     // IntegrationRules::TetrahedronIntegrationRule
  std::vector<IP> r;
  switch (order) {
    case 0: case 1: r.push_back({0.25, 0.25, 0.25, 1. / 6.}); break;
    case 2: tet4b(r, 0.58541019662496845446, 1. / 24.); break;
    case 3: r.push_back({0.25, 0.25, 0.25, -2. / 15.}); tet4b(r, 0.5, 0.075); break;
    case 4:
      tet4(r, 1. / 14., 343. / 45000.);
      r.push_back({0.25, 0.25, 0.25, -74. / 5625.});
      tet6(r, 0.10059642383320079500, 28. / 1125.);
      break;
    default: break;
  }
  return r;
}
// Gauss-Legendre on [0,1] with n = (order|1)/2 + 1 points, ascending
// (QuadratureFunctions1D::GaussLegendre; pinned by test/MFEM/MCore.jl:8-13 for order 6 -> 4 pts).
std::vector<IP> seg_rule(int order) {
  int real_order = order | 1;
  int n = real_order / 2 + 1;
  std::vector<IP> r(n);
  for (int i = 1; i <= (n + 1) / 2; i++) {
    double z = std::cos(M_PI * (i - 0.25) / (n + 0.5)), pp = 0, p1 = 0;
    for (int it = 0; it < 100; it++) {
      p1 = 1.0; double p2 = 0.0;
      for (int j = 1; j <= n; j++) { double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j; }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double dz = p1 / pp; z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    // re-evaluate derivative at the converged root
    { p1 = 1.0; double p2 = 0.0;
      for (int j = 1; j <= n; j++) { double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j; }
      pp = n * (z * p1 - p2) / (z * z - 1.0); }
    double x = 0.5 * (1.0 - z), w = 1.0 / ((1.0 - z * z) * pp * pp);   // weight on [0,1]
    r[i - 1] = {x, 0, 0, w};
    r[n - i] = {1.0 - x, 0, 0, w};
  }
  return r;
}

// ------------------------------------------------------------------------------------
// Edge table: Mesh::GetVertexToVertexTable + DSTable::Push (first-appearance numbering)
// ------------------------------------------------------------------------------------
struct EdgeTable {
  std::vector<int> head, next, col;   // per-row singly linked lists keyed by the smaller vertex
  EdgeTable(int nv) : head(nv, -1) {}
  int push(int a, int b) {
    int r = a <= b ? a : b, c = a <= b ? b : a;
    for (int n = head[r]; n >= 0; n = next[n]) if (col[n] == c) return n;
    int idx = (int)col.size();
    col.push_back(c); next.push_back(head[r]); head[r] = idx;
    return idx;
  }
  int size() const { return (int)col.size(); }
};

const int TRI_EDGES[3][2] = {{0, 1}, {1, 2}, {2, 0}};                                   // Geometry::Constants<TRIANGLE>::Edges
const int TET_EDGES[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};          // Geometry::Constants<TETRAHEDRON>::Edges

// ------------------------------------------------------------------------------------
// 2-D mesh tables — deps/MFEM/Common.h:17-63, Common.cpp:20-142
// ------------------------------------------------------------------------------------
struct Element2 {               // NNFEM_Element, Common.h:17-40 (one heap object per element, heap matrices)
  std::vector<double> h, hx, hy;   // elem_ndof x ngauss, (r,k) at r*ngauss+k
  std::vector<double> hs;          // 3 x ngauss
  std::vector<double> w;
  double area;
  double coord[6];
  int ngauss, ndof;
  int dof[6], node[3], edge[3];
};
struct Mesh2 {
  int nelem = 0, nnode = 0, ngauss = 0, ndof = 0, order = 0, degree = 0, elem_ndof = 0, lorder = 0;
  std::vector<double> GaussPts;    // ngauss x 2, column-major (Common.cpp:110-111, API.cpp:21-24)
  std::vector<Element2*> elements;
  void clear() { for (auto e : elements) delete e; elements.clear(); }
} mmesh;

double heron(const double* c) {    // Common.cpp:9-15
  double a = std::sqrt((c[0] - c[2]) * (c[0] - c[2]) + (c[1] - c[3]) * (c[1] - c[3]));
  double b = std::sqrt((c[4] - c[2]) * (c[4] - c[2]) + (c[5] - c[3]) * (c[5] - c[3]));
  double cc = std::sqrt((c[0] - c[4]) * (c[0] - c[4]) + (c[1] - c[5]) * (c[1] - c[5]));
  double s = (a + b + cc) / 2.0;
  return std::sqrt(s * (s - a) * (s - b) * (s - cc));
}

// Nodal H1 P1/P2 basis on the reference simplex in barycentric form; L = lambdas, gL = physical
// gradients of the lambdas (nv x dim).  Vertex functions first, then one per edge in geometry order.
template <int DIM>
void h1_basis(int degree, const double* L, const double (*gL)[DIM], double* phi, double (*gphi)[DIM]) {
  const int nv = DIM + 1;
  if (degree == 1) {
    for (int i = 0; i < nv; i++) { phi[i] = L[i]; for (int c = 0; c < DIM; c++) gphi[i][c] = gL[i][c]; }
    return;
  }
  for (int i = 0; i < nv; i++) {
    phi[i] = L[i] * (2.0 * L[i] - 1.0);
    for (int c = 0; c < DIM; c++) gphi[i][c] = (4.0 * L[i] - 1.0) * gL[i][c];
  }
  const int ne = DIM == 2 ? 3 : 6;
  for (int e = 0; e < ne; e++) {
    int a = DIM == 2 ? TRI_EDGES[e][0] : TET_EDGES[e][0];
    int b = DIM == 2 ? TRI_EDGES[e][1] : TET_EDGES[e][1];
    phi[nv + e] = 4.0 * L[a] * L[b];
    for (int c = 0; c < DIM; c++) gphi[nv + e][c] = 4.0 * (L[a] * gL[b][c] + L[b] * gL[a][c]);
  }
}

long long* mesh2_init(double* vertices, int num_vertices, int* element_indices, int num_elements,
                      int _order, int _lorder, int _degree, long long* nedges_ptr) {
  Mesh2& M = mmesh;
  M.order = _order; M.lorder = _lorder; M.degree = _degree;
  M.nelem = num_elements; M.nnode = num_vertices;
  std::vector<IP> rule = tri_rule(_order);
  const int g = (int)rule.size();
  M.ngauss = g * num_elements;
  M.elem_ndof = (_degree == 1) ? 3 : 6;                                   // Common.cpp:49
  // orientation fix: Mesh::CheckElementOrientation(true) via FinalizeTriMesh(1,0,true), Common.cpp:40
  std::vector<int> ev(element_indices, element_indices + 3 * (size_t)num_elements);
  for (int e = 0; e < num_elements; e++) {
    int* vi = &ev[3 * (size_t)e];
    const double *v0 = vertices + 3 * (size_t)vi[0], *v1 = vertices + 3 * (size_t)vi[1], *v2 = vertices + 3 * (size_t)vi[2];
    double det = (v1[0] - v0[0]) * (v2[1] - v0[1]) - (v1[1] - v0[1]) * (v2[0] - v0[0]);
    if (det < 0.0) { int t = vi[0]; vi[0] = vi[1]; vi[1] = t; }
  }
  // edge numbering: GetElementToEdgeTable
  EdgeTable et(num_vertices);
  std::vector<int> el_to_edge(3 * (size_t)num_elements);
  for (int e = 0; e < num_elements; e++)
    for (int j = 0; j < 3; j++)
      el_to_edge[3 * (size_t)e + j] = et.push(ev[3 * (size_t)e + TRI_EDGES[j][0]], ev[3 * (size_t)e + TRI_EDGES[j][1]]);
  int nedges = et.size();
  *nedges_ptr = nedges;
  M.ndof = (_degree == 1) ? M.nnode : (M.nnode + nedges);                 // Common.cpp:52
  M.GaussPts.assign(2 * (size_t)M.ngauss, 0.0);
  const int d = M.elem_ndof;
  size_t i_gp = 0;
  M.elements.reserve(num_elements);
  for (int e = 0; e < num_elements; e++) {                                 // Common.cpp:59-131
    Element2* el = new Element2;
    el->ngauss = g; el->ndof = d;
    el->h.assign(d * g, 0.0); el->hx.assign(d * g, 0.0); el->hy.assign(d * g, 0.0);
    el->hs.assign(3 * g, 0.0); el->w.assign(g, 0.0);
    for (int k = 0; k < 3; k++) { el->node[k] = ev[3 * (size_t)e + k]; el->dof[k] = el->node[k]; }
    for (int k = 0; k < 3; k++) { el->edge[k] = el_to_edge[3 * (size_t)e + k]; el->dof[k + 3] = el->edge[k] + M.nnode; }
    for (int k = 0; k < 3; k++) { el->coord[2 * k] = vertices[3 * (size_t)el->node[k]]; el->coord[2 * k + 1] = vertices[3 * (size_t)el->node[k] + 1]; }
    el->area = heron(el->coord);                                           // Common.cpp:83
    const double x1 = el->coord[0], y1 = el->coord[1], x2 = el->coord[2], y2 = el->coord[3], x3 = el->coord[4], y3 = el->coord[5];
    // inverse Jacobian of the affine map (ElementTransformation::InverseJacobian = adj/det)
    double det = (x2 - x1) * (y3 - y1) - (x3 - x1) * (y2 - y1);
    double gL[3][2];
    gL[1][0] = (y3 - y1) / det;  gL[1][1] = -(x3 - x1) / det;
    gL[2][0] = -(y2 - y1) / det; gL[2][1] = (x2 - x1) / det;
    gL[0][0] = -gL[1][0] - gL[2][0]; gL[0][1] = -gL[1][1] - gL[2][1];
    for (int i = 0; i < g; i++) {
      const IP& ip = rule[i];
      el->hs[0 * g + i] = 1 - ip.x - ip.y; el->hs[1 * g + i] = ip.x; el->hs[2 * g + i] = ip.y;   // Common.cpp:106-108
      M.GaussPts[i_gp] = x1 * el->hs[0 * g + i] + x2 * el->hs[1 * g + i] + x3 * el->hs[2 * g + i];
      M.GaussPts[M.ngauss + i_gp] = y1 * el->hs[0 * g + i] + y2 * el->hs[1 * g + i] + y3 * el->hs[2 * g + i];
      i_gp++;
      el->w[i] = ip.w * el->area / 0.5;                                     // Common.cpp:116
      double L[3] = {1 - ip.x - ip.y, ip.x, ip.y}, phi[6], gphi[6][2];
      h1_basis<2>(_degree, L, gL, phi, gphi);                               // CalcPhysShape / CalcPhysDShape, :119-120
      for (int k = 0; k < d; k++) { el->h[k * g + i] = phi[k]; el->hx[k * g + i] = gphi[k][0]; el->hy[k * g + i] = gphi[k][1]; }
    }
    M.elements.push_back(el);
  }
  long long* edges = (long long*)malloc(sizeof(long long) * 2 * (size_t)(nedges > 0 ? nedges : 1));   // Common.cpp:133-140
  // edge i joins row r (smaller vertex) and col[i]; recover r by walking the rows
  std::vector<int> row_of(nedges);
  for (int r = 0; r < num_vertices; r++) for (int n = et.head[r]; n >= 0; n = et.next[n]) row_of[n] = r;
  for (int i = 0; i < nedges; i++) { edges[i] = row_of[i] + 1; edges[nedges + i] = et.col[i] + 1; }
  return edges;
}

// ------------------------------------------------------------------------------------
// 3-D mesh tables — deps/MFEM3/Common.h:13-63, Common.cpp:9-148
// ------------------------------------------------------------------------------------
struct Element3 {               // NNFEM_Element3
  std::vector<double> h, hx, hy, hz, hs, w;
  double volume;
  double coord[12];
  int ngauss, ndof;
  int dof[10], node[4], edge[6];
};
struct Mesh3 {
  int nelem = 0, nnode = 0, ngauss = 0, ndof = 0, order = 0, degree = 0, elem_ndof = 0;
  std::vector<double> GaussPts;    // ngauss x 3 column-major
  std::vector<Element3*> elements;
  void clear() { for (auto e : elements) delete e; elements.clear(); }
} mmesh3;

double det3(const double* a, const double* b, const double* c) {   // columns/rows a,b,c
  return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
}

long long* mesh3_init(double* vertices, int num_vertices, int* element_indices, int num_elements,
                      int _order, int _degree, long long* nedges_ptr) {
  Mesh3& M = mmesh3;
  M.order = _order; M.degree = _degree; M.nelem = num_elements; M.nnode = num_vertices;
  std::vector<IP> rule = tet_rule(_order);
  const int g = (int)rule.size();
  M.ngauss = g * num_elements;
  std::vector<int> ev(element_indices, element_indices + 4 * (size_t)num_elements);
  for (int e = 0; e < num_elements; e++) {            // FinalizeTetMesh(1,0,true) -> CheckElementOrientation, Common.cpp:22
    int* vi = &ev[4 * (size_t)e];
    const double* v[4]; for (int j = 0; j < 4; j++) v[j] = vertices + 3 * (size_t)vi[j];
    double a[3], b[3], c[3];
    for (int k = 0; k < 3; k++) { a[k] = v[1][k] - v[0][k]; b[k] = v[2][k] - v[0][k]; c[k] = v[3][k] - v[0][k]; }
    if (det3(a, b, c) < 0.0) { int t = vi[0]; vi[0] = vi[1]; vi[1] = t; }
  }
  EdgeTable et(num_vertices);
  std::vector<int> el_to_edge(6 * (size_t)num_elements);
  for (int e = 0; e < num_elements; e++)
    for (int j = 0; j < 6; j++)
      el_to_edge[6 * (size_t)e + j] = et.push(ev[4 * (size_t)e + TET_EDGES[j][0]], ev[4 * (size_t)e + TET_EDGES[j][1]]);
  int nedges = et.size();
  *nedges_ptr = nedges;
  if (_degree == 1) { M.elem_ndof = 4; M.ndof = M.nnode; }              // Common.cpp:40-53
  else if (_degree == 2) { M.elem_ndof = 10; M.ndof = M.nnode + nedges; }
  else return nullptr;
  const int d = M.elem_ndof;
  M.GaussPts.assign(3 * (size_t)M.ngauss, 0.0);
  size_t i_gp = 0;
  M.elements.reserve(num_elements);
  for (int e = 0; e < num_elements; e++) {                               // Common.cpp:62-134
    Element3* el = new Element3;
    el->ngauss = g; el->ndof = d;
    el->h.assign(d * g, 0.0); el->hx.assign(d * g, 0.0); el->hy.assign(d * g, 0.0); el->hz.assign(d * g, 0.0);
    el->hs.assign(4 * g, 0.0); el->w.assign(g, 0.0);
    for (int k = 0; k < 4; k++) { el->node[k] = ev[4 * (size_t)e + k]; el->dof[k] = el->node[k]; }
    for (int k = 0; k < 6; k++) { el->edge[k] = el_to_edge[6 * (size_t)e + k]; el->dof[k + 4] = el->edge[k] + M.nnode; }
    for (int k = 0; k < 4; k++) for (int c = 0; c < 3; c++) el->coord[3 * k + c] = vertices[3 * (size_t)el->node[k] + c];
    const double* X = el->coord;
    double J[3][3];   // J[r][c] = d x_r / d xi_c
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) J[r][c] = X[3 * (c + 1) + r] - X[r];
    double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                 J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
    el->volume = det * (1. / 6.);                                         // Mesh::GetElementVolume, Common.cpp:88
    double inv[3][3];
    inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det; inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det; inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
    inv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det; inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
    inv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det; inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det; inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
    double gL[4][3];
    for (int c = 0; c < 3; c++) { for (int k = 0; k < 3; k++) gL[c + 1][k] = inv[c][k]; }
    for (int k = 0; k < 3; k++) gL[0][k] = -gL[1][k] - gL[2][k] - gL[3][k];
    for (int i = 0; i < g; i++) {
      const IP& ip = rule[i];
      double L[4] = {1 - ip.x - ip.y - ip.z, ip.x, ip.y, ip.z};
      for (int j = 0; j < 4; j++) el->hs[j * g + i] = L[j];               // degree-1 shapes, Common.cpp:95-103
      for (int c = 0; c < 3; c++)
        M.GaussPts[c * (size_t)M.ngauss + i_gp] = X[c] * L[0] + X[3 + c] * L[1] + X[6 + c] * L[2] + X[9 + c] * L[3];
      i_gp++;
      el->w[i] = ip.w * el->volume * 6.0;                                   // Common.cpp:117
      double phi[10], gphi[10][3];
      h1_basis<3>(_degree, L, gL, phi, gphi);
      for (int k = 0; k < d; k++) { el->h[k * g + i] = phi[k]; el->hx[k * g + i] = gphi[k][0]; el->hy[k * g + i] = gphi[k][1]; el->hz[k * g + i] = gphi[k][2]; }
    }
    M.elements.push_back(el);
  }
  long long* edges = (long long*)malloc(sizeof(long long) * 2 * (size_t)(nedges > 0 ? nedges : 1));
  std::vector<int> row_of(nedges);
  for (int r = 0; r < num_vertices; r++) for (int n = et.head[r]; n >= 0; n = et.next[n]) row_of[n] = r;
  for (int i = 0; i < nedges; i++) { edges[i] = row_of[i] + 1; edges[nedges + i] = et.col[i] + 1; }
  return edges;
}

const double pts[] = {(-1 / std::sqrt(3.0) + 1.0) / 2.0, (1 / std::sqrt(3.0) + 1.0) / 2.0};   // deps/FemStiffness1/UnivariateFemStiffness.h:5

}  // namespace

extern "C" {

// =====================================================================================
// Mesh API — deps/MFEM/API.cpp:3-66, deps/MFEM3/API.cpp:3-67
// =====================================================================================
long long* init_nnfem_mesh(double* vertices, int num_vertices, int* element_indices, int num_elements,
                           int order, int lorder, int degree, long long* nedges) {
  if (mmesh.elements.size() > 0) { printf("WARNING: Internal mesh is being overwritten!\n"); mmesh.clear(); }   // API.cpp:6-10
  return mesh2_init(vertices, num_vertices, element_indices, num_elements, order, lorder, degree, nedges);
}
int mfem_get_ngauss() { return mmesh.ngauss; }
void mfem_get_gauss(double* x, double* y) {
  memcpy(x, mmesh.GaussPts.data(), mmesh.ngauss * sizeof(double));
  memcpy(y, mmesh.GaussPts.data() + mmesh.ngauss, mmesh.ngauss * sizeof(double));
}
void mfem_get_gauss_weights(double* w) { size_t s = 0; for (auto el : mmesh.elements) for (int k = 0; k < el->ngauss; k++) w[s++] = el->w[k]; }
void mfem_get_area(double* a) { for (int i = 0; i < mmesh.nelem; i++) a[i] = mmesh.elements[i]->area; }
int mfem_get_elem_ndof() { return mmesh.elements[0]->ndof; }
int mfem_get_ndof() { return mmesh.ndof; }
void mfem_get_connectivity(long long* conn) { size_t p = 0; for (auto el : mmesh.elements) for (int k = 0; k < el->ndof; k++) conn[p++] = el->dof[k] + 1; }
void mfem_get_element_to_vertices(long long* elems) {
  for (int i = 0; i < mmesh.nelem; i++) for (int k = 0; k < 3; k++) elems[k * (size_t)mmesh.nelem + i] = mmesh.elements[i]->node[k] + 1;
}
int get_LineIntegralN() { return (int)seg_rule(mmesh.lorder).size(); }         // Common.cpp:350-352
void get_LineIntegralPnW(double* p, double* w) { auto r = seg_rule(mmesh.lorder); for (size_t i = 0; i < r.size(); i++) { p[i] = r[i].x; w[i] = r[i].w; } }
void oracle_segment_rule(int order, int* n, double* p, double* w) { auto r = seg_rule(order); *n = (int)r.size(); for (size_t i = 0; i < r.size(); i++) { p[i] = r[i].x; w[i] = r[i].w; } }
void oracle_free(void* p) { free(p); }

long long* init_nnfem_mesh3(double* vertices, int num_vertices, int* element_indices, int num_elements,
                            int order, int degree, long long* nedges) {
  if (mmesh3.elements.size() > 0) { printf("WARNING: Internal mesh is being overwritten!\n"); mmesh3.clear(); }
  return mesh3_init(vertices, num_vertices, element_indices, num_elements, order, degree, nedges);
}
int mfem_get_ngauss3() { return mmesh3.ngauss; }
void mfem_get_gauss3(double* x, double* y, double* z) {
  size_t G = mmesh3.ngauss;
  memcpy(x, mmesh3.GaussPts.data(), G * sizeof(double)); memcpy(y, mmesh3.GaussPts.data() + G, G * sizeof(double));
  memcpy(z, mmesh3.GaussPts.data() + 2 * G, G * sizeof(double));
}
void mfem_get_gauss_weights3(double* w) { size_t s = 0; for (auto el : mmesh3.elements) for (int k = 0; k < el->ngauss; k++) w[s++] = el->w[k]; }
void mfem_get_volume3(double* a) { for (int i = 0; i < mmesh3.nelem; i++) a[i] = mmesh3.elements[i]->volume; }
int mfem_get_elem_ndof3() { return mmesh3.elements[0]->ndof; }
int mfem_get_ndof3() { return mmesh3.ndof; }
void mfem_get_connectivity3(long long* conn) { size_t p = 0; for (auto el : mmesh3.elements) for (int k = 0; k < el->ndof; k++) conn[p++] = el->dof[k] + 1; }
void mfem_get_element_to_vertices3(long long* elems) {
  for (int i = 0; i < mmesh3.nelem; i++) for (int k = 0; k < 4; k++) elems[k * (size_t)mmesh3.nelem + i] = mmesh3.elements[i]->node[k] + 1;
}

// =====================================================================================
// 2-D ops
// =====================================================================================
// deps/MFEM/FemLaplace1/FemLaplaceScalar.h:3-26 (and :59-61)
// element range [e0, e1): the reference's loop body unchanged; slot / Gauss-point counters start where the full loop would be at e0
static void laplace2_forward_range(int e0, int e1, int64* indices, double* vv, const double* kappa) {
  int d = mmesh.elem_ndof;
  size_t s = 0, nz = 0;
  for (int i = 0; i < e0; i++) { s += mmesh.elements[i]->ngauss; nz += (size_t)mmesh.elements[i]->ngauss * d * d; }
  for (int i = e0; i < e1; i++) {
    Element2* elem = mmesh.elements[i];
    std::vector<double> D(d * 2);                       // Eigen::MatrixXd D(elem_ndof,2) per element (:9)
    for (int k = 0; k < elem->ngauss; k++) {
      for (int r = 0; r < d; r++) { D[2 * r] = elem->hx[r * elem->ngauss + k]; D[2 * r + 1] = elem->hy[r * elem->ngauss + k]; }
      std::vector<double> N(d * d);                     // heap temporary per Gauss point (:15)
      double c = kappa[s++], w = elem->w[k];
      for (int p = 0; p < d; p++) for (int q = 0; q < d; q++) N[p * d + q] = (D[2 * p] * D[2 * q] + D[2 * p + 1] * D[2 * q + 1]) * c * w;
      for (int p = 0; p < d; p++) for (int q = 0; q < d; q++) {
        indices[2 * nz] = elem->dof[p]; indices[2 * nz + 1] = elem->dof[q]; vv[nz] = N[p * d + q]; nz++;
      }
    }
  }
}
void FemLaplaceScalar_forward_Julia(int64* indices, double* vv, const double* kappa) { laplace2_forward_range(0, mmesh.nelem, indices, vv, kappa); }
// FemLaplaceScalar.h:28-54
static void laplace2_backward_range(int e0, int e1, double* grad_kappa, const double* grad_vv) {
  int d = mmesh.elem_ndof;
  size_t nz = 0, s = 0;
  for (int i = 0; i < e0; i++) { s += mmesh.elements[i]->ngauss; nz += (size_t)mmesh.elements[i]->ngauss * d * d; }
  for (int i = e0; i < e1; i++) {
    Element2* elem = mmesh.elements[i];
    std::vector<double> D(d * 2);
    for (int k = 0; k < elem->ngauss; k++) {
      for (int r = 0; r < d; r++) { D[2 * r] = elem->hx[r * elem->ngauss + k]; D[2 * r + 1] = elem->hy[r * elem->ngauss + k]; }
      std::vector<double> N(d * d);
      double w = elem->w[k];
      for (int p = 0; p < d; p++) for (int q = 0; q < d; q++) N[p * d + q] = (D[2 * p] * D[2 * q] + D[2 * p + 1] * D[2 * q + 1]) * w;
      double v = 0.0;
      for (int p = 0; p < d; p++) for (int q = 0; q < d; q++) { v += grad_vv[nz] * N[p * d + q]; nz++; }
      grad_kappa[s++] = v;
    }
  }
}
void oracle_FemLaplaceScalar_backward(double* grad_kappa, const double* grad_vv) { laplace2_backward_range(0, mmesh.nelem, grad_kappa, grad_vv); }
// bench.py's CPU arm only: the same two loops split over `nthreads` contiguous element blocks (the reference itself has no threading; every
// slot is written by exactly one thread, so the outputs are bit-identical to the serial call).
void oracle_FemLaplaceScalar_forward_backward_mt(int nthreads, int64* indices, double* vv, const double* kappa, double* grad_kappa, const double* grad_vv) {
  if (nthreads < 1) nthreads = 1;
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) {
    const int e0 = (int)((long long)mmesh.nelem * t / nthreads), e1 = (int)((long long)mmesh.nelem * (t + 1) / nthreads);
    th.emplace_back([=] { laplace2_forward_range(e0, e1, indices, vv, kappa); laplace2_backward_range(e0, e1, grad_kappa, grad_vv); });
  }
  for (auto& x : th) x.join();
}
// FemLaplaceScalar.h:65-92 — dense column-major ngauss x N Jacobian
void pcl_FemLaplaceScalar_Jacobian(double* H) {
  size_t s = 0, nz = 0;
  int d = mmesh.elem_ndof;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    for (int k = 0; k < elem->ngauss; k++) {
      double w = elem->w[k];
      for (int p = 0; p < d; p++) for (int q = 0; q < d; q++) {
        double n = (elem->hx[p * elem->ngauss + k] * elem->hx[q * elem->ngauss + k] + elem->hy[p * elem->ngauss + k] * elem->hy[q * elem->ngauss + k]) * w;
        H[s + nz * (size_t)mmesh.ngauss] = n; nz++;
      }
      s++;
    }
  }
}

// deps/MFEM/ComputeFemStiffnessMatrixMfem/ComputeFemStiffnessMatrixMfem.h:4-48 (and :84-88)
void ComputeFemStiffnessMatrixMfem_forward_Julia(int64* indices, double* vv, const double* hmat) {
  int d = mmesh.elem_ndof, D2 = 2 * d;
  std::vector<double> B(3 * D2);
  double K[9];
  size_t k = 0, k0 = 0;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    std::fill(B.begin(), B.end(), 0.0);
    for (int j = 0; j < elem->ngauss; j++) {
      for (int r = 0; r < d; r++) {
        double hx = elem->hx[r * elem->ngauss + j], hy = elem->hy[r * elem->ngauss + j];
        B[0 * D2 + r] = hx; B[1 * D2 + r + d] = hy; B[2 * D2 + r] = hy; B[2 * D2 + r + d] = hx;
      }
      for (int p = 0; p < 3; p++) for (int q = 0; q < 3; q++) K[3 * p + q] = hmat[k0++];
      std::vector<double> KB(3 * D2), NN(D2 * D2);      // heap temporaries per Gauss point (:29)
      for (int p = 0; p < 3; p++) for (int s = 0; s < D2; s++) KB[p * D2 + s] = K[3 * p] * B[s] + K[3 * p + 1] * B[D2 + s] + K[3 * p + 2] * B[2 * D2 + s];
      for (int l = 0; l < D2; l++) for (int s = 0; s < D2; s++)
        NN[l * D2 + s] = (B[l] * KB[s] + B[D2 + l] * KB[D2 + s] + B[2 * D2 + l] * KB[2 * D2 + s]) * elem->w[j];
      std::vector<int> dofs(D2);
      for (int p = 0; p < d; p++) { dofs[p] = elem->dof[p]; dofs[p + d] = elem->dof[p] + mmesh.ndof; }
      for (int l = 0; l < D2; l++) for (int s = 0; s < D2; s++) {
        indices[2 * k] = dofs[l]; indices[2 * k + 1] = dofs[s]; vv[k] = NN[l * D2 + s]; k++;
      }
    }
  }
}
// ComputeFemStiffnessMatrixMfem.h:50-81
void oracle_ComputeFemStiffnessMatrixMfem_backward(double* grad_hmat, const double* grad_vv) {
  int d = mmesh.elem_ndof, D2 = 2 * d;
  std::vector<double> B(3 * D2), K(D2 * D2);
  size_t k0 = 0, k = 0;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    std::fill(B.begin(), B.end(), 0.0);
    for (int j = 0; j < elem->ngauss; j++) {
      for (int r = 0; r < d; r++) {
        double hx = elem->hx[r * elem->ngauss + j], hy = elem->hy[r * elem->ngauss + j];
        B[0 * D2 + r] = hx; B[1 * D2 + r + d] = hy; B[2 * D2 + r] = hy; B[2 * D2 + r + d] = hx;
      }
      for (int l = 0; l < D2; l++) for (int s = 0; s < D2; s++) K[l * D2 + s] = grad_vv[k++];
      std::vector<double> BK(3 * D2);
      for (int p = 0; p < 3; p++) for (int s = 0; s < D2; s++) { double a = 0; for (int l = 0; l < D2; l++) a += B[p * D2 + l] * K[l * D2 + s]; BK[p * D2 + s] = a; }
      for (int p = 0; p < 3; p++) for (int q = 0; q < 3; q++) {
        double a = 0; for (int s = 0; s < D2; s++) a += BK[p * D2 + s] * B[q * D2 + s];
        grad_hmat[k0++] = a * elem->w[j];
      }
    }
  }
}

// deps/MFEM/ComputeFemMassMatrix1/ComputeFemMassMatrixMfem.h:4-29
void oracle_ComputeFemMassMatrix1_forward(int64* indices, double* vv, const double* rho) {
  int d = mmesh.elem_ndof;
  size_t k = 0, k0 = 0;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    for (int j = 0; j < elem->ngauss; j++) {
      std::vector<double> NN(d * d);
      double c = rho[k0++], w = elem->w[j];
      for (int l = 0; l < d; l++) for (int s = 0; s < d; s++) NN[l * d + s] = elem->h[l * elem->ngauss + j] * elem->h[s * elem->ngauss + j] * c * w;
      for (int l = 0; l < d; l++) for (int s = 0; s < d; s++) {
        indices[2 * k] = elem->dof[l]; indices[2 * k + 1] = elem->dof[s]; vv[k] = NN[l * d + s]; k++;
      }
    }
  }
}
// ComputeFemMassMatrixMfem.h:31-56 (the op shell zero-fills grad_rho first, .cpp:139)
void oracle_ComputeFemMassMatrix1_backward(double* grad_rho, const double* grad_vv) {
  int d = mmesh.elem_ndof;
  size_t k = 0, k0 = 0;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    for (int j = 0; j < elem->ngauss; j++) {
      double w = elem->w[j];
      grad_rho[k0] = 0.0;
      for (int l = 0; l < d; l++) for (int s = 0; s < d; s++) {
        grad_rho[k0] += elem->h[l * elem->ngauss + j] * elem->h[s * elem->ngauss + j] * w * grad_vv[k]; k++;
      }
      k0++;
    }
  }
}

// deps/MFEM/FemSource1/FemSourceScalar.h:4-15 (rhs must be pre-zeroed by the caller, .cpp:73)
void FemSourceScalar_forward_Julia(double* rhs, const double* f) {
  size_t k = 0;
  int d = mmesh.elem_ndof;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    for (int j = 0; j < elem->ngauss; j++) {
      for (int r = 0; r < d; r++) rhs[elem->dof[r]] += f[k] * elem->h[r * elem->ngauss + j] * elem->w[j];
      k++;
    }
  }
}
// FemSourceScalar.h:17-32
void oracle_FemSourceScalar_backward(double* grad_f, const double* grad_rhs) {
  size_t k = 0;
  int d = mmesh.elem_ndof;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    for (int j = 0; j < elem->ngauss; j++) {
      grad_f[k] = 0.0;
      for (int r = 0; r < d; r++) grad_f[k] += elem->h[r * elem->ngauss + j] * elem->w[j] * grad_rhs[elem->dof[r]];
      k++;
    }
  }
}

// =====================================================================================
// Gauss-point operators and matrix-free terms on the 2-D tables.  Like the reference bodies these
// ACCUMULATE where the reference writes `+=` (the callers zero-fill) and assign where it writes `=`.
// =====================================================================================
// deps/MFEM/FemToGaussPoints/FemToGaussPointsMfem.h:6-18
void oracle_FemToGaussPointsMfem_forward(double* out, const double* u) {
  size_t k = 0;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    const int g = elem->ngauss;
    for (int j = 0; j < g; j++) {
      out[k] = u[elem->node[0]] * elem->hs[0 * g + j] + u[elem->node[1]] * elem->hs[1 * g + j] + u[elem->node[2]] * elem->hs[2 * g + j];
      k++;
    }
  }
}
// FemToGaussPointsMfem.h:20-35
void oracle_FemToGaussPointsMfem_backward(double* grad_u, const double* grad_out) {
  size_t k = 0;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    const int g = elem->ngauss;
    for (int j = 0; j < g; j++) {
      grad_u[elem->node[0]] += grad_out[k] * elem->hs[0 * g + j];
      grad_u[elem->node[1]] += grad_out[k] * elem->hs[1 * g + j];
      grad_u[elem->node[2]] += grad_out[k] * elem->hs[2 * g + j];
      k++;
    }
  }
}
// deps/MFEM/DofToGaussPoints/DofToGaussPointsMfem.h:6-18
void oracle_DofToGaussPointsMfem_forward(double* out, const double* u) {
  size_t k = 0;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    const int g = elem->ngauss;
    for (int j = 0; j < g; j++) {
      out[k] = 0.0;
      for (int r = 0; r < elem->ndof; r++) out[k] += u[elem->dof[r]] * elem->h[r * g + j];
      k++;
    }
  }
}
// DofToGaussPointsMfem.h:20-34
void oracle_DofToGaussPointsMfem_backward(double* grad_u, const double* grad_out) {
  size_t k = 0;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    const int g = elem->ngauss;
    for (int j = 0; j < g; j++) {
      for (int r = 0; r < elem->ndof; r++) grad_u[elem->dof[r]] += grad_out[k] * elem->h[r * g + j];
      k++;
    }
  }
}
// deps/MFEM/FemGrad/FemGradMfem.h:8-23
void oracle_FemGradMfem_forward(double* out, const double* u) {
  size_t k = 0;
  const int d = mmesh.elem_ndof;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    const int g = elem->ngauss;
    for (int j = 0; j < g; j++) {
      out[k] = 0.0;
      for (int r = 0; r < d; r++) out[k] += u[elem->dof[r]] * elem->hx[r * g + j];
      k++;
      out[k] = 0.0;
      for (int r = 0; r < d; r++) out[k] += u[elem->dof[r]] * elem->hy[r * g + j];
      k++;
    }
  }
}
// FemGradMfem.h:25-42
void oracle_FemGradMfem_backward(double* grad_u, const double* grad_out) {
  size_t k = 0;
  const int d = mmesh.elem_ndof;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    const int g = elem->ngauss;
    for (int j = 0; j < g; j++) {
      for (int r = 0; r < d; r++) grad_u[elem->dof[r]] += grad_out[k] * elem->hx[r * g + j];
      k++;
      for (int r = 0; r < d; r++) grad_u[elem->dof[r]] += grad_out[k] * elem->hy[r * g + j];
      k++;
    }
  }
}
// deps/MFEM/EvalStrainOnGaussPtsMfem/EvalStrainOnGaussPts.h:4-16
void oracle_EvalStrainOnGaussPts_forward(double* epsilon, const double* u) {
  const int d = mmesh.elem_ndof;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    const int g = elem->ngauss;
    for (int j = 0; j < g; j++)
      for (int p = 0; p < d; p++) {
        epsilon[((size_t)i * g + j) * 3] += elem->hx[p * g + j] * u[elem->dof[p]];
        epsilon[((size_t)i * g + j) * 3 + 1] += elem->hy[p * g + j] * u[elem->dof[p] + mmesh.ndof];
        epsilon[((size_t)i * g + j) * 3 + 2] += elem->hy[p * g + j] * u[elem->dof[p]] + elem->hx[p * g + j] * u[elem->dof[p] + mmesh.ndof];
      }
  }
}
// EvalStrainOnGaussPts.h:18-29
void oracle_EvalStrainOnGaussPts_backward(double* grad_u, const double* grad_epsilon) {
  const int d = mmesh.elem_ndof;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    const int g = elem->ngauss;
    for (int j = 0; j < g; j++)
      for (int p = 0; p < d; p++) {
        const double* ge = grad_epsilon + ((size_t)i * g + j) * 3;
        grad_u[elem->dof[p]] += elem->hx[p * g + j] * ge[0] + elem->hy[p * g + j] * ge[2];
        grad_u[elem->dof[p] + mmesh.ndof] += elem->hy[p * g + j] * ge[1] + elem->hx[p * g + j] * ge[2];
      }
  }
}
// deps/MFEM/ComputeStrainEnergyTermMfem/ComputeStrainEnergyTermMfem.h:4-19
void oracle_ComputeStrainEnergyTermMfem_forward(double* out, const double* sigma) {
  const int d = mmesh.elem_ndof;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    const int g = elem->ngauss;
    for (int j = 0; j < g; j++) {
      const double s11 = sigma[3 * ((size_t)i * g + j)], s22 = sigma[3 * ((size_t)i * g + j) + 1], s12 = sigma[3 * ((size_t)i * g + j) + 2];
      for (int p = 0; p < d; p++) {
        out[elem->dof[p]] += s11 * elem->hx[p * g + j] * elem->w[j];
        out[elem->dof[p] + mmesh.ndof] += s22 * elem->hy[p * g + j] * elem->w[j];
        out[elem->dof[p]] += s12 * elem->hy[p * g + j] * elem->w[j];
        out[elem->dof[p] + mmesh.ndof] += s12 * elem->hx[p * g + j] * elem->w[j];
      }
    }
  }
}
// ComputeStrainEnergyTermMfem.h:21-35
void oracle_ComputeStrainEnergyTermMfem_backward(double* grad_sigma, const double* grad_out) {
  const int d = mmesh.elem_ndof;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    const int g = elem->ngauss;
    for (int j = 0; j < g; j++)
      for (int p = 0; p < d; p++) {
        double* gs = grad_sigma + 3 * ((size_t)i * g + j);
        gs[0] += elem->hx[p * g + j] * elem->w[j] * grad_out[elem->dof[p]];
        gs[1] += elem->hy[p * g + j] * elem->w[j] * grad_out[elem->dof[p] + mmesh.ndof];
        gs[2] += elem->hy[p * g + j] * elem->w[j] * grad_out[elem->dof[p]] + elem->hx[p * g + j] * elem->w[j] * grad_out[elem->dof[p] + mmesh.ndof];
      }
  }
}
// deps/MFEM/ComputeLaplaceTermMfem/ComputeLaplaceTermMfem.h:4-17
void oracle_ComputeLaplaceTermMfem_forward(double* out, const double* nu, const double* u) {
  const int d = mmesh.elem_ndof;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    const int g = elem->ngauss;
    for (int j = 0; j < g; j++) {
      const double nu_val = nu[(size_t)i * g + j];
      for (int p = 0; p < d; p++)
        for (int q = 0; q < d; q++)
          out[elem->dof[p]] += nu_val * (elem->hx[p * g + j] * elem->hx[q * g + j] * u[elem->dof[q]] * elem->w[j] +
                                         elem->hy[p * g + j] * elem->hy[q * g + j] * u[elem->dof[q]] * elem->w[j]);
    }
  }
}
// ComputeLaplaceTermMfem.h:19-39
void oracle_ComputeLaplaceTermMfem_backward(double* grad_nu, double* grad_u, const double* grad_out, const double* nu, const double* u) {
  const int d = mmesh.elem_ndof;
  for (int i = 0; i < mmesh.nelem; i++) {
    Element2* elem = mmesh.elements[i];
    const int g = elem->ngauss;
    for (int j = 0; j < g; j++) {
      const double nu_val = nu[(size_t)i * g + j];
      for (int p = 0; p < d; p++)
        for (int q = 0; q < d; q++) {
          grad_nu[(size_t)i * g + j] += grad_out[elem->dof[p]] * (elem->hx[p * g + j] * elem->hx[q * g + j] * u[elem->dof[q]] * elem->w[j] +
                                                                  elem->hy[p * g + j] * elem->hy[q * g + j] * u[elem->dof[q]] * elem->w[j]);
          grad_u[elem->dof[q]] += grad_out[elem->dof[p]] * nu_val * elem->hx[p * g + j] * elem->hx[q * g + j] * elem->w[j];
          grad_u[elem->dof[q]] += grad_out[elem->dof[p]] * nu_val * elem->hy[p * g + j] * elem->hy[q * g + j] * elem->w[j];
        }
    }
  }
}
// deps/MFEM/PlaneStrainAndStress/PlaneStrainAndStress.h:5-19 (mode 0) and :46-60 (mode 1)
void oracle_PlaneMatrix_forward(double* out, const double* E, const double* nu, int N, int mode) {
  for (int i = 0; i < N; i++) {
    if (mode == 0) {
      double s = E[i] * (1 - nu[i]) / (1 + nu[i]) / (1 - 2 * nu[i]);
      for (int k = 0; k < 9; k++) out[9 * i + k] = (k % 4 == 0) ? s : s * nu[i] / (1 - nu[i]);
    } else {
      double s = E[i] / (1 + nu[i]) / (1 - 2 * nu[i]);
      out[9 * i] = s * (1 - nu[i]); out[9 * i + 1] = s * nu[i]; out[9 * i + 2] = 0.0;
      out[9 * i + 3] = s * nu[i]; out[9 * i + 4] = s * (1 - nu[i]); out[9 * i + 5] = 0.0;
      out[9 * i + 6] = 0.0; out[9 * i + 7] = 0.0; out[9 * i + 8] = s * (1 - 2 * nu[i]) / 2.0;
    }
  }
}
// PlaneStrainAndStress.h:21-44, 62-78.  The reference runs a reverse-mode tape (`had`, not vendored) over the forward expressions;
// restated here as forward-mode dual numbers over the SAME expressions (one pass per input), which yields the same derivatives.
struct Dual { double v, d; };
static inline Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
static inline Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
static inline Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
static inline Dual operator/(Dual a, Dual b) { return {a.v / b.v, (a.d * b.v - a.v * b.d) / (b.v * b.v)}; }
static inline Dual cst(double c) { return {c, 0.0}; }
static Dual plane_L(Dual Ei, Dual nui, const double* go, int mode) {
  if (mode == 0) {
    Dual s = Ei * (cst(1) - nui) / (cst(1) + nui) / (cst(1) - cst(2) * nui);
    Dual t = s * nui / (cst(1) - nui);
    Dual L = cst(0);
    for (int k = 0; k < 9; k++) L = L + cst(go[k]) * ((k % 4 == 0) ? s : t);
    return L;
  }
  Dual s = Ei / (cst(1) + nui) / (cst(1) - cst(2) * nui);
  return cst(go[0]) * s * (cst(1) - nui) + cst(go[1]) * s * nui + cst(go[3]) * s * nui + cst(go[4]) * s * (cst(1) - nui) +
         cst(go[8]) * s * (cst(1) - cst(2) * nui) / cst(2.0);
}
void oracle_PlaneMatrix_backward(double* grad_nu, double* grad_E, const double* grad_out, const double* E, const double* nu, int N, int mode) {
  for (int i = 0; i < N; i++) {
    grad_E[i] = plane_L({E[i], 1.0}, {nu[i], 0.0}, grad_out + 9 * i, mode).d;
    grad_nu[i] = plane_L({E[i], 0.0}, {nu[i], 1.0}, grad_out + 9 * i, mode).d;
  }
}

// =====================================================================================
// ImposeDirichlet — deps/MFEM/ImposeDirichlet/ImposeDirichlet.h:27-93
// =====================================================================================
// forward, two calls like the op shell (ImposeDirichlet.cpp:104-128): count, then copy
struct DirichletState { std::vector<int64> ii, jj; std::vector<double> vv, rhs; } g_dir;
long long oracle_ImposeDirichlet_forward(const int64* indices, const double* v_ipt, const int64* bd, const double* rhs_ipt,
                                         const double* bdval, int N, int bdN, int sN) {
  g_dir = DirichletState();
  std::map<int64, double> bdMap;
  for (int i = 0; i < N; i++) g_dir.rhs.push_back(rhs_ipt[i]);
  for (int i = 0; i < bdN; i++) bdMap[bd[i] - 1] = bdval[i];
  for (int k = 0; k < sN; k++) {
    int64 i = indices[2 * k], j = indices[2 * k + 1];
    if (bdMap.count(i) == 0 && bdMap.count(j) == 0) { g_dir.ii.push_back(i); g_dir.jj.push_back(j); g_dir.vv.push_back(v_ipt[k]); }
    if (bdMap.count(i) == 0 && bdMap.count(j) > 0) g_dir.rhs[i] = g_dir.rhs[i] - v_ipt[k] * bdMap[j];
  }
  for (auto& it : bdMap) { g_dir.ii.push_back(it.first); g_dir.jj.push_back(it.first); g_dir.vv.push_back(1.0); g_dir.rhs[it.first] = it.second; }
  return (long long)g_dir.vv.size();
}
void oracle_ImposeDirichlet_copy(int64* oindices, double* ov, double* orhs) {      // ImposeDirichlet.h:53-60
  for (size_t i = 0; i < g_dir.rhs.size(); i++) orhs[i] = g_dir.rhs[i];
  for (size_t i = 0; i < g_dir.ii.size(); i++) { oindices[2 * i] = g_dir.ii[i]; oindices[2 * i + 1] = g_dir.jj[i]; ov[i] = g_dir.vv[i]; }
}
// DirichletBd (deps/DirichletBd/DirichletBd.h:8-60): two calls like the op shell (count, then copy).  bdset = bd and bd + (m+1)(n+1);
// a boundary dof's column in the second matrix is its 1-based position in that concatenated list (std::map assignment: last wins).
struct DirichletBdState { std::vector<int64> ii1, jj1, ii2, jj2; std::vector<double> vv1, vv2; } g_dbd;
void oracle_DirichletBd_forward(const int64* ii, const int64* jj, const double* vv, int N, const int* bd, int bdn, int m, int n, long long* n1, long long* n2) {
  g_dbd = DirichletBdState();
  const int off = (m + 1) * (n + 1);
  std::set<int> bdset;
  std::map<int, int> position;
  for (int half = 0, k = 1; half < 2; half++)
    for (int i = 0; i < bdn; i++) { const int dof = bd[i] + half * off; bdset.insert(dof); position[dof] = k++; }      // :17-29
  for (int s = 0; s < N; s++) {                                                                                         // :31-41
    const bool rb = bdset.count((int)ii[s]) > 0, cb = bdset.count((int)jj[s]) > 0;
    if (!rb && !cb) { g_dbd.ii1.push_back(ii[s]); g_dbd.jj1.push_back(jj[s]); g_dbd.vv1.push_back(vv[s]); }
    if (!rb && cb) { g_dbd.ii2.push_back(ii[s]); g_dbd.jj2.push_back(position[(int)jj[s]]); g_dbd.vv2.push_back(vv[s]); }
  }
  for (int dof : bdset) { g_dbd.ii1.push_back(dof); g_dbd.jj1.push_back(dof); g_dbd.vv1.push_back(1.0); }               // :42-44
  *n1 = (long long)g_dbd.vv1.size(); *n2 = (long long)g_dbd.vv2.size();
}
void oracle_DirichletBd_copy(int64* ii1, int64* jj1, double* vv1, int64* ii2, int64* jj2, double* vv2) {                // fill(), :47-92
  for (size_t i = 0; i < g_dbd.vv1.size(); i++) { ii1[i] = g_dbd.ii1[i]; jj1[i] = g_dbd.jj1[i]; vv1[i] = g_dbd.vv1[i]; }
  for (size_t i = 0; i < g_dbd.vv2.size(); i++) { ii2[i] = g_dbd.ii2[i]; jj2[i] = g_dbd.jj2[i]; vv2[i] = g_dbd.vv2[i]; }
}
void oracle_DirichletBd_backward(double* grad_vv, const int64* ii, const int64* jj, const double* grad_vv1, const double* grad_vv2, int N, const int* bd,
                                 int bdn, int m, int n) {                                                                  // :96-112
  std::set<int> bdset(bd, bd + bdn);
  for (int i = 0; i < bdn; i++) bdset.insert(bd[i] + (m + 1) * (n + 1));
  size_t k1 = 0, k2 = 0;
  for (int s = 0; s < N; s++) {
    grad_vv[s] = 0.0;
    const bool rb = bdset.count((int)ii[s]) > 0, cb = bdset.count((int)jj[s]) > 0;
    if (!rb && !cb) grad_vv[s] += grad_vv1[k1++];
    if (!rb && cb) grad_vv[s] += grad_vv2[k2++];
  }
}
// ImposeDirichlet.h:63-93; outputs are zero-filled first like the Grad op shell (.cpp:229-232)
void oracle_ImposeDirichlet_backward(double* grad_vv_ipt, double* grad_rhs_ipt, double* grad_bdval, const double* grad_vv,
                                     const double* grad_rhs, const int64* indices, const double* v_ipt, const int64* bd,
                                     const double* bdval, int N, int bdN, int sN) {
  for (int k = 0; k < sN; k++) grad_vv_ipt[k] = 0.0;
  for (int i = 0; i < N; i++) grad_rhs_ipt[i] = 0.0;
  for (int i = 0; i < bdN; i++) grad_bdval[i] = 0.0;
  std::map<int64, double> bdMap;
  std::map<int64, int64> bdMapIdx;
  for (int i = 0; i < bdN; i++) bdMap[bd[i] - 1] = bdval[i];
  for (int i = 0; i < bdN; i++) bdMapIdx[bd[i] - 1] = i;
  size_t z = 0;
  for (int k = 0; k < sN; k++) {
    int64 i = indices[2 * k], j = indices[2 * k + 1];
    if (bdMap.count(i) == 0 && bdMap.count(j) == 0) grad_vv_ipt[k] = grad_vv[z++];
    if (bdMap.count(i) == 0 && bdMap.count(j) > 0) {
      grad_vv_ipt[k] -= bdMap[j] * grad_rhs[i];
      grad_bdval[bdMapIdx[j]] -= v_ipt[k] * grad_rhs[i];
    }
  }
  for (int i = 0; i < N; i++) if (bdMap.count(i) == 0) grad_rhs_ipt[i] = grad_rhs[i];
  for (auto& it : bdMapIdx) grad_bdval[it.second] += grad_rhs[it.first];
}

// =====================================================================================
// 3-D ops
// =====================================================================================
// deps/MFEM3/FemLaplace1/FemLaplaceScalarT.h:3-27 (and :61)
void FemLaplaceScalarT_forward_Julia(int64* indices, double* vv, const double* kappa) {
  size_t s = 0, nz = 0;
  int d = mmesh3.elem_ndof;
  for (int i = 0; i < mmesh3.nelem; i++) {
    Element3* elem = mmesh3.elements[i];
    std::vector<double> D(d * 3);
    int g = elem->ngauss;
    for (int k = 0; k < g; k++) {
      for (int r = 0; r < d; r++) { D[3 * r] = elem->hx[r * g + k]; D[3 * r + 1] = elem->hy[r * g + k]; D[3 * r + 2] = elem->hz[r * g + k]; }
      std::vector<double> N(d * d);
      double c = kappa[s++], w = elem->w[k];
      for (int p = 0; p < d; p++) for (int q = 0; q < d; q++)
        N[p * d + q] = (D[3 * p] * D[3 * q] + D[3 * p + 1] * D[3 * q + 1] + D[3 * p + 2] * D[3 * q + 2]) * c * w;
      for (int p = 0; p < d; p++) for (int q = 0; q < d; q++) {
        indices[2 * nz] = elem->dof[p]; indices[2 * nz + 1] = elem->dof[q]; vv[nz] = N[p * d + q]; nz++;
      }
    }
  }
}
// FemLaplaceScalarT.h:29-56
void oracle_FemLaplaceScalarT_backward(double* grad_kappa, const double* grad_vv) {
  size_t nz = 0, s = 0;
  int d = mmesh3.elem_ndof;
  for (int i = 0; i < mmesh3.nelem; i++) {
    Element3* elem = mmesh3.elements[i];
    std::vector<double> D(d * 3);
    int g = elem->ngauss;
    for (int k = 0; k < g; k++) {
      for (int r = 0; r < d; r++) { D[3 * r] = elem->hx[r * g + k]; D[3 * r + 1] = elem->hy[r * g + k]; D[3 * r + 2] = elem->hz[r * g + k]; }
      double w = elem->w[k], v = 0.0;
      for (int p = 0; p < d; p++) for (int q = 0; q < d; q++) {
        v += grad_vv[nz] * ((D[3 * p] * D[3 * q] + D[3 * p + 1] * D[3 * q + 1] + D[3 * p + 2] * D[3 * q + 2]) * w); nz++;
      }
      grad_kappa[s++] = v;
    }
  }
}
// deps/MFEM3/ComputeFemMassMatrixMfem3/ComputeFemMassMatrixMfemT.h:4-27 — ONE slot per (e,p,q), N = nelem*d^2
void oracle_ComputeFemMassMatrixMfemT_forward(int64* indices, double* vv, const double* rho) {
  int d = mmesh3.elem_ndof;
  size_t nz = 0;
  for (int i = 0; i < mmesh3.nelem; i++) {
    Element3* elem = mmesh3.elements[i];
    int g = elem->ngauss;
    for (int p = 0; p < d; p++) for (int q = 0; q < d; q++) {
      double s = 0.0;
      for (int k = 0; k < g; k++) s += elem->h[p * g + k] * elem->h[q * g + k] * elem->w[k] * rho[k];
      indices[2 * nz] = elem->dof[p]; indices[2 * nz + 1] = elem->dof[q]; vv[nz] = s; nz++;
    }
    rho += g;
  }
}
// EXTENSION (Q5): the reference's Grad op body is empty (ComputeFemMassMatrixMfemT.cpp:136-140); this is
// the mathematically implied adjoint of the forward above.  Parity unpinned by the reference.
void oracle_ComputeFemMassMatrixMfemT_backward(double* grad_rho, const double* grad_vv) {
  int d = mmesh3.elem_ndof;
  size_t base = 0, s = 0;
  for (int i = 0; i < mmesh3.nelem; i++) {
    Element3* elem = mmesh3.elements[i];
    int g = elem->ngauss;
    for (int k = 0; k < g; k++) {
      double v = 0.0;
      for (int p = 0; p < d; p++) for (int q = 0; q < d; q++) v += elem->h[p * g + k] * elem->h[q * g + k] * elem->w[k] * grad_vv[base + p * d + q];
      grad_rho[s++] = v;
    }
    base += d * d;
  }
}
// deps/MFEM3/FemSource/FemSourceScalarT.h:4-15 / :17-32
void FemSourceScalarT_forward_Julia(double* rhs, const double* f) {
  size_t k = 0;
  int d = mmesh3.elem_ndof;
  for (int i = 0; i < mmesh3.nelem; i++) {
    Element3* elem = mmesh3.elements[i];
    for (int j = 0; j < elem->ngauss; j++) {
      for (int r = 0; r < d; r++) rhs[elem->dof[r]] += f[k] * elem->h[r * elem->ngauss + j] * elem->w[j];
      k++;
    }
  }
}
void oracle_FemSourceScalarT_backward(double* grad_f, const double* grad_rhs) {
  size_t k = 0;
  int d = mmesh3.elem_ndof;
  for (int i = 0; i < mmesh3.nelem; i++) {
    Element3* elem = mmesh3.elements[i];
    for (int j = 0; j < elem->ngauss; j++) {
      grad_f[k] = 0.0;
      for (int r = 0; r < d; r++) grad_f[k] += elem->h[r * elem->ngauss + j] * elem->w[j] * grad_rhs[elem->dof[r]];
      k++;
    }
  }
}
// EXTENSION (N2): 3-D elasticity stiffness, not in the reference.  ComputeFemStiffnessMatrixMfem.h:4-81
// generalised to Voigt order [xx, yy, zz, yz, xz, xy], H = 6x6 row-major per Gauss point, local dofs
// component-blocked [dof, dof+ndof, dof+2ndof] (layout of deps/MFEM3/PMLElasticMfem3/ComputePmlElasticTermT.h:48-50).
static void build_B3(const Element3* elem, int j, int d, std::vector<double>& B) {
  int D3 = 3 * d, g = elem->ngauss;
  std::fill(B.begin(), B.end(), 0.0);
  for (int r = 0; r < d; r++) {
    double hx = elem->hx[r * g + j], hy = elem->hy[r * g + j], hz = elem->hz[r * g + j];
    B[0 * D3 + r] = hx; B[1 * D3 + r + d] = hy; B[2 * D3 + r + 2 * d] = hz;
    B[3 * D3 + r + d] = hz; B[3 * D3 + r + 2 * d] = hy;
    B[4 * D3 + r] = hz;     B[4 * D3 + r + 2 * d] = hx;
    B[5 * D3 + r] = hy;     B[5 * D3 + r + d] = hx;
  }
}
void oracle_ComputeFemStiffnessMatrixMfemT_forward(int64* indices, double* vv, const double* hmat) {
  int d = mmesh3.elem_ndof, D3 = 3 * d;
  std::vector<double> B(6 * D3), KB(6 * D3);
  size_t k = 0, k0 = 0;
  for (int i = 0; i < mmesh3.nelem; i++) {
    Element3* elem = mmesh3.elements[i];
    for (int j = 0; j < elem->ngauss; j++) {
      build_B3(elem, j, d, B);
      const double* K = hmat + k0; k0 += 36;
      for (int p = 0; p < 6; p++) for (int s = 0; s < D3; s++) { double a = 0; for (int q = 0; q < 6; q++) a += K[6 * p + q] * B[q * D3 + s]; KB[p * D3 + s] = a; }
      for (int l = 0; l < D3; l++) for (int s = 0; s < D3; s++) {
        double a = 0; for (int p = 0; p < 6; p++) a += B[p * D3 + l] * KB[p * D3 + s];
        int cl = l / d, cs = s / d;
        indices[2 * k] = elem->dof[l % d] + (int64)cl * mmesh3.ndof; indices[2 * k + 1] = elem->dof[s % d] + (int64)cs * mmesh3.ndof;
        vv[k] = a * elem->w[j]; k++;
      }
    }
  }
}
void oracle_ComputeFemStiffnessMatrixMfemT_backward(double* grad_hmat, const double* grad_vv) {
  int d = mmesh3.elem_ndof, D3 = 3 * d;
  std::vector<double> B(6 * D3), BK(6 * D3);
  size_t k = 0, k0 = 0;
  for (int i = 0; i < mmesh3.nelem; i++) {
    Element3* elem = mmesh3.elements[i];
    for (int j = 0; j < elem->ngauss; j++) {
      build_B3(elem, j, d, B);
      const double* G = grad_vv + k; k += (size_t)D3 * D3;
      for (int p = 0; p < 6; p++) for (int s = 0; s < D3; s++) { double a = 0; for (int l = 0; l < D3; l++) a += B[p * D3 + l] * G[l * D3 + s]; BK[p * D3 + s] = a; }
      for (int p = 0; p < 6; p++) for (int q = 0; q < 6; q++) {
        double a = 0; for (int s = 0; s < D3; s++) a += BK[p * D3 + s] * B[q * D3 + s];
        grad_hmat[k0++] = a * elem->w[j];
      }
    }
  }
}

// deps/MFEM3/ComputeLaplaceTermMfem/ComputeLaplaceTermMfemT.h:4-22
void oracle_ComputeLaplaceTermMfemT_forward(double* out, const double* nu, const double* u) {
  const int d = mmesh3.elem_ndof;
  for (int i = 0; i < mmesh3.nelem; i++) {
    Element3* elem = mmesh3.elements[i];
    const int g = elem->ngauss;
    for (int j = 0; j < g; j++) {
      const double nu_val = nu[(size_t)i * g + j];
      for (int p = 0; p < d; p++)
        for (int q = 0; q < d; q++)
          out[elem->dof[p]] += nu_val * (elem->hx[p * g + j] * elem->hx[q * g + j] * u[elem->dof[q]] * elem->w[j] +
                                         elem->hy[p * g + j] * elem->hy[q * g + j] * u[elem->dof[q]] * elem->w[j] +
                                         elem->hz[p * g + j] * elem->hz[q * g + j] * u[elem->dof[q]] * elem->w[j]);
    }
  }
}
// ComputeLaplaceTermMfemT.h:24-46
void oracle_ComputeLaplaceTermMfemT_backward(double* grad_nu, double* grad_u, const double* grad_out, const double* nu, const double* u) {
  const int d = mmesh3.elem_ndof;
  for (int i = 0; i < mmesh3.nelem; i++) {
    Element3* elem = mmesh3.elements[i];
    const int g = elem->ngauss;
    for (int j = 0; j < g; j++) {
      const double nu_val = nu[(size_t)i * g + j];
      for (int p = 0; p < d; p++)
        for (int q = 0; q < d; q++) {
          grad_nu[(size_t)i * g + j] += grad_out[elem->dof[p]] * (elem->hx[p * g + j] * elem->hx[q * g + j] * u[elem->dof[q]] * elem->w[j] +
                                                                  elem->hy[p * g + j] * elem->hy[q * g + j] * u[elem->dof[q]] * elem->w[j] +
                                                                  elem->hz[p * g + j] * elem->hz[q * g + j] * u[elem->dof[q]] * elem->w[j]);
          grad_u[elem->dof[q]] += grad_out[elem->dof[p]] * nu_val * elem->hx[p * g + j] * elem->hx[q * g + j] * elem->w[j];
          grad_u[elem->dof[q]] += grad_out[elem->dof[p]] * nu_val * elem->hy[p * g + j] * elem->hy[q * g + j] * elem->w[j];
          grad_u[elem->dof[q]] += grad_out[elem->dof[p]] * nu_val * elem->hz[p * g + j] * elem->hz[q * g + j] * elem->w[j];
        }
    }
  }
}
// Extension (no 3-D twin in the reference): shape tables of Element3 at the Gauss points, exported so that the tests can check the 3-D
// Gauss-point gathers of libadfem_cuda against the oracle's own h / hx / hy / hz (layout [r*g + k] per element, elements concatenated).
void oracle_shape_tables3(double* h, double* hx, double* hy, double* hz) {
  size_t o = 0;
  for (int i = 0; i < mmesh3.nelem; i++) {
    Element3* elem = mmesh3.elements[i];
    const size_t n = (size_t)elem->ndof * elem->ngauss;
    for (size_t k = 0; k < n; k++) { h[o + k] = elem->h[k]; hx[o + k] = elem->hx[k]; hy[o + k] = elem->hy[k]; hz[o + k] = elem->hz[k]; }
    o += n;
  }
}

// =====================================================================================
// Structured-grid Q1 ops
// =====================================================================================
static void make_Bs2(double h, double Bs[4][2][4]) {     // UnivariateFemStiffness.h:20-26 : Bs[2*ej+ei], xi=pts[ei], eta=pts[ej]
  for (int ei = 0; ei < 2; ei++) for (int ej = 0; ej < 2; ej++) {
    double xi = pts[ei], eta = pts[ej];
    double r0[4] = {-1 / h * (1 - eta), 1 / h * (1 - eta), -1 / h * eta, 1 / h * eta};
    double r1[4] = {-1 / h * (1 - xi), -1 / h * xi, 1 / h * (1 - xi), 1 / h * xi};
    for (int c = 0; c < 4; c++) { Bs[2 * ej + ei][0][c] = r0[c]; Bs[2 * ej + ei][1][c] = r1[c]; }
  }
}
// deps/FemStiffness1/UnivariateFemStiffness.h:7-77 (rank3=1, Forward_UFS) and :129-196 (rank3=0, Forward2); 1-based ii/jj
void oracle_UnivariateFemStiffness_forward(int64* ii, int64* jj, double* vv, const double* hmat, int m, int n, double h, int rank3) {
  double Bs[4][2][4]; make_Bs2(h, Bs);
  size_t z = 0;
  for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) {
    int idx[4] = {j * (m + 1) + i, j * (m + 1) + i + 1, (j + 1) * (m + 1) + i, (j + 1) * (m + 1) + i + 1};
    for (int ei = 0; ei < 2; ei++) for (int ej = 0; ej < 2; ej++) {
      const double* K = rank3 ? hmat + 16 * ((size_t)i + (size_t)j * m) + 4 * (ei + ej * 2) : hmat;
      const double (*B)[4] = Bs[2 * ej + ei];
      for (int p = 0; p < 4; p++) for (int q = 0; q < 4; q++) {
        double KBq0 = K[0] * B[0][q] + K[1] * B[1][q], KBq1 = K[2] * B[0][q] + K[3] * B[1][q];
        ii[z] = idx[p] + 1; jj[z] = idx[q] + 1; vv[z] = (B[0][p] * KBq0 + B[1][p] * KBq1) * 0.25 * h * h; z++;
      }
    }
  }
}
// UnivariateFemStiffness.h:80-124 (rank3=1) / :199-247 (rank3=0, accumulates into 4 numbers)
void oracle_UnivariateFemStiffness_backward(double* grad_hmat, const double* grad_vv, int m, int n, double h, int rank3) {
  double Bs[4][2][4]; make_Bs2(h, Bs);
  if (!rank3) for (int t = 0; t < 4; t++) grad_hmat[t] = 0.0;
  size_t k = 0;
  for (int ei = 0; ei < m; ei++) for (int ej = 0; ej < n; ej++)
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) {
      size_t ids = 16 * ((size_t)ei + (size_t)ej * m) + 4 * (i + j * 2);
      const double (*B)[4] = Bs[2 * j + i];
      double dK[2][2] = {{0, 0}, {0, 0}};
      for (int p = 0; p < 4; p++) for (int q = 0; q < 4; q++) {
        double gO = grad_vv[k++];
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) dK[a][b] += B[a][p] * gO * B[b][q];
      }
      for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) {
        double v = dK[a][b] * 0.25 * h * h;
        if (rank3) grad_hmat[ids + 2 * a + b] = v; else grad_hmat[2 * a + b] += v;
      }
    }
}
static void make_B3x8(double h, double xi, double eta, double B[3][8]) {     // FemStiffness.h:24-26
  double r0[4] = {-1 / h * (1 - eta), 1 / h * (1 - eta), -1 / h * eta, 1 / h * eta};
  double r1[4] = {-1 / h * (1 - xi), -1 / h * xi, 1 / h * (1 - xi), 1 / h * xi};
  for (int c = 0; c < 4; c++) { B[0][c] = r0[c]; B[0][c + 4] = 0; B[1][c] = 0; B[1][c + 4] = r1[c]; B[2][c] = r1[c]; B[2][c + 4] = r0[c]; }
}
static void quad_idx8(int i, int j, int m, int n, int idx[8]) {
  int b[4] = {j * (m + 1) + i, j * (m + 1) + i + 1, (j + 1) * (m + 1) + i, (j + 1) * (m + 1) + i + 1};
  for (int c = 0; c < 4; c++) { idx[c] = b[c]; idx[c + 4] = b[c] + (m + 1) * (n + 1); }
}
// deps/FemStiffness/FemStiffness.h:7-70 — constant H read COLUMN-major (:18-20), one Omega for every cell
void oracle_FemStiffness_forward(int64* ii, int64* jj, double* vv, const double* hmat, int m, int n, double h) {
  double K[3][3], Omega[8][8];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) K[i][j] = hmat[j * 3 + i];
  for (int p = 0; p < 8; p++) for (int q = 0; q < 8; q++) Omega[p][q] = 0;
  for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) {
    double B[3][8]; make_B3x8(h, pts[i], pts[j], B);
    for (int p = 0; p < 8; p++) for (int q = 0; q < 8; q++) {
      double a = 0; for (int r = 0; r < 3; r++) for (int s = 0; s < 3; s++) a += B[r][p] * K[r][s] * B[s][q];
      Omega[p][q] += a * 0.25 * h * h;
    }
  }
  size_t z = 0;
  for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) {
    int idx[8]; quad_idx8(i, j, m, n, idx);
    for (int p = 0; p < 8; p++) for (int q = 0; q < 8; q++) { ii[z] = idx[p] + 1; jj[z] = idx[q] + 1; vv[z] = Omega[p][q]; z++; }
  }
}
// FemStiffness.h:73-109
void oracle_FemStiffness_backward(double* grad_hmat, const double* grad_vv, int m, int n, double h) {
  for (int i = 0; i < 9; i++) grad_hmat[i] = 0.0;
  size_t k = 0;
  double Bq[4][3][8];
  for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) make_B3x8(h, pts[i], pts[j], Bq[2 * i + j]);
  for (int ci = 0; ci < m; ci++) for (int cj = 0; cj < n; cj++) {
    const double* dO = grad_vv + k; k += 64;
    double dK[3][3] = {{0}};
    for (int t = 0; t < 4; t++)
      for (int r = 0; r < 3; r++) for (int s = 0; s < 3; s++) {
        double a = 0; for (int p = 0; p < 8; p++) for (int q = 0; q < 8; q++) a += Bq[t][r][p] * dO[p * 8 + q] * Bq[t][s][q];
        dK[r][s] += a * 0.25 * h * h;
      }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) grad_hmat[j * 3 + i] += dK[i][j];
  }
}
// deps/SpatialFemStiffness/SpatialFemStiffness.h:7-81 — per-Gauss H row-major at hmat[36*elem+9*k], k=2q+p,
// paired with Bs[k] built as (xi=pts[k/2], eta=pts[k%2])  (quirk Q6, :18-26 vs :34-41)
void oracle_SpatialFemStiffness_forward(int64* ii, int64* jj, double* vv, const double* hmat, int m, int n, double h) {
  double Bs[4][3][8];
  { int k = 0; for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) make_B3x8(h, pts[i], pts[j], Bs[k++]); }
  size_t z = 0;
  for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) {
    size_t elem = (size_t)j * m + i;
    int idx[8]; quad_idx8(i, j, m, n, idx);
    for (int p = 0; p < 2; p++) for (int q = 0; q < 2; q++) {
      int k = 2 * q + p;
      const double* K = hmat + 36 * elem + 9 * k;
      for (int r = 0; r < 8; r++) for (int s = 0; s < 8; s++) {
        double a = 0; for (int x = 0; x < 3; x++) for (int y = 0; y < 3; y++) a += Bs[k][x][r] * K[3 * x + y] * Bs[k][y][s];
        ii[z] = idx[r] + 1; jj[z] = idx[s] + 1; vv[z] = a * 0.25 * h * h; z++;
      }
    }
  }
}
// SpatialFemStiffness.h:85-133
void oracle_SpatialFemStiffness_backward(double* grad_hmat, const double* grad_vv, int m, int n, double h) {
  double Bs[4][3][8];
  { int k = 0; for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) make_B3x8(h, pts[i], pts[j], Bs[k++]); }
  size_t rs_ = 0;
  for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) {
    size_t elem = (size_t)j * m + i;
    for (int p = 0; p < 2; p++) for (int q = 0; q < 2; q++) {
      int k = 2 * q + p;
      const double* G = grad_vv + rs_; rs_ += 64;       // G[r*8+s] = d loss / d Omega(r,s)
      for (int x = 0; x < 3; x++) for (int y = 0; y < 3; y++) {
        double a = 0; for (int r = 0; r < 8; r++) for (int s = 0; s < 8; s++) a += Bs[k][x][r] * G[r * 8 + s] * Bs[k][y][s];
        grad_hmat[36 * elem + 9 * k + 3 * x + y] = a * 0.25 * h * h;
      }
    }
  }
}
// deps/SpatialVaryingTangentElastic/SpatialVaryingTangentElastic.h:1-31 / :35-70
void oracle_SVT_forward(double* hmat, const double* mu, long long m, long long n, int type) {
  size_t k = 0, off = 4 * (size_t)m * n;
  for (size_t i = 0; i < off; i++) {
    if (type == 1) { hmat[k++] = mu[i]; hmat[k++] = 0.; hmat[k++] = 0.; hmat[k++] = mu[i]; }
    else if (type == 2) { hmat[k++] = mu[i]; hmat[k++] = 0.; hmat[k++] = 0.; hmat[k++] = mu[i + off]; }
    else { hmat[k++] = mu[i]; hmat[k++] = mu[i + 2 * off]; hmat[k++] = mu[i + 2 * off]; hmat[k++] = mu[i + off]; }
  }
}
void oracle_SVT_backward(double* grad_mu, const double* grad_hmat, long long m, long long n, int type) {
  size_t k = 0, off = 4 * (size_t)m * n;
  for (size_t i = 0; i < off; i++) {
    if (type == 1) grad_mu[i] = grad_hmat[k] + grad_hmat[k + 3];
    else if (type == 2) { grad_mu[i] = grad_hmat[k]; grad_mu[i + off] = grad_hmat[k + 3]; }
    else { grad_mu[i] = grad_hmat[k]; grad_mu[i + off] = grad_hmat[k + 3]; grad_mu[i + 2 * off] = grad_hmat[k + 1] + grad_hmat[k + 2]; }
    k += 4;
  }
}

// =====================================================================================
// Structured-grid Q1 scalar siblings; ii / jj 0-based as these ops emit them
// =====================================================================================
// the four B0^T B0 h^2/4 matrices, k = 2q + p with xi = pts[p], eta = pts[q] — deps/FemLaplace/FemLaplace.h:13-23
static void make_laplace_locals(double h, double B[4][4][4]) {
  for (int q = 0; q < 2; q++) for (int p = 0; p < 2; p++) {
    const int k = q * 2 + p;
    const double xi = pts[p], eta = pts[q];
    const double B0[2][4] = {{-1 / h * (1 - eta), 1 / h * (1 - eta), -1 / h * eta, 1 / h * eta},
                             {-1 / h * (1 - xi), -1 / h * xi, 1 / h * (1 - xi), 1 / h * xi}};
    for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) B[k][a][b] = (B0[0][a] * B0[0][b] + B0[1][a] * B0[1][b]) * 0.25 * h * h;
  }
}
// the four A A^T h^2/4 matrices — deps/FemMass/FemMass.h:12-20
static void make_mass_locals(double h, double Me[4][4][4]) {
  for (int q = 0; q < 2; q++) for (int p = 0; p < 2; p++) {
    const double xi = pts[p], eta = pts[q];
    const double A[4] = {(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta};
    for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) Me[q * 2 + p][a][b] = A[a] * A[b] * 0.25 * h * h;
  }
}
// deps/FemLaplace/FemLaplace.h:11-49
void oracle_FemLaplace_forward(int64* ii, int64* jj, double* vv, const double* K, int m, int n, double h) {
  double B[4][4][4]; make_laplace_locals(h, B);
  size_t k_gauss = 0, k = 0;
  for (int j = 0; j < n; j++) for (int i = 0; i < m; i++) {
    const int idx[4] = {j * (m + 1) + i, j * (m + 1) + i + 1, (j + 1) * (m + 1) + i, (j + 1) * (m + 1) + i + 1};
    for (int q = 0; q < 2; q++) for (int p = 0; p < 2; p++) {
      const double kv = K[k_gauss++];
      for (int i_ = 0; i_ < 4; i_++) for (int j_ = 0; j_ < 4; j_++) { ii[k] = idx[i_]; jj[k] = idx[j_]; vv[k] = kv * B[2 * q + p][i_][j_]; k++; }
    }
  }
}
// FemLaplace.h:51-82 (accumulates; the op shell zero-fills)
void oracle_FemLaplace_backward(double* grad_K, const double* grad_vv, int m, int n, double h) {
  double B[4][4][4]; make_laplace_locals(h, B);
  size_t k_gauss = 0, k = 0;
  for (int j = 0; j < n; j++) for (int i = 0; i < m; i++)
    for (int q = 0; q < 2; q++) for (int p = 0; p < 2; p++) {
      for (int i_ = 0; i_ < 4; i_++) for (int j_ = 0; j_ < 4; j_++) { grad_K[k_gauss] += B[2 * q + p][i_][j_] * grad_vv[k]; k++; }
      k_gauss++;
    }
}
// deps/FemMass/FemMass.h:10-43
void oracle_FemMass_forward(int64* ii, int64* jj, double* vv, const double* rho, int m, int n, double h) {
  double Me[4][4][4]; make_mass_locals(h, Me);
  size_t k = 0;
  for (int j = 0; j < n; j++) for (int i = 0; i < m; i++) {
    const size_t elem_idx = (size_t)j * m + i;
    const int idx[4] = {j * (m + 1) + i, j * (m + 1) + i + 1, (j + 1) * (m + 1) + i, (j + 1) * (m + 1) + i + 1};
    for (int q = 0; q < 2; q++) for (int p = 0; p < 2; p++) {
      const double rho_ = rho[elem_idx * 4 + 2 * q + p];
      for (int i_ = 0; i_ < 4; i_++) for (int j_ = 0; j_ < 4; j_++) { ii[k] = idx[i_]; jj[k] = idx[j_]; vv[k] = rho_ * Me[2 * q + p][i_][j_]; k++; }
    }
  }
}
// FemMass.h:45-79
void oracle_FemMass_backward(double* grad_rho, const double* grad_vv, int m, int n, double h) {
  double Me[4][4][4]; make_mass_locals(h, Me);
  size_t k = 0;
  for (int j = 0; j < n; j++) for (int i = 0; i < m; i++) {
    const size_t elem_idx = (size_t)j * m + i;
    for (int q = 0; q < 2; q++) for (int p = 0; p < 2; p++)
      for (int i_ = 0; i_ < 4; i_++) for (int j_ = 0; j_ < 4; j_++) { grad_rho[elem_idx * 4 + 2 * q + p] += grad_vv[k] * Me[2 * q + p][i_][j_]; k++; }
  }
}
// deps/FemSource/FemSource.h:8-26 (cell loop i outer, j inner; accumulates into a zero-filled rhs)
void oracle_FemSource_forward(double* rhs, const double* f, int m, int n, double h) {
  for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) {
    const size_t idx = (size_t)j * m + i;
    for (int p = 0; p < 2; p++) for (int q = 0; q < 2; q++) {
      const double xi = pts[p], eta = pts[q];
      const size_t k = idx * 4 + 2 * q + p;
      const double val1 = f[k] * h * h * 0.25;
      rhs[j * (m + 1) + i] += val1 * (1 - xi) * (1 - eta);
      rhs[j * (m + 1) + i + 1] += val1 * xi * (1 - eta);
      rhs[(j + 1) * (m + 1) + i] += val1 * (1 - xi) * eta;
      rhs[(j + 1) * (m + 1) + i + 1] += val1 * xi * eta;
    }
  }
}
// FemSource.h:30-47
void oracle_FemSource_backward(double* grad_f, const double* grad_rhs, int m, int n, double h) {
  for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) {
    const size_t idx = (size_t)j * m + i;
    for (int p = 0; p < 2; p++) for (int q = 0; q < 2; q++) {
      const double xi = pts[p], eta = pts[q];
      const size_t k = idx * 4 + 2 * q + p;
      grad_f[k] += h * h * 0.25 * (1 - xi) * (1 - eta) * grad_rhs[j * (m + 1) + i];
      grad_f[k] += h * h * 0.25 * xi * (1 - eta) * grad_rhs[j * (m + 1) + i + 1];
      grad_f[k] += h * h * 0.25 * (1 - xi) * eta * grad_rhs[(j + 1) * (m + 1) + i];
      grad_f[k] += h * h * 0.25 * xi * eta * grad_rhs[(j + 1) * (m + 1) + i + 1];
    }
  }
}

}  // extern "C"
