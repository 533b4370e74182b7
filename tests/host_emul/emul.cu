// tests/host_emul/emul.cu — TEST INFRASTRUCTURE ONLY (never linked into libadfem_cuda.so, never imported by the package).
//
// The kernel bodies of adfem.jl_b200/csrc/gauss_ops.cuh are `__host__ __device__`; this file compiles them for the HOST and runs
// them in plain loops that mirror the kernels of gauss_ops.cu (same indexing of inputs and outputs), so that the arithmetic the GPU
// executes can be checked against the oracle on a machine without a GPU (tests/test_host_emulation.py).  It is a checker of the
// device code, not a CPU path of the product: the product's entry points still fail without a CUDA device.
#include <cstdint>

#include "gauss_ops.cuh"
#include "quad_ops.cuh"
#include "grid_elast.cuh"
#include "grid_gauss.cuh"
#include "tet_grid.cuh"
#include "row_gather.cuh"
#include "tet_gauss.cuh"
#include "tet_scalar.cuh"

using namespace adfem;

// Lanes of a warp phase run one after the other here; on the GPU they run together, so a phase must not depend on the order.  emul_set_reverse(1)
// makes every phase visit its lanes 31..0 instead of 0..31: the tests require bit-identical results for both orders.
static int g_reverse = 0;
#define EMUL_LANES(lane) for (int lane##_i = 0, lane = g_reverse ? 31 : 0; lane##_i < 32; lane##_i++, lane = g_reverse ? 31 - lane##_i : lane##_i)

namespace {

DevMesh make_mesh(int dim, int degree, int order, int nv, int ne, int ndof, const double* coords, const int* verts, const int* conn, bool& ok) {
  DevMesh m{};
  m.dim = dim; m.ne = ne; m.nv = nv; m.ndof = ndof; m.heron = 1;
  m.d = dim == 2 ? (degree == 1 ? 3 : 6) : (degree == 1 ? 4 : 10);
  m.coords = coords; m.verts = verts; m.conn = conn;
  ok = dim == 2 ? triangle_rule(order, m.rule) : tetrahedron_rule(order, m.rule);
  m.g = m.rule.n;
  return m;
}

template <int DIM, int DEG, int B, bool W>
void run_gather(const DevMesh& m, const double* in, double* out) {
  for (int e = 0; e < m.ne; e++) gp_gather_body<DIM, DEG, B, W>(m, e, in, out);      // k_gp_gather
}
template <int DIM, int DEG, int B, bool W>
void run_scatter(const DevMesh& m, const long long* ap, const int* ae, const uint8_t* al, const double* in, double* out) {
  constexpr int NC = GpShape<DIM, DEG, B>::NC;
  const int nrows = B == GB_P1SHAPE ? m.nv : m.ndof;                                   // launch_gp_scatter
  for (int r = 0; r < nrows; r++) {                                                    // k_gp_scatter
    double acc[NC];
    gp_scatter_row<DIM, DEG, B, W>(m, ap, ae, al, r, in, acc);
    for (int c = 0; c < NC; c++) out[r + (size_t)c * nrows] = acc[c];
  }
}
template <int DIM, int DEG>
int run_kind(const DevMesh& m, const long long* ap, const int* ae, const uint8_t* al, int basis, bool weighted, bool to_gauss, const double* in,
             double* out) {
  if (to_gauss) {
    switch (basis) {
      case GB_P1SHAPE: run_gather<DIM, DEG, GB_P1SHAPE, false>(m, in, out); return 0;
      case GB_SHAPE: run_gather<DIM, DEG, GB_SHAPE, false>(m, in, out); return 0;
      case GB_GRAD: run_gather<DIM, DEG, GB_GRAD, false>(m, in, out); return 0;
      case GB_STRAIN: if (weighted) run_gather<DIM, DEG, GB_STRAIN, true>(m, in, out); else run_gather<DIM, DEG, GB_STRAIN, false>(m, in, out); return 0;
    }
    return 1;
  }
  switch (basis) {
    case GB_P1SHAPE: run_scatter<DIM, DEG, GB_P1SHAPE, false>(m, ap, ae, al, in, out); return 0;
    case GB_SHAPE: run_scatter<DIM, DEG, GB_SHAPE, false>(m, ap, ae, al, in, out); return 0;
    case GB_GRAD: run_scatter<DIM, DEG, GB_GRAD, false>(m, ap, ae, al, in, out); return 0;
    case GB_STRAIN: if (weighted) run_scatter<DIM, DEG, GB_STRAIN, true>(m, ap, ae, al, in, out); else run_scatter<DIM, DEG, GB_STRAIN, false>(m, ap, ae, al, in, out); return 0;
  }
  return 1;
}

#define EMUL_DISPATCH(dim, degree, CALL)          \
  do {                                            \
    if (dim == 2 && degree == 1) { CALL(2, 1); }  \
    else if (dim == 2) { CALL(2, 2); }            \
    else if (degree == 1) { CALL(3, 1); }         \
    else { CALL(3, 2); }                          \
  } while (0)

}  // namespace

// one thread per node of k_grid_gp_scatter (gauss_ops.cu)
template <int B, bool W>
static void run_grid_scatter(const GridTri& gt, int heron, const QuadRule& rule, const double* in, double* out, const double* xy = nullptr) {
  constexpr int NC = GpShape<2, 1, B>::NC;
  const long long nn = (long long)(gt.m + 1) * (gt.n + 1);
  for (long long r = 0; r < nn; r++) {
    double acc[NC];
    grid_scatter_node<B, W>(gt, heron, rule, rule.n, (int)(r / (gt.m + 1)), (int)(r % (gt.m + 1)), in, acc, xy);
    for (int c = 0; c < NC; c++) out[r + c * nn] = acc[c];
  }
}

// k_row_gather_fwd (row_gather.cuh): every CTA as a loop over its threads (phase 1) and the flat copy (phase 2); -1 when a CTA's rows exceed RG_CAP
template <int DIM, int DEG, int OP>
static int run_row_gather(const DevMesh& m, const long long* ap, const int* ae, const uint8_t* al, const long long* rowptr, const int* colind,
                          const double* coef, double* vals, int rows) {
  static double acc[RG_CAP];
  for (int r0 = 0; r0 < m.ndof; r0 += rows) {
    const int r1 = r0 + rows < m.ndof ? r0 + rows : m.ndof;
    const long long rs0 = rowptr[r0];
    const int total = (int)(rowptr[r1] - rs0);
    if (total > RG_CAP) return -1;
    for (int c = 0; c < RG_CAP; c++) acc[c] = -7.0e300;
    for (int t = 0; t < r1 - r0; t++) { const int r = g_reverse ? r1 - 1 - t : r0 + t; rg_row<DIM, DEG, OP>(m, ap, ae, al, rowptr, colind, r, rs0, coef, acc); }
    for (int idx = 0; idx < total; idx++) vals[rs0 + idx] = acc[idx];
  }
  return 0;
}

// k_row_gather_elast_fwd (row_gather.cuh)
template <int DIM>
static int run_row_gather_elast(const DevMesh& m, const long long* ap, const int* ae, const uint8_t* al, const long long* rowptr, const int* colind,
                                long long nnz, const double* hbar, double* vals) {
  constexpr int NC = DIM;
  static double acc[9 * 8192];
  for (int r0 = 0; r0 < m.ndof; r0 += RGE_THREADS) {
    const int r1 = r0 + RGE_THREADS < m.ndof ? r0 + RGE_THREADS : m.ndof;
    const long long rs0 = rowptr[r0];
    const int T = (int)(rowptr[r1] - rs0);
    if (T > 8192) return -1;
    for (int c = 0; c < NC * NC * T; c++) acc[c] = -7.0e300;
    for (int t = 0; t < r1 - r0; t++) { const int r = g_reverse ? r1 - 1 - t : r0 + t; rg_row_elast<DIM>(m, ap, ae, al, rowptr, colind, r, rs0, T, hbar, acc); }
    for (int a = 0; a < NC; a++)
      for (int idx = 0; idx < NC * T; idx++) vals[NC * ((long long)a * nnz + rs0) + idx] = acc[a * NC * T + idx];
  }
  return 0;
}

// one thread per node of k_tet_gp_scatter (gauss_ops.cu)
template <int B, bool W>
static void run_tet_scatter(const GridTet& gt, const QuadRule& rule, const double* in, double* out) {
  constexpr int NC = GpShape<3, 1, B>::NC;
  const long long n1 = gt.n + 1, nn = n1 * n1 * (gt.l + 1);
  for (long long r = 0; r < nn; r++) {
    double acc[NC];
    tg_scatter_node<B, W>(gt, rule, rule.n, (int)(r % n1), (int)((r / n1) % n1), (int)(r / (n1 * n1)), in, acc);
    for (int c = 0; c < NC; c++) out[r + c * nn] = acc[c];
  }
}

extern "C" {

void emul_set_reverse(int on) { g_reverse = on; }

// kind / adjoint as adfem_gauss_op / adfem_gauss_op_adjoint (include/adfem_cuda.h); the kind -> (basis, weighted, direction) table is the one
// of gp_kind() in adfem_cuda.cu
int emul_gauss_op(int dim, int degree, int order, int nv, int ne, int ndof, const double* coords, const int* verts, const int* conn,
                  const long long* adj_ptr, const int* adj_elem, const uint8_t* adj_loc, int kind, int adjoint, const double* in, double* out) {
  bool ok;
  const DevMesh m = make_mesh(dim, degree, order, nv, ne, ndof, coords, verts, conn, ok);
  if (!ok || kind < 0 || kind > 4) return 1;
  const int basis = kind == 4 ? GB_STRAIN : kind;
  const bool weighted = kind == 4, scatter_fwd = kind == 4;
  const bool to_gauss = adjoint ? scatter_fwd : !scatter_fwd;
  int rc = 1;
#define CALL_K(DIM, DEG) rc = run_kind<DIM, DEG>(m, adj_ptr, adj_elem, adj_loc, basis, weighted, to_gauss, in, out)
  EMUL_DISPATCH(dim, degree, CALL_K);
#undef CALL_K
  return rc;
}

int emul_laplace_term(int dim, int degree, int order, int nv, int ne, int ndof, const double* coords, const int* verts, const int* conn,
                      const long long* adj_ptr, const int* adj_elem, const uint8_t* adj_loc, const double* nu, const double* u, double* out) {
  bool ok;
  const DevMesh m = make_mesh(dim, degree, order, nv, ne, ndof, coords, verts, conn, ok);
  if (!ok) return 1;
#define CALL_L(DIM, DEG) for (int r = 0; r < m.ndof; r++) out[r] = laplace_term_row<DIM, DEG>(m, adj_ptr, adj_elem, adj_loc, r, nu, u)
  EMUL_DISPATCH(dim, degree, CALL_L);
#undef CALL_L
  return 0;
}

int emul_laplace_term_grad_nu(int dim, int degree, int order, int nv, int ne, int ndof, const double* coords, const int* verts, const int* conn,
                              const double* u, const double* go, double* grad_nu) {
  bool ok;
  const DevMesh m = make_mesh(dim, degree, order, nv, ne, ndof, coords, verts, conn, ok);
  if (!ok) return 1;
#define CALL_G(DIM, DEG) for (int e = 0; e < m.ne; e++) laplace_term_grad_nu_body<DIM, DEG>(m, e, u, go, grad_nu)
  EMUL_DISPATCH(dim, degree, CALL_G);
#undef CALL_G
  return 0;
}

// option "coef_presum": k_presum_coef / k_expand_grad of gauss_ops.cu
int emul_presum_coef(int dim, int order, long long ne, int ns2, const double* coef, double* hbar) {
  QuadRule r;
  if (!(dim == 2 ? triangle_rule(order, r) : tetrahedron_rule(order, r))) return 1;
  for (long long i = 0; i < ne * ns2; i++) hbar[i] = presum_coef_body(r, r.n, ns2, i, coef);
  return 0;
}
int emul_expand_grad(int dim, int order, long long ne, int ns2, const double* gbar, double* grad) {
  QuadRule r;
  if (!(dim == 2 ? triangle_rule(order, r) : tetrahedron_rule(order, r))) return 1;
  for (long long i = 0; i < ne * r.n * ns2; i++) grad[i] = expand_grad_body(r, r.n, ns2, i, gbar);
  return 0;
}

// structured Q1 scalar siblings: k_quad_scalar_fwd / _bwd, k_quad_source_fwd / _bwd of grid_ops.cu
void emul_quad_scalar(int op, const double* coef, int m, int n, double h, long long* ii, long long* jj, double* vv) {
  for (long long t = 0; t < 4LL * m * n; t++) quad_scalar_fwd_body(op, t, coef, m, h, ii, jj, vv);
}
void emul_quad_scalar_grad(int op, const double* grad_vv, int m, int n, double h, double* grad_coef) {
  for (long long t = 0; t < 4LL * m * n; t++) grad_coef[t] = quad_scalar_bwd_body(op, t, grad_vv, h);
}
void emul_quad_source(const double* f, int m, int n, double h, double* rhs) {
  for (long long node = 0; node < (long long)(m + 1) * (n + 1); node++) rhs[node] = quad_source_node(node, f, m, n, h);
}
void emul_quad_source_grad(const double* grad_rhs, int m, int n, double h, double* grad_f) {
  for (long long t = 0; t < 4LL * m * n; t++) grad_f[t] = quad_source_bwd_body(t, grad_rhs, m, h);
}

// fused constitutive pre-step: k_presum_plane / k_expand_plane_grad of gauss_ops.cu (P1 triangles)
int emul_presum_plane(int order, long long ne, int mode, const double* E, const double* nu, double* hbar) {
  QuadRule r;
  if (!triangle_rule(order, r)) return 1;
  for (long long i = 0; i < ne * 9; i++) hbar[i] = presum_plane_body(r, r.n, mode, i, E, nu);
  return 0;
}
int emul_expand_plane_grad(int order, long long ne, int mode, const double* E, const double* nu, const double* gbar, double* gE, double* gnu) {
  QuadRule r;
  if (!triangle_rule(order, r)) return 1;
  for (long long t = 0; t < ne * r.n; t++) expand_plane_grad_body(r, r.n, mode, t, E, nu, gbar, gE, gnu);
  return 0;
}

// structured P1 elasticity (grid_elast.cuh): every warp of k_grid_elast_fwd / k_grid_elast_adj as three loops over its 32 lanes per row
// plane_mode < 0: coef = H; plane_mode = 0 | 1: coef = E, coef2 = nu (fused constitutive step)
// xy != nullptr: MAPPED instantiation (node positions from the coordinate array [node][2]; xs / ys are not read)
static int emul_grid_elast_fwd_impl(int m, int n, const double* xs, const double* ys, int order, int heron, int rows_per_warp, long long nnz, const double* coef,
                                    double* vals, int plane_mode, const double* coef2, const double* xy) {
  QuadRule rule;
  if (!triangle_rule(order, rule) || rule.n != GE_G) return 1;
  const GridTri gt{m, n, xs, ys};
  const int strips = (m + 1 + GE_COLS - 1) / GE_COLS, chunks = (n + 1 + rows_per_warp - 1) / rows_per_warp;
  static double smem[GE_FWD_WARP_DOUBLES];
  for (long long gw = 0; gw < (long long)strips * chunks; gw++) {
    const int strip = (int)(gw % strips), chunk = (int)(gw / strips), j0 = strip * GE_COLS;
    const int i0 = chunk * rows_per_warp, i1 = ge_min(i0 + rows_per_warp, n + 1);
    for (int k = 0; k < GE_FWD_WARP_DOUBLES; k++) smem[k] = -7.0e300;           // poison: nothing may be read before it is written
    double *P = smem, *C = P + GE_HROW, *stage = C + GE_HROW;
    auto load = [&](int lane, int ci, double* buf) {
      if (plane_mode >= 0) ge_load_cell_row_plane<GE_G>(lane, rule, m, n, ci, j0, plane_mode, coef, coef2, buf);
      else ge_load_cell_row<GE_G>(lane, rule, m, n, ci, j0, coef, buf);
    };
    EMUL_LANES(lane) load(lane, i0 - 1, P);
    long long rowbase = grid_rowptr(i0, 0, m, n);
    for (int i = i0; i < i1; i++) {
      EMUL_LANES(lane) load(lane, i, C);
      if (xy) { EMUL_LANES(lane) ge_node<true>(lane, heron, gt, i, j0, P, C, stage, xy); }
      else { EMUL_LANES(lane) ge_node(lane, heron, gt, i, j0, P, C, stage); }
      EMUL_LANES(lane) ge_store_node_row(lane, m, n, i, j0, rowbase, nnz, stage, vals);
      rowbase += ge_prefix(m + 1, m, i > 0, i < n);
      double* t = P; P = C; C = t;
    }
  }
  return 0;
}
int emul_grid_elast_fwd(int m, int n, const double* xs, const double* ys, int order, int heron, int rows_per_warp, long long nnz, const double* coef,
                        double* vals, int plane_mode, const double* coef2) {
  return emul_grid_elast_fwd_impl(m, n, xs, ys, order, heron, rows_per_warp, nnz, coef, vals, plane_mode, coef2, nullptr);
}
int emul_grid_elast_fwd_mapped(int m, int n, const double* xy, int order, int heron, int rows_per_warp, long long nnz, const double* coef, double* vals) {
  return emul_grid_elast_fwd_impl(m, n, nullptr, nullptr, order, heron, rows_per_warp, nnz, coef, vals, -1, nullptr, xy);
}
static int emul_grid_elast_adj_impl(int m, int n, const double* xs, const double* ys, int order, int heron, int rows_per_warp, long long nnz, const double* dvals,
                                    double* grad, int plane_mode, const double* E, const double* nu, double* grad2, const double* xy) {
  QuadRule rule;
  if (!triangle_rule(order, rule) || rule.n != GE_G) return 1;
  const GridTri gt{m, n, xs, ys};
  const int strips = (m + GE_COLS - 1) / GE_COLS, chunks = (n + rows_per_warp - 1) / rows_per_warp;
  static double smem[GE_ADJ_WARP_DOUBLES];
  for (long long gw = 0; gw < (long long)strips * chunks; gw++) {
    const int strip = (int)(gw % strips), chunk = (int)(gw / strips), c0 = strip * GE_COLS;
    const int r0 = chunk * rows_per_warp, r1 = ge_min(r0 + rows_per_warp, n);
    for (int k = 0; k < GE_ADJ_WARP_DOUBLES; k++) smem[k] = -7.0e300;
    double *lo = smem, *hi = lo + 2 * GE_NROW, *gst = hi + 2 * GE_NROW;
    long long rowbase = grid_rowptr(r0, 0, m, n);
    EMUL_LANES(lane) ge_load_node_row(lane, m, n, r0, c0, rowbase, nnz, dvals, lo);
    for (int ci = r0; ci < r1; ci++) {
      rowbase += ge_prefix(m + 1, m, ci > 0, ci < n);
      EMUL_LANES(lane) ge_load_node_row(lane, m, n, ci + 1, c0, rowbase, nnz, dvals, hi);
      if (xy) { EMUL_LANES(lane) ge_cell_adjoint<true>(lane, heron, gt, ci, c0, lo, hi, gst, xy); }
      else { EMUL_LANES(lane) ge_cell_adjoint(lane, heron, gt, ci, c0, lo, hi, gst); }
      EMUL_LANES(lane) {
        if (plane_mode >= 0) ge_store_cell_row_plane<GE_G>(lane, rule, m, ci, c0, plane_mode, E, nu, gst, grad, grad2);
        else ge_store_cell_row<GE_G>(lane, rule, m, ci, c0, gst, grad);
      }
      double* t = lo; lo = hi; hi = t;
    }
  }
  return 0;
}

int emul_grid_elast_adj(int m, int n, const double* xs, const double* ys, int order, int heron, int rows_per_warp, long long nnz, const double* dvals,
                        double* grad, int plane_mode, const double* E, const double* nu, double* grad2) {
  return emul_grid_elast_adj_impl(m, n, xs, ys, order, heron, rows_per_warp, nnz, dvals, grad, plane_mode, E, nu, grad2, nullptr);
}
int emul_grid_elast_adj_mapped(int m, int n, const double* xy, int order, int heron, int rows_per_warp, long long nnz, const double* dvals, double* grad) {
  return emul_grid_elast_adj_impl(m, n, nullptr, nullptr, order, heron, rows_per_warp, nnz, dvals, grad, -1, nullptr, nullptr, nullptr, xy);
}

// structured scatter-type Gauss-point operators (grid_gauss.cuh): k_grid_gp_scatter / k_grid_laplace_term of gauss_ops.cu
// xy != nullptr: mapped grid (positions from the coordinate array, xs / ys unused)
static int emul_grid_gp_scatter_impl(int m, int n, const double* xs, const double* ys, int order, int heron, int basis, int weighted, const double* in, double* out,
                                     const double* xy) {
  QuadRule rule;
  if (!triangle_rule(order, rule)) return 1;
  const GridTri gt{m, n, xs, ys};
  switch (basis) {
    case GB_P1SHAPE: run_grid_scatter<GB_P1SHAPE, false>(gt, heron, rule, in, out, xy); return 0;
    case GB_SHAPE: run_grid_scatter<GB_SHAPE, false>(gt, heron, rule, in, out, xy); return 0;
    case GB_GRAD: run_grid_scatter<GB_GRAD, false>(gt, heron, rule, in, out, xy); return 0;
    case GB_STRAIN: if (weighted) run_grid_scatter<GB_STRAIN, true>(gt, heron, rule, in, out, xy); else run_grid_scatter<GB_STRAIN, false>(gt, heron, rule, in, out, xy); return 0;
  }
  return 1;
}
int emul_grid_gp_scatter(int m, int n, const double* xs, const double* ys, int order, int heron, int basis, int weighted, const double* in, double* out) {
  return emul_grid_gp_scatter_impl(m, n, xs, ys, order, heron, basis, weighted, in, out, nullptr);
}
int emul_grid_gp_scatter_mapped(int m, int n, const double* xy, int order, int heron, int basis, int weighted, const double* in, double* out) {
  return emul_grid_gp_scatter_impl(m, n, nullptr, nullptr, order, heron, basis, weighted, in, out, xy);
}
int emul_grid_laplace_term(int m, int n, const double* xs, const double* ys, int order, int heron, const double* nu, const double* u, double* out) {
  QuadRule rule;
  if (!triangle_rule(order, rule)) return 1;
  const GridTri gt{m, n, xs, ys};
  for (long long r = 0; r < (long long)(m + 1) * (n + 1); r++) out[r] = grid_laplace_term_node(gt, heron, rule, rule.n, (int)(r / (m + 1)), (int)(r % (m + 1)), nu, u);
  return 0;
}
int emul_grid_laplace_term_mapped(int m, int n, const double* xy, int order, int heron, const double* nu, const double* u, double* out) {
  QuadRule rule;
  if (!triangle_rule(order, rule)) return 1;
  const GridTri gt{m, n, nullptr, nullptr};
  for (long long r = 0; r < (long long)(m + 1) * (n + 1); r++) out[r] = grid_laplace_term_node(gt, heron, rule, rule.n, (int)(r / (m + 1)), (int)(r % (m + 1)), nu, u, xy);
  return 0;
}

// structured tetrahedral grid (tet_grid.cuh): every warp (= node) of k_tet_grid_elast_fwd as three loops over its lanes
int emul_tet_grid_elast_fwd(int n, int l, const double* xs, const double* ys, const double* zs, long long nnz, const long long* rowptr, const double* hbar,
                            double* vals) {
  static TetGridTables tab;
  build_tet_grid_tables(tab);
  const GridTet gt{n, l, xs, ys, zs, &tab};
  static double smem[TG_WARP_DOUBLES];
  const long long n1 = n + 1, nn = n1 * n1 * (l + 1);
  for (long long node = 0; node < nn; node++) {
    const int i = (int)(node % n1), j = (int)((node / n1) % n1), k = (int)(node / (n1 * n1)), par = (i + j + k) & 1;
    for (int c = 0; c < TG_WARP_DOUBLES; c++) smem[c] = -7.0e300;
    double *tb = smem, *stage = tb + 32 * TG_BLK;
    EMUL_LANES(lane) tg_tet_block(lane, gt, par, i, j, k, hbar, tb);
    const int mask = tg_row_mask(gt, par, i, j, k);
    if (tg_popc(mask) != (int)(rowptr[node + 1] - rowptr[node])) return 2;          // the closed-form row must be the symbolic one
    EMUL_LANES(lane) tg_gather_rows(lane, gt, par, mask, tb, stage);
    EMUL_LANES(lane) tg_store_rows(lane, rowptr[node], tg_popc(mask), nnz, stage, vals);
  }
  return 0;
}

int emul_tet_grid_elast_adj(int n, int l, const double* xs, const double* ys, const double* zs, int order, long long nnz, const long long* rowptr,
                            const double* dvals, double* grad) {
  static TetGridTables tab;
  build_tet_grid_tables(tab);
  QuadRule rule;
  if (!tetrahedron_rule(order, rule)) return 1;
  const GridTet gt{n, l, xs, ys, zs, &tab};
  alignas(16) static double smem[TG_ADJ_WARP_DOUBLES];
  const long long ne = 5LL * n * n * l;
  for (long long e0 = 0; e0 < ne; e0 += 32) {
    for (int c = 0; c < TG_ADJ_WARP_DOUBLES; c++) smem[c] = -7.0e300;
    EMUL_LANES(lane) tg_tet_adjoint(lane, gt, e0, ne, nnz, rowptr, dvals, smem);
    EMUL_LANES(lane) tg_store_grad(lane, rule, rule.n, e0, ne, smem, grad);
  }
  return 0;
}

int emul_row_gather_fwd(int dim, int degree, int order, int nv, int ne, int ndof, const double* coords, const int* verts, const int* conn,
                        const long long* adj_ptr, const int* adj_elem, const uint8_t* adj_loc, const long long* rowptr, const int* colind, int op,
                        const double* coef, double* vals, int rows) {
  bool ok;
  const DevMesh m = make_mesh(dim, degree, order, nv, ne, ndof, coords, verts, conn, ok);
  if (!ok || (op != OP_LAPLACE && op != OP_MASS)) return 1;
  int rc = 1;
#define CALL_RG(DIM, DEG) rc = op == OP_LAPLACE ? run_row_gather<DIM, DEG, OP_LAPLACE>(m, adj_ptr, adj_elem, adj_loc, rowptr, colind, coef, vals, rows) \
                                                  : run_row_gather<DIM, DEG, OP_MASS>(m, adj_ptr, adj_elem, adj_loc, rowptr, colind, coef, vals, rows)
  EMUL_DISPATCH(dim, degree, CALL_RG);
#undef CALL_RG
  return rc;
}

int emul_row_gather_elast_fwd(int dim, int order, int nv, int ne, const double* coords, const int* verts, const int* conn, const long long* adj_ptr,
                              const int* adj_elem, const uint8_t* adj_loc, const long long* rowptr, const int* colind, const double* hbar, double* vals) {
  bool ok;
  const DevMesh m = make_mesh(dim, 1, order, nv, ne, nv, coords, verts, conn, ok);
  if (!ok) return 1;
  const long long nnz = rowptr[nv];
  return dim == 2 ? run_row_gather_elast<2>(m, adj_ptr, adj_elem, adj_loc, rowptr, colind, nnz, hbar, vals)
                  : run_row_gather_elast<3>(m, adj_ptr, adj_elem, adj_loc, rowptr, colind, nnz, hbar, vals);
}

int emul_tet_gp_scatter(int n, int l, const double* xs, const double* ys, const double* zs, int order, int basis, int weighted, const double* in,
                        double* out) {
  static TetGridTables tab;
  build_tet_grid_tables(tab);
  QuadRule rule;
  if (!tetrahedron_rule(order, rule)) return 1;
  const GridTet gt{n, l, xs, ys, zs, &tab};
  switch (basis) {
    case GB_P1SHAPE: run_tet_scatter<GB_P1SHAPE, false>(gt, rule, in, out); return 0;
    case GB_SHAPE: run_tet_scatter<GB_SHAPE, false>(gt, rule, in, out); return 0;
    case GB_GRAD: run_tet_scatter<GB_GRAD, false>(gt, rule, in, out); return 0;
    case GB_STRAIN: if (weighted) run_tet_scatter<GB_STRAIN, true>(gt, rule, in, out); else run_tet_scatter<GB_STRAIN, false>(gt, rule, in, out); return 0;
  }
  return 1;
}
int emul_tet_laplace_term(int n, int l, const double* xs, const double* ys, const double* zs, int order, const double* nu, const double* u, double* out) {
  static TetGridTables tab;
  build_tet_grid_tables(tab);
  QuadRule rule;
  if (!tetrahedron_rule(order, rule)) return 1;
  const GridTet gt{n, l, xs, ys, zs, &tab};
  const long long n1 = n + 1;
  for (long long r = 0; r < n1 * n1 * (l + 1); r++) out[r] = tg_laplace_term_node(gt, rule, rule.n, (int)(r % n1), (int)((r / n1) % n1), (int)(r / (n1 * n1)), nu, u);
  return 0;
}

// structured tetrahedral grid, scalar operators (tet_scalar.cuh): CTAs of k_tet_grid_scalar_fwd, threads of k_tet_grid_scalar_adj
int emul_tet_grid_scalar(int n, int l, const double* xs, const double* ys, const double* zs, int order, int op, int adjoint, const long long* rowptr,
                         const double* in, double* out) {
  static TetGridTables tab;
  build_tet_grid_tables(tab);
  QuadRule rule;
  if (!tetrahedron_rule(order, rule) || (op != OP_LAPLACE && op != OP_MASS)) return 1;
  const GridTet gt{n, l, xs, ys, zs, &tab};
  const long long n1 = n + 1, nn = n1 * n1 * (l + 1), ne = 5LL * n * n * l;
  if (adjoint) {
    for (long long e = 0; e < ne; e++) {
      if (op == OP_LAPLACE) tgs_tet_adjoint<OP_LAPLACE>(gt, rule, rule.n, e, rowptr, in, out); else tgs_tet_adjoint<OP_MASS>(gt, rule, rule.n, e, rowptr, in, out);
    }
    return 0;
  }
  static double acc[RG_CAP];
  for (long long r0 = 0; r0 < nn; r0 += RG_THREADS) {
    const long long r1 = r0 + RG_THREADS < nn ? r0 + RG_THREADS : nn, rs0 = rowptr[r0];
    if (rowptr[r1] - rs0 > RG_CAP) return -1;
    for (int c = 0; c < RG_CAP; c++) acc[c] = -7.0e300;
    for (long long t = 0; t < r1 - r0; t++) {
      const long long r = g_reverse ? r1 - 1 - t : r0 + t;
      const int i = (int)(r % n1), j = (int)((r / n1) % n1), k = (int)(r / (n1 * n1));
      if (op == OP_LAPLACE) tgs_row<OP_LAPLACE>(gt, rule, rule.n, i, j, k, rowptr[r], rs0, in, acc); else tgs_row<OP_MASS>(gt, rule, rule.n, i, j, k, rowptr[r], rs0, in, acc);
    }
    for (long long idx = 0; idx < rowptr[r1] - rs0; idx++) out[rs0 + idx] = acc[idx];
  }
  return 0;
}

// k_quad_stiff1_svt_fwd / _bwd of grid_ops.cu
void emul_quad_stiffness1_svt(const double* mu, int type, int m, int n, double h, long long* ii, long long* jj, double* vv) {
  for (long long t = 0; t < 4LL * m * n; t++) quad_stiff1_svt_fwd_body(t, mu, type, m, n, h, ii, jj, vv);
}
void emul_quad_stiffness1_svt_grad(const double* grad_vv, int type, int m, int n, double h, double* grad_mu) {
  for (long long t = 0; t < 4LL * m * n; t++) quad_stiff1_svt_bwd_body(t, grad_vv, type, m, n, h, grad_mu);
}

void emul_plane_matrix(int mode, long long n, const double* E, const double* nu, double* H) {
  for (long long i = 0; i < n; i++) plane_matrix_body(mode, E[i], nu[i], H + 9 * i);
}
void emul_plane_matrix_grad(int mode, long long n, const double* E, const double* nu, const double* gH, double* gE, double* gnu) {
  for (long long i = 0; i < n; i++) plane_matrix_grad_body(mode, E[i], nu[i], gH + 9 * i, gE + i, gnu + i);
}
}
