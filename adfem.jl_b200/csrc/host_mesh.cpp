#include "host_mesh.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>

#include <sys/mman.h>

namespace adfem {

void advise_huge_pages(void* p, size_t bytes) {
#ifdef MADV_HUGEPAGE
  const uintptr_t huge = (uintptr_t)2 << 20;
  const uintptr_t a = ((uintptr_t)p + huge - 1) & ~(huge - 1), e = ((uintptr_t)p + bytes) & ~(huge - 1);
  if (p && e > a) (void)madvise((void*)a, e - a, MADV_HUGEPAGE);
#else
  (void)p; (void)bytes;
#endif
}

void host_parallel_for(long long n, int nthreads, const std::function<void(long long, long long)>& fn, long long serial_below) {
  int nt = nthreads;
  if (nt <= 0) { if (const char* s = getenv("ADFEM_HOST_THREADS")) nt = atoi(s); }
  if (nt <= 0) { unsigned h = std::thread::hardware_concurrency(); nt = h == 0 ? 4 : (int)std::min(h, 64u); }
  if (nt <= 1 || n < serial_below) { fn(0LL, n); return; }
  std::vector<std::thread> th;
  const long long chunk = (n + nt - 1) / nt;
  for (int t = 0; t < nt; t++) {
    const long long a = t * chunk, b = std::min(n, a + chunk);
    if (a < b) th.emplace_back([=, &fn] { fn(a, b); });
  }
  for (auto& x : th) x.join();
}

void radix_sort_by_key(std::vector<KeyId>& a, int nthreads) {
  const long long n = (long long)a.size();
  constexpr int BITS = 11, NB = 1 << BITS;
  const int T = std::max(1, nthreads);
  uint64_t any = 0, all = ~0ULL;
  for (const KeyId& p : a) { any |= p.key; all &= p.key; }
  const uint64_t varying = any ^ all;
  std::vector<KeyId> b;
  b.reserve((size_t)n);
  advise_huge_pages(b.data(), (size_t)n * sizeof(KeyId));
  {
    KeyId* p = b.data();
    host_parallel_for(n, T, [&](long long i0, long long i1) { memset(static_cast<void*>(p + i0), 0, (size_t)(i1 - i0) * sizeof(KeyId)); }, 1 << 16);
  }
  b.resize((size_t)n);
  std::vector<long long> cut(T + 1);
  for (int t = 0; t <= T; t++) cut[t] = n * t / T;
  std::vector<std::vector<long long>> hist(T, std::vector<long long>(NB));
  for (int shift = 0; shift < 64; shift += BITS) {
    if (((varying >> shift) & (NB - 1)) == 0) continue;
    {
      std::vector<std::thread> th;
      for (int t = 0; t < T; t++) th.emplace_back([&, t] {
        std::vector<long long>& h = hist[t];
        std::fill(h.begin(), h.end(), 0);
        for (long long i = cut[t]; i < cut[t + 1]; i++) h[(a[i].key >> shift) & (NB - 1)]++;
      });
      for (auto& x : th) x.join();
    }
    long long run = 0;
    for (int dgt = 0; dgt < NB; dgt++)
      for (int t = 0; t < T; t++) { const long long c = hist[t][dgt]; hist[t][dgt] = run; run += c; }
    {
      std::vector<std::thread> th;
      for (int t = 0; t < T; t++) th.emplace_back([&, t] {
        std::vector<long long>& h = hist[t];
        for (long long i = cut[t]; i < cut[t + 1]; i++) b[h[(a[i].key >> shift) & (NB - 1)]++] = a[i];
      });
      for (auto& x : th) x.join();
    }
    a.swap(b);
  }
}

std::vector<int> soa_copy(const std::vector<int>& aos, long long ne, int kcount) {
  std::vector<int> out;
  const size_t total = (size_t)ne * kcount;
  out.reserve(total);
  advise_huge_pages(out.data(), total * sizeof(int));
  int* o = out.data();
  host_parallel_for((long long)total, 0, [&](long long a, long long b) { memset(static_cast<void*>(o + a), 0, (size_t)(b - a) * sizeof(int)); }, 1 << 20);
  out.resize(total);
  o = out.data();
  const int* in = aos.data();
  host_parallel_for(ne, 0, [&](long long e0, long long e1) {
    for (long long e = e0; e < e1; e++)
      for (int k = 0; k < kcount; k++) o[(size_t)k * ne + e] = in[(size_t)e * kcount + k];
  }, 1 << 16);
  return out;
}

namespace {
// static block partition of [0, n) over host threads (ADFEM_HOST_THREADS, default: the hardware's); element loops below are independent per element
template <class F> void par_elems(long long n, F fn) {
  int nt = 0;
  if (const char* s = getenv("ADFEM_HOST_THREADS")) nt = atoi(s);
  if (nt <= 0) { unsigned h = std::thread::hardware_concurrency(); nt = h == 0 ? 4 : (int)std::min(h, 64u); }
  if (nt <= 1 || n < 65536) { fn(0LL, n); return; }
  std::vector<std::thread> th;
  const long long chunk = (n + nt - 1) / nt;
  for (int t = 0; t < nt; t++) {
    const long long a = t * chunk, b = std::min(n, a + chunk);
    if (a < b) th.emplace_back([=] { fn(a, b); });
  }
  for (auto& x : th) x.join();
}

// zero reserved (not yet constructed) storage of plain data: maps its pages from the calling thread; the resize that follows finds them mapped
template <class T> void prefault(T* p, size_t n) { if (n) memset(static_cast<void*>(p), 0, n * sizeof(T)); }

// local edges in MFEM geometry order (Geometry::Constants<TRIANGLE/TETRAHEDRON>::Edges)
const int kTriEdges[3][2] = {{0, 1}, {1, 2}, {2, 0}};
const int kTetEdges[6][2] = {{0, 1}, {0, 2}, {0, 3}, {1, 2}, {1, 3}, {2, 3}};

// Global edge ids in first-appearance order while walking elements 0..E-1 and their local edges in
// geometry order — the numbering MFEM's GetElementToEdgeTable produces and the reference exposes
// through `edges` and `dof[k+nv]` (deps/MFEM/Common.cpp:70-74,133-140).
int host_threads() {
  int nt = 0;
  if (const char* s = getenv("ADFEM_HOST_THREADS")) nt = atoi(s);
  if (nt <= 0) { unsigned h = std::thread::hardware_concurrency(); nt = h == 0 ? 4 : (int)std::min(h, 64u); }
  return nt;
}

// The numbering of EdgeNumbering below (edge ids in order of first appearance over (element, local edge)) without the sequential walk: sort the
// (edge key, appearance) records by key — the first record of a key is its first appearance —, then rank the edges by that first appearance.
// ids[e * nel + j] = edge id of local edge j of element e; lo / hi = end points per edge id.  Same arrays as the walk (tests/test_host_plan.py).
void number_edges_sorted(const std::vector<int>& verts, int ne, int nvl, int nel, const int (*local)[2], int nt, std::vector<int>& ids,
                         std::vector<int>& lo, std::vector<int>& hi) {
  const long long nrec = (long long)ne * nel;
  std::vector<KeyId> rec;
  rec.reserve((size_t)nrec);
  advise_huge_pages(rec.data(), (size_t)nrec * sizeof(KeyId));
  {
    KeyId* p = rec.data();
    host_parallel_for(nrec, nt, [&](long long a, long long b) { memset(static_cast<void*>(p + a), 0, (size_t)(b - a) * sizeof(KeyId)); }, 1 << 16);
  }
  rec.resize((size_t)nrec);
  host_parallel_for(ne, nt, [&](long long e0, long long e1) {
    for (long long e = e0; e < e1; e++) {
      const int* vi = &verts[(size_t)e * nvl];
      for (int j = 0; j < nel; j++) {
        const int a = vi[local[j][0]], b = vi[local[j][1]];
        const uint64_t l = (uint64_t)(a <= b ? a : b), h = (uint64_t)(a <= b ? b : a);
        rec[(size_t)e * nel + j] = KeyId{l << 32 | h, (int)(e * nel + j)};
      }
    }
  }, 1 << 14);
  radix_sort_by_key(rec, nt);
  // groups of equal keys; group g starts at record start[g], which carries the smallest appearance index of the edge
  std::vector<KeyId> first;             // (first appearance, group)
  std::vector<long long> start;
  for (long long i = 0; i < nrec; i++)
    if (i == 0 || rec[i].key != rec[i - 1].key) { first.push_back(KeyId{(uint64_t)rec[i].id, (int)start.size()}); start.push_back(i); }
  const long long nedge = (long long)start.size();
  start.push_back(nrec);
  radix_sort_by_key(first, nt);         // by first appearance: position = edge id
  std::vector<int> id_of_group((size_t)nedge);
  lo.resize((size_t)nedge); hi.resize((size_t)nedge);
  host_parallel_for(nedge, nt, [&](long long a, long long b) {
    for (long long id = a; id < b; id++) {
      const int g = first[id].id;
      id_of_group[g] = (int)id;
      const uint64_t key = rec[start[g]].key;
      lo[id] = (int)(key >> 32); hi[id] = (int)(key & 0xffffffffu);
    }
  }, 1 << 14);
  ids.resize((size_t)nrec);
  host_parallel_for(nedge, nt, [&](long long a, long long b) {
    for (long long g = a; g < b; g++)
      for (long long i = start[g]; i < start[g + 1]; i++) ids[rec[i].id] = id_of_group[g];
  }, 1 << 14);
}

struct EdgeNumbering {
  std::vector<int> head, next, hi, lo;
  // `expect`: a guess of the edge count (the lists are reserved and advised for huge pages: the walk below is one dependent random access after
  // the other on renumbered meshes; a wrong guess only means the vectors grow the usual way)
  EdgeNumbering(int nv, size_t expect) {
    head.reserve(nv); advise_huge_pages(head.data(), (size_t)nv * sizeof(int)); head.assign(nv, -1);
    for (std::vector<int>* v : {&next, &hi, &lo}) { v->reserve(expect); advise_huge_pages(v->data(), expect * sizeof(int)); }
  }
  int id(int a, int b) {
    int r = a <= b ? a : b, c = a <= b ? b : a;
    for (int n = head[r]; n >= 0; n = next[n])
      if (hi[n] == c) return n;
    int i = (int)hi.size();
    hi.push_back(c); lo.push_back(r); next.push_back(head[r]); head[r] = i;
    return i;
  }
};
}  // namespace

std::string HostMesh::build(int dim_, const double* vertices, int vstride, int nv_, const int* elems, int ne_, int order_,
                            int degree_, int lorder_) {
  dim = dim_; nv = nv_; ne = ne_; order = order_; degree = degree_; lorder = lorder_;
  if (dim != 2 && dim != 3) return "dim must be 2 or 3";
  if (degree != 1 && degree != 2) return "degree must be equal to 1 or 2";   // deps/MFEM3/Common.cpp:52 (BDM1 is out of scope)
  if (!(dim == 2 ? triangle_rule(order, rule) : tetrahedron_rule(order, rule))) return "unsupported quadrature order";
  g = rule.n;
  const int nvl = dim + 1, nel = dim == 2 ? 3 : 6;
  d = degree == 1 ? nvl : nvl + nel;
  // copies and the range check in element / vertex blocks over the host threads (the first touch of the new arrays is most of their cost)
  coords.clear(); coords.reserve((size_t)nv * dim);
  verts.clear(); verts.reserve((size_t)ne * nvl);
  advise_huge_pages(coords.data(), coords.capacity() * sizeof(double));
  advise_huge_pages(verts.data(), verts.capacity() * sizeof(int));
  par_elems(nv, [&](long long i0, long long i1) { prefault(coords.data() + (size_t)i0 * dim, (size_t)(i1 - i0) * dim); });
  par_elems(ne, [&](long long e0, long long e1) { prefault(verts.data() + (size_t)e0 * nvl, (size_t)(e1 - e0) * nvl); });
  coords.resize((size_t)nv * dim);
  verts.resize((size_t)ne * nvl);
  par_elems(nv, [&](long long i0, long long i1) {
    for (long long i = i0; i < i1; i++)
      for (int c = 0; c < dim; c++) coords[(size_t)i * dim + c] = vertices[(size_t)i * vstride + c];
  });
  std::atomic<bool> bad(false);
  par_elems(ne, [&](long long e0, long long e1) {
    bool b = false;
    for (size_t i = (size_t)e0 * nvl; i < (size_t)e1 * nvl; i++) { const int v = elems[i]; verts[i] = v; b |= v < 0 || v >= nv_; }
    if (b) bad.store(true);
  });
  if (bad.load()) return "element vertex index out of range";
  // orientation fix (quirk Q3)
  par_elems(ne, [&](long long e0, long long e1) {
  for (long long e = e0; e < e1; e++) {
    int* vi = &verts[(size_t)e * nvl];
    const double* X = coords.data();
    double det;
    if (dim == 2) {
      const double *v0 = X + 2 * (size_t)vi[0], *v1 = X + 2 * (size_t)vi[1], *v2 = X + 2 * (size_t)vi[2];
      det = (v1[0] - v0[0]) * (v2[1] - v0[1]) - (v1[1] - v0[1]) * (v2[0] - v0[0]);
    } else {
      const double *v0 = X + 3 * (size_t)vi[0], *v1 = X + 3 * (size_t)vi[1], *v2 = X + 3 * (size_t)vi[2], *v3 = X + 3 * (size_t)vi[3];
      double a[3], b[3], c[3];
      for (int k = 0; k < 3; k++) { a[k] = v1[k] - v0[k]; b[k] = v2[k] - v0[k]; c[k] = v3[k] - v0[k]; }
      det = a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
    }
    if (det < 0.0) { int t = vi[0]; vi[0] = vi[1]; vi[1] = t; }
  }
  });
  // edges + connectivity.  P1: the connectivity is the vertex list and the edge numbering (an output of the mesh constructor only) is left to
  // ensure_edges(): its sequential first-appearance walk costs more than every other table together on large meshes.
  conn.clear(); conn.reserve((size_t)ne * d);
  advise_huge_pages(conn.data(), conn.capacity() * sizeof(int));
  par_elems(ne, [&](long long e0, long long e1) { prefault(conn.data() + (size_t)e0 * d, (size_t)(e1 - e0) * d); });
  conn.resize((size_t)ne * d);
  edges_built = false; nedges = 0; edge_lo.clear(); edge_hi.clear();
  if (degree == 1) {
    par_elems(ne, [&](long long e0, long long e1) { memcpy(conn.data() + (size_t)e0 * d, verts.data() + (size_t)e0 * d, (size_t)(e1 - e0) * d * sizeof(int)); });
  } else {
    const int nt = host_threads();
    if (nt > 1 && ne >= 32768 && (long long)ne * nel < 1500000000LL) {          // threaded, sort-based (3 x 16 B per record of scratch)
      std::vector<int> ids;
      number_edges_sorted(verts, ne, nvl, nel, dim == 2 ? kTriEdges : kTetEdges, nt, ids, edge_lo, edge_hi);
      nedges = (long long)edge_lo.size();
      host_parallel_for(ne, nt, [&](long long e0, long long e1) {
        for (long long e = e0; e < e1; e++) {
          const int* vi = &verts[(size_t)e * nvl];
          int* ce = &conn[(size_t)e * d];
          for (int k = 0; k < nvl; k++) ce[k] = vi[k];
          for (int j = 0; j < nel; j++) ce[nvl + j] = nv + ids[(size_t)e * nel + j];
        }
      }, 1 << 14);
      edges_built = true;
    } else {
    EdgeNumbering en(nv, expected_edges());
    for (int e = 0; e < ne; e++) {
      const int* vi = &verts[(size_t)e * nvl];
      int* ce = &conn[(size_t)e * d];
      if (e + 8 < ne) {          // the chain heads of the vertices a few elements ahead (random addresses on a renumbered mesh)
        const int* vn = &verts[(size_t)(e + 8) * nvl];
        for (int k = 0; k < nvl; k++) __builtin_prefetch(&en.head[vn[k]]);
      }
      for (int k = 0; k < nvl; k++) ce[k] = vi[k];
      for (int j = 0; j < nel; j++) {
        int a = dim == 2 ? kTriEdges[j][0] : kTetEdges[j][0], b = dim == 2 ? kTriEdges[j][1] : kTetEdges[j][1];
        ce[nvl + j] = nv + en.id(vi[a], vi[b]);
      }
    }
    nedges = (long long)en.hi.size();
    edge_lo.swap(en.lo);
    edge_hi.swap(en.hi);
    edges_built = true;
    }
  }
  long long nd = degree == 1 ? (long long)nv : (long long)nv + nedges;
  if (nd > 2147483647LL) return "too many dofs for 32-bit dof ids";
  ndof = (int)nd;
  return "";
}

// Euler: E = V + F - 1 for a simply connected triangulation; tetrahedral grids have about 7 edges per vertex, 1.2-1.4 per element
size_t HostMesh::expected_edges() const { return dim == 2 ? (size_t)nv + ne + 64 : (size_t)nv + (size_t)ne * 3 / 2 + 64; }

void HostMesh::ensure_edges() const {
  if (edges_built) return;
  const int nvl = dim + 1, nel = dim == 2 ? 3 : 6;
  const int nt = host_threads();
  if (nt > 1 && ne >= 32768 && (long long)ne * nel < 1500000000LL) {
    std::vector<int> ids;
    number_edges_sorted(verts, ne, nvl, nel, dim == 2 ? kTriEdges : kTetEdges, nt, ids, edge_lo, edge_hi);
    nedges = (long long)edge_lo.size();
    edges_built = true;
    return;
  }
  EdgeNumbering en(nv, expected_edges());
  for (int e = 0; e < ne; e++) {
    const int* vi = &verts[(size_t)e * nvl];
    for (int j = 0; j < nel; j++) en.id(vi[dim == 2 ? kTriEdges[j][0] : kTetEdges[j][0]], vi[dim == 2 ? kTriEdges[j][1] : kTetEdges[j][1]]);
  }
  nedges = (long long)en.hi.size();
  edge_lo.swap(en.lo);
  edge_hi.swap(en.hi);
  edges_built = true;
}

void HostMesh::dof_position(int dof, double* x) const {
  if (dof < nv) {
    for (int c = 0; c < dim; c++) x[c] = coords[(size_t)dof * dim + c];
  } else {
    int e = dof - nv;
    for (int c = 0; c < dim; c++) x[c] = 0.5 * (coords[(size_t)edge_lo[e] * dim + c] + coords[(size_t)edge_hi[e] * dim + c]);
  }
}

namespace {
double tri_area_heron(const double* p1, const double* p2, const double* p3) {
  double a = std::sqrt((p1[0] - p2[0]) * (p1[0] - p2[0]) + (p1[1] - p2[1]) * (p1[1] - p2[1]));
  double b = std::sqrt((p3[0] - p2[0]) * (p3[0] - p2[0]) + (p3[1] - p2[1]) * (p3[1] - p2[1]));
  double c = std::sqrt((p1[0] - p3[0]) * (p1[0] - p3[0]) + (p1[1] - p3[1]) * (p1[1] - p3[1]));
  double s = (a + b + c) / 2.0;
  return std::sqrt(s * (s - a) * (s - b) * (s - c));
}
double tet_volume(const double* v0, const double* v1, const double* v2, const double* v3) {
  double J[3][3];
  const double* V[3] = {v1, v2, v3};
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) J[r][c] = V[c][r] - v0[r];
  double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
               J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  return det * (1. / 6.);
}
}  // namespace

void HostMesh::measure(double* a) const {
  const int nvl = dim + 1;
  par_elems(ne, [&](long long e0, long long e1) {
    for (long long e = e0; e < e1; e++) {
      const int* vi = &verts[(size_t)e * nvl];
      const double* X = coords.data();
      a[e] = dim == 2 ? tri_area_heron(X + 2 * (size_t)vi[0], X + 2 * (size_t)vi[1], X + 2 * (size_t)vi[2])
                      : tet_volume(X + 3 * (size_t)vi[0], X + 3 * (size_t)vi[1], X + 3 * (size_t)vi[2], X + 3 * (size_t)vi[3]);
    }
  });
}

void HostMesh::gauss_weights(double* w) const {
  std::vector<double> a(ne);
  measure(a.data());
  par_elems(ne, [&](long long e0, long long e1) {
    for (long long e = e0; e < e1; e++)
      for (int k = 0; k < g; k++) w[(size_t)e * g + k] = dim == 2 ? rule.w[k] * a[e] / 0.5 : rule.w[k] * a[e] * 6.0;
  });
}

void HostMesh::gauss_points(double* xyz) const {
  const int nvl = dim + 1;
  const size_t G = (size_t)ne * g;
  par_elems(ne, [&](long long e0, long long e1) {
    for (long long e = e0; e < e1; e++) {
      const int* vi = &verts[(size_t)e * nvl];
      for (int k = 0; k < g; k++) {
        double L[4];
        if (dim == 2) { L[0] = 1 - rule.x[k] - rule.y[k]; L[1] = rule.x[k]; L[2] = rule.y[k]; }
        else { L[0] = 1 - rule.x[k] - rule.y[k] - rule.z[k]; L[1] = rule.x[k]; L[2] = rule.y[k]; L[3] = rule.z[k]; }
        for (int c = 0; c < dim; c++) {
          double s = 0;
          for (int j = 0; j < nvl; j++) s += coords[(size_t)vi[j] * dim + c] * L[j];
          xyz[c * G + (size_t)e * g + k] = s;
        }
      }
    }
  });
}

}  // namespace adfem
