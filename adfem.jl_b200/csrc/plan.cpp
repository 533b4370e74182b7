#include "plan.h"
#include "tet_grid_tables.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <type_traits>
#include <utility>

namespace adfem {

int default_threads() {
  if (const char* s = getenv("ADFEM_HOST_THREADS")) { int v = atoi(s); if (v > 0) return v; }
  unsigned n = std::thread::hardware_concurrency();
  return n == 0 ? 4 : (int)std::min(n, 64u);
}

namespace {

// blocks [cut[t], cut[t+1]) of [0,n) with equal shares of the weight whose running sum is prefix(i) (non-decreasing, prefix(0) = 0): P2 dofs are
// numbered vertices first, and a vertex row costs three times an edge row
template <class W> void parallel_for_weighted(long long n, int nthreads, W prefix, const std::function<void(long long, long long, int)>& fn) {
  if (nthreads <= 1 || n < 4096) { fn(0, n, 0); return; }
  std::vector<long long> cut(nthreads + 1, n);
  cut[0] = 0;
  const long long total = prefix(n);
  for (int t = 1; t < nthreads; t++) {
    const long long want = (long long)((double)total * t / nthreads);
    long long lo = cut[t - 1], hi = n;
    while (lo < hi) { const long long mid = (lo + hi) / 2; if (prefix(mid) < want) lo = mid + 1; else hi = mid; }
    cut[t] = lo;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) if (cut[t] < cut[t + 1]) th.emplace_back(fn, cut[t], cut[t + 1], t);
  for (auto& x : th) x.join();
}

// ADFEM_DEBUG_PLAN=1: wall time of each host phase on stderr
struct PhaseTimer {
  const bool on = getenv("ADFEM_DEBUG_PLAN") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    const auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "  [plan] %-34s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  }
};

// v = n zeros, with the first touch of the pages spread over the worker threads: a fresh multi-GB array costs a page fault per 4 KB, which a single
// thread pays at 1-2 GB/s.  The threads zero the reserved storage, the resize that follows value-initialises pages that are already mapped.
template <class T> void assign_parallel(std::vector<T>& v, size_t n, int nthreads) {
  static_assert(std::is_trivially_copyable<T>::value, "plain data only");
  v.clear();
  const size_t bytes = n * sizeof(T);
  if (nthreads > 1 && bytes >= ((size_t)8 << 20)) {
    v.reserve(n);
    char* p = reinterpret_cast<char*>(v.data());
    advise_huge_pages(p, bytes);
    std::vector<std::thread> th;
    const size_t chunk = ((bytes + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
    for (size_t b = 0; b < bytes; b += chunk) th.emplace_back([=] { memset(p + b, 0, std::min(chunk, bytes - b)); });
    for (auto& x : th) x.join();
  }
  v.resize(n);
}

// static block partition of [0,n) over worker threads
void parallel_for(long long n, int nthreads, const std::function<void(long long, long long, int)>& fn) {
  if (nthreads <= 1 || n < 4096) { fn(0, n, 0); return; }
  std::vector<std::thread> th;
  long long chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    long long b = (long long)t * chunk, e = std::min(n, b + chunk);
    if (b >= e) break;
    th.emplace_back(fn, b, e, t);
  }
  for (auto& x : th) x.join();
}

inline uint64_t spread2(uint64_t x) {   // 21 bits -> every 2nd bit
  x &= 0x1fffff;
  x = (x | x << 16) & 0x0000ffff0000ffffULL; x = (x | x << 8) & 0x00ff00ff00ff00ffULL;
  x = (x | x << 4) & 0x0f0f0f0f0f0f0f0fULL;  x = (x | x << 2) & 0x3333333333333333ULL;
  x = (x | x << 1) & 0x5555555555555555ULL;
  return x;
}
inline uint64_t spread3(uint64_t x) {   // 21 bits -> every 3rd bit
  x &= 0x1fffff;
  x = (x | x << 32) & 0x1f00000000ffffULL; x = (x | x << 16) & 0x1f0000ff0000ffULL;
  x = (x | x << 8) & 0x100f00f00f00f00fULL; x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
  x = (x | x << 2) & 0x1249249249249249ULL;
  return x;
}

struct Morton {
  int dim; double lo[3], inv[3];
  Morton(const HostMesh& m) : dim(m.dim) {
    double hi[3];
    for (int c = 0; c < dim; c++) { lo[c] = 1e300; hi[c] = -1e300; }
    for (int i = 0; i < m.nv; i++)
      for (int c = 0; c < dim; c++) { double v = m.coords[(size_t)i * dim + c]; lo[c] = std::min(lo[c], v); hi[c] = std::max(hi[c], v); }
    // one common scale so that cells stay isotropic
    double ext = 0;
    for (int c = 0; c < dim; c++) ext = std::max(ext, hi[c] - lo[c]);
    for (int c = 0; c < dim; c++) inv[c] = ext > 0 ? 2097151.0 / ext : 0.0;
  }
  uint64_t code(const double* x) const {
    uint64_t q[3];
    for (int c = 0; c < dim; c++) { double t = (x[c] - lo[c]) * inv[c]; q[c] = t <= 0 ? 0 : (t >= 2097151.0 ? 2097151 : (uint64_t)t); }
    return dim == 2 ? (spread2(q[0]) | spread2(q[1]) << 1) : (spread3(q[0]) | spread3(q[1]) << 1 | spread3(q[2]) << 2);
  }
};

// ids sorted by (Morton code of their position, id)
void morton_order(long long n, int nthreads, const std::function<uint64_t(long long)>& code, std::vector<int>& order) {
  std::vector<KeyId> keyed;
  assign_parallel(keyed, (size_t)n, nthreads);
  parallel_for(n, nthreads, [&](long long b, long long e, int) { for (long long i = b; i < e; i++) keyed[i] = KeyId{code(i), (int)i}; });
  if (nthreads > 1 && n >= 65536) radix_sort_by_key(keyed, nthreads);
  else std::sort(keyed.begin(), keyed.end());
  assign_parallel(order, (size_t)n, nthreads);
  parallel_for(n, nthreads, [&](long long b, long long e, int) { for (long long i = b; i < e; i++) order[i] = keyed[i].id; });
}

}  // namespace

// ------------------------------------------------------------------------------------------------
std::string ScalarPattern::build(const HostMesh& m, int nthreads) {
  n = m.ndof;
  const int d = m.d;
  PhaseTimer pt;
  const long long nslot_rows = (long long)m.ne * d;
  if (nslot_rows > 2147483647LL) return "ne*elem_ndof exceeds 32-bit";
  // dof -> (element, local) adjacency, ascending element inside a row.  Threads take element blocks: counts and fill positions are claimed with
  // relaxed atomic increments, which leaves the order inside a row to the scheduler — so every row (a handful of entries) is sorted afterwards.
  assign_parallel(adj_ptr, (size_t)n + 1, nthreads);
  parallel_for(nslot_rows, nthreads, [&](long long b, long long e, int) {
    long long* cnt = adj_ptr.data() + 1;
    const int* c = m.conn.data();
    for (long long i = b; i < e; i++) {
      if (i + 24 < e) __builtin_prefetch(&cnt[c[i + 24]], 1);
      __atomic_fetch_add(&cnt[c[i]], 1LL, __ATOMIC_RELAXED);
    }
  });
  for (int r = 0; r < n; r++) adj_ptr[r + 1] += adj_ptr[r];
  assign_parallel(adj_elem, (size_t)nslot_rows, nthreads); assign_parallel(adj_loc, (size_t)nslot_rows, nthreads);
  {
    std::vector<long long> cur;
    assign_parallel(cur, (size_t)n, nthreads);
    parallel_for(n, nthreads, [&](long long b, long long e, int) { memcpy(cur.data() + b, adj_ptr.data() + b, (size_t)(e - b) * sizeof(long long)); });
    parallel_for(m.ne, nthreads, [&](long long b, long long e, int) {
      const int* c = m.conn.data() + (size_t)b * d;
      const int* cend = m.conn.data() + (size_t)e * d;
      for (long long el = b; el < e; el++)
        for (int p = 0; p < d; p++, c++) {
          if (c + 24 < cend) __builtin_prefetch(&cur[c[24]], 1);
          const long long at = __atomic_fetch_add(&cur[*c], 1LL, __ATOMIC_RELAXED); adj_elem[at] = (int)el; adj_loc[at] = (uint8_t)p;
        }
    });
  }
  parallel_for(n, nthreads, [&](long long b, long long e, int) {
    std::vector<uint64_t> key;
    for (long long r = b; r < e; r++) {
      const long long r0 = adj_ptr[r], len = adj_ptr[r + 1] - r0;
      int* el = adj_elem.data() + r0;
      uint8_t* lo = adj_loc.data() + r0;
      if (len <= 32) {                          // insertion sort of the (element, local) pairs (local decides only inside a degenerate element)
        for (long long i = 1; i < len; i++) {
          const int ke = el[i]; const uint8_t kl = lo[i];
          long long j = i - 1;
          for (; j >= 0 && (el[j] > ke || (el[j] == ke && lo[j] > kl)); j--) { el[j + 1] = el[j]; lo[j + 1] = lo[j]; }
          el[j + 1] = ke; lo[j + 1] = kl;
        }
      } else {
        key.resize((size_t)len);
        for (long long i = 0; i < len; i++) key[i] = (uint64_t)(uint32_t)el[i] << 8 | lo[i];
        std::sort(key.begin(), key.end());
        for (long long i = 0; i < len; i++) { el[i] = (int)(key[i] >> 8); lo[i] = (uint8_t)(key[i] & 0xff); }
      }
    }
  });
  {
    std::atomic<bool> rep(false);
    parallel_for(m.ne, nthreads, [&](long long b, long long e, int) {
      bool r = false;
      for (long long el = b; el < e; el++) {
        const int* ce = &m.conn[(size_t)el * d];
        for (int p = 1; p < d; p++) for (int q = 0; q < p; q++) r |= ce[p] == ce[q];
      }
      if (r) rep.store(true);
    });
    repeated_dofs = rep.load();
  }
  pt.lap("pattern: adjacency");
  if (pt.on) {      // FNV-1a over the adjacency: lets two builds of the host code be compared
    uint64_t hsh = 1469598103934665603ULL;
    auto mix = [&](uint64_t v) { hsh = (hsh ^ v) * 1099511628211ULL; };
    for (long long v : adj_ptr) mix((uint64_t)v);
    for (long long i = 0; i < nslot_rows; i++) mix(((uint64_t)(uint32_t)adj_elem[i] << 8) | adj_loc[i]);
    fprintf(stderr, "  [plan] adjacency checksum %016llx\n", (unsigned long long)hsh);
    pt.lap("pattern: (checksum)");
  }
  // pass 1: the distinct columns of every row, ascending (one sort of the row's column ids); a thread appends the rows of its range to a buffer of
  // its own, which is copied to its place once the row pointers are known
  std::vector<int> rowlen;
  assign_parallel(rowlen, (size_t)n, nthreads);
  const int nbuf = std::max(1, nthreads);
  std::vector<std::vector<int>> colbuf(nbuf);
  std::vector<long long> buf_first(nbuf, -1);
  auto row_work = [&](long long r) { return adj_ptr[r] + r; };       // running sum of (incident elements + 1) per row
  parallel_for_weighted(n, nthreads, row_work, [&](long long b, long long e, int tid) {
    std::vector<int> cols;
    std::vector<int>& out = colbuf[tid];
    buf_first[tid] = b;
    // untouched reserve (virtual memory only): a row has at most one column per gathered dof; twice the usual row length of the element family
    out.reserve((size_t)std::min<long long>((adj_ptr[e] - adj_ptr[b]) * d, (e - b) * (m.dim == 2 ? (m.degree == 1 ? 16 : 40) : (m.degree == 1 ? 32 : 64))));
    advise_huge_pages(out.data(), out.capacity() * sizeof(int));
    constexpr long long AHEAD = 12;      // rows: on a renumbered mesh every incident element's connectivity is a cache miss; ask for it early
    for (long long r = b; r < e; r++) {
      if (r + AHEAD < e)
        for (long long a = adj_ptr[r + AHEAD]; a < adj_ptr[r + AHEAD + 1]; a++) __builtin_prefetch(&m.conn[(size_t)adj_elem[a] * d]);
      const long long deg = adj_ptr[r + 1] - adj_ptr[r];
      if (deg * d <= 40) {                      // the usual row: insert into a small sorted set
        int uq[160], cnt = 0;
        for (long long a = adj_ptr[r]; a < adj_ptr[r + 1]; a++) {
          const int* ce = &m.conn[(size_t)adj_elem[a] * d];
          for (int q = 0; q < d; q++) {
            const int c = ce[q];
            int i = 0, dup = 0;                       // position and presence by branch-free counts over the (short) set
            for (int k = 0; k < cnt; k++) { i += uq[k] < c; dup += uq[k] == c; }
            if (dup) continue;
            for (int k = cnt; k > i; k--) uq[k] = uq[k - 1];
            uq[i] = c; cnt++;
          }
        }
        out.insert(out.end(), uq, uq + cnt);
        rowlen[r] = cnt;
        continue;
      }
      cols.clear();
      for (long long a = adj_ptr[r]; a < adj_ptr[r + 1]; a++) {
        const int* ce = &m.conn[(size_t)adj_elem[a] * d];
        cols.insert(cols.end(), ce, ce + d);
      }
      std::sort(cols.begin(), cols.end());
      const size_t before = out.size();
      for (size_t i = 0; i < cols.size(); i++) if (i == 0 || cols[i] != cols[i - 1]) out.push_back(cols[i]);
      rowlen[r] = (int)(out.size() - before);
    }
  });
  assign_parallel(rowptr, (size_t)n + 1, nthreads);
  for (int r = 0; r < n; r++) rowptr[r + 1] = rowptr[r] + rowlen[r];
  nnz = rowptr[n];
  if (nnz > 4294967295LL) return "scalar nnz exceeds 32-bit slot map";
  pt.lap("pattern: columns of every row");
  assign_parallel(colind, (size_t)nnz, nthreads);
  assign_parallel(slot_nnz, (size_t)m.ne * d * d, nthreads);
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nbuf; t++)
      if (buf_first[t] >= 0 && !colbuf[t].empty())
        th.emplace_back([&, t] { memcpy(colind.data() + rowptr[buf_first[t]], colbuf[t].data(), colbuf[t].size() * sizeof(int)); std::vector<int>().swap(colbuf[t]); });
    for (auto& x : th) x.join();
  }
  // pass 2: slot map — local entry (e, p, q) lands in row conn[e][p] at the position of conn[e][q] among the row's columns
  parallel_for_weighted(n, nthreads, row_work, [&](long long b, long long e, int) {
    constexpr long long AHEAD = 12;
    for (long long r = b; r < e; r++) {
      if (r + AHEAD < e)
        for (long long a = adj_ptr[r + AHEAD]; a < adj_ptr[r + AHEAD + 1]; a++) {
          __builtin_prefetch(&m.conn[(size_t)adj_elem[a] * d]);
          __builtin_prefetch(&slot_nnz[((size_t)adj_elem[a] * d + adj_loc[a]) * d], 1);
        }
      const int* cb = colind.data() + rowptr[r];
      const int* cend = colind.data() + rowptr[r + 1];
      for (long long a = adj_ptr[r]; a < adj_ptr[r + 1]; a++) {
        const int el = adj_elem[a], p = adj_loc[a];
        const int* ce = &m.conn[(size_t)el * d];
        uint32_t* dst = &slot_nnz[((size_t)el * d + p) * d];
        const int len = (int)(cend - cb);
        if (len <= 64) {          // short row: count the smaller columns (no branches, vectorised) instead of a search with unpredictable ones
          for (int q = 0; q < d; q++) {
            const int c = ce[q];
            int pos = 0;
            for (int k = 0; k < len; k++) pos += cb[k] < c;
            dst[q] = (uint32_t)(rowptr[r] + pos);
          }
        } else {
          for (int q = 0; q < d; q++) dst[q] = (uint32_t)(rowptr[r] + (std::lower_bound(cb, cend, ce[q]) - cb));
        }
      }
    }
  });
  pt.lap("pattern: columns + slot map");
  return "";
}

// ------------------------------------------------------------------------------------------------
namespace {
// node (i, j) of Mesh(gm, gn, h): incident (element, local vertex) pairs in ascending element order and the ascending columns of its row
struct GridNode {
  int nadj = 0, ncol = 0;
  int elem[6]; uint8_t loc[6]; int col[7];
  GridNode(int i, int j, int gm, int gn) {
    const long long r = (long long)i * (gm + 1) + j;
    auto cell = [&](int ci, int cj) { return 2 * ((long long)ci * gm + cj); };
    if (i >= 1 && j >= 1) { elem[nadj] = (int)(cell(i - 1, j - 1) + 1); loc[nadj++] = 2; }            // top-right corner of the cell below-left
    if (i >= 1 && j < gm) { elem[nadj] = (int)cell(i - 1, j); loc[nadj++] = 2; elem[nadj] = (int)(cell(i - 1, j) + 1); loc[nadj++] = 0; }    // top-left corner
    if (i < gn && j >= 1) { elem[nadj] = (int)cell(i, j - 1); loc[nadj++] = 1; elem[nadj] = (int)(cell(i, j - 1) + 1); loc[nadj++] = 1; }    // bottom-right corner
    if (i < gn && j < gm) { elem[nadj] = (int)cell(i, j); loc[nadj++] = 0; }                            // bottom-left corner
    if (i >= 1) col[ncol++] = (int)(r - (gm + 1));
    if (i >= 1 && j < gm) col[ncol++] = (int)(r - (gm + 1) + 1);
    if (j >= 1) col[ncol++] = (int)(r - 1);
    col[ncol++] = (int)r;
    if (j < gm) col[ncol++] = (int)(r + 1);
    if (i < gn && j >= 1) col[ncol++] = (int)(r + (gm + 1) - 1);
    if (i < gn) col[ncol++] = (int)(r + (gm + 1));
  }
  int position(int c) const { int p = 0; while (p < ncol && col[p] != c) p++; return p; }
};
}  // namespace

std::string ScalarPattern::build_tri_grid(const HostMesh& m, int gm, int gn, int nthreads) {
  if (m.dim != 2 || m.degree != 1 || m.d != 3 || gm < 1 || gn < 1 || (long long)(gm + 1) * (gn + 1) != m.ndof || 2LL * gm * gn != m.ne)
    return "not the structured triangulation";
  PhaseTimer pt;
  n = m.ndof;
  const long long nslot_rows = 3LL * m.ne, w = gm + 1;
  if (nslot_rows > 2147483647LL) return "ne*elem_ndof exceeds 32-bit";
  assign_parallel(adj_ptr, (size_t)n + 1, nthreads);
  assign_parallel(rowptr, (size_t)n + 1, nthreads);
  parallel_for(n, nthreads, [&](long long b, long long e, int) {
    for (long long r = b; r < e; r++) { const GridNode g((int)(r / w), (int)(r % w), gm, gn); adj_ptr[r + 1] = g.nadj; rowptr[r + 1] = g.ncol; }
  });
  for (long long r = 0; r < n; r++) { adj_ptr[r + 1] += adj_ptr[r]; rowptr[r + 1] += rowptr[r]; }
  nnz = rowptr[n];
  if (adj_ptr[n] != nslot_rows) return "structured triangulation: incidence count mismatch";
  if (nnz > 4294967295LL) return "scalar nnz exceeds 32-bit slot map";
  assign_parallel(adj_elem, (size_t)nslot_rows, nthreads); assign_parallel(adj_loc, (size_t)nslot_rows, nthreads);
  assign_parallel(colind, (size_t)nnz, nthreads);
  assign_parallel(slot_nnz, (size_t)m.ne * 9, nthreads);
  parallel_for(n, nthreads, [&](long long b, long long e, int) {
    for (long long r = b; r < e; r++) {
      const GridNode g((int)(r / w), (int)(r % w), gm, gn);
      for (int a = 0; a < g.nadj; a++) { adj_elem[adj_ptr[r] + a] = g.elem[a]; adj_loc[adj_ptr[r] + a] = g.loc[a]; }
      for (int c = 0; c < g.ncol; c++) colind[rowptr[r] + c] = g.col[c];
    }
  });
  pt.lap("pattern (structured): adjacency, columns");
  // slot map: local entry (p, q) of a triangle lands in the row of its vertex p at the position of its vertex q
  parallel_for((long long)gn * gm, nthreads, [&](long long b, long long e, int) {
    for (long long c = b; c < e; c++) {
      const int ci = (int)(c / gm), cj = (int)(c % gm);
      const GridNode bl(ci, cj, gm, gn), br(ci, cj + 1, gm, gn), tl(ci + 1, cj, gm, gn), tr(ci + 1, cj + 1, gm, gn);
      const long long a = (long long)ci * w + cj;
      const GridNode* node[2][3] = {{&bl, &br, &tl}, {&tl, &br, &tr}};
      const long long vert[2][3] = {{a, a + 1, a + w}, {a + w, a + 1, a + w + 1}};
      for (int t = 0; t < 2; t++)
        for (int p = 0; p < 3; p++)
          for (int q = 0; q < 3; q++)
            slot_nnz[((size_t)(2 * c + t) * 3 + p) * 3 + q] = (uint32_t)(rowptr[vert[t][p]] + node[t][p]->position((int)vert[t][q]));
    }
  });
  pt.lap("pattern (structured): slot map");
  return "";
}

std::string ScalarPattern::build_tet_grid(const HostMesh& m, int gn, int gl, int nthreads) {
  const long long n1 = gn + 1, plane = n1 * n1;
  if (m.dim != 3 || m.degree != 1 || m.d != 4 || gn < 1 || gl < 1 || plane * (gl + 1) != m.ndof || 5LL * gn * gn * gl != m.ne)
    return "not the structured tetrahedral grid";
  PhaseTimer pt;
  TetGridTables T;
  build_tet_grid_tables(T);
  n = m.ndof;
  const long long nslot_rows = 4LL * m.ne;
  if (nslot_rows > 2147483647LL) return "ne*elem_ndof exceeds 32-bit";
  // incident tetrahedra (table entries whose cube exists) and the 27-bit mask of the row entries of node (i, j, k)
  auto node = [&](long long r, int& i, int& j, int& k) { i = (int)(r % n1); j = (int)((r / n1) % n1); k = (int)(r / plane); };
  auto exists = [&](const int* inc, int i, int j, int k) {
    const int ci = i + inc[0], cj = j + inc[1], ck = k + inc[2];
    return ci >= 0 && ci < gn && cj >= 0 && cj < gn && ck >= 0 && ck < gl;
  };
  auto popc = [](unsigned v) { return __builtin_popcount(v); };
  assign_parallel(adj_ptr, (size_t)n + 1, nthreads);
  assign_parallel(rowptr, (size_t)n + 1, nthreads);
  std::vector<uint32_t> maskv;
  assign_parallel(maskv, (size_t)n, nthreads);
  parallel_for(n, nthreads, [&](long long b, long long e, int) {
    for (long long r = b; r < e; r++) {
      int i, j, k;
      node(r, i, j, k);
      const int par = (i + j + k) & 1;
      int cnt = 0; unsigned mask = 0;
      for (int t = 0; t < T.ninc[par]; t++)
        if (exists(T.inc[par][t], i, j, k)) { cnt++; for (int q = 0; q < 4; q++) mask |= 1u << T.vslot[par][t][q]; }
      adj_ptr[r + 1] = cnt; rowptr[r + 1] = popc(mask); maskv[r] = mask;
    }
  });
  for (long long r = 0; r < n; r++) { adj_ptr[r + 1] += adj_ptr[r]; rowptr[r + 1] += rowptr[r]; }
  nnz = rowptr[n];
  if (adj_ptr[n] != nslot_rows) return "structured tetrahedral grid: incidence count mismatch";
  if (nnz > 4294967295LL) return "scalar nnz exceeds 32-bit slot map";
  pt.lap("pattern (structured tetrahedra): counts");
  assign_parallel(adj_elem, (size_t)nslot_rows, nthreads); assign_parallel(adj_loc, (size_t)nslot_rows, nthreads);
  assign_parallel(colind, (size_t)nnz, nthreads);
  assign_parallel(slot_nnz, (size_t)m.ne * 16, nthreads);
  pt.lap("pattern (structured tetrahedra): allocation");
  std::atomic<bool> bad(false);
  // per node: columns from the mask, incident tetrahedra from the table (the node's local position is read from the mesh)
  parallel_for(n, nthreads, [&](long long b, long long e, int) {
    for (long long r = b; r < e; r++) {
      int i, j, k;
      node(r, i, j, k);
      const int par = (i + j + k) & 1;
      const unsigned mask = maskv[r];
      long long at = rowptr[r];
      for (int s = 0; s < 27; s++)
        if ((mask >> s) & 1) colind[at++] = (int)(r + (long long)(s / 9 - 1) * plane + (long long)((s / 3) % 3 - 1) * n1 + (s % 3 - 1));
      long long a = adj_ptr[r];
      for (int t = 0; t < T.ninc[par]; t++) {
        const int* inc = T.inc[par][t];
        if (!exists(inc, i, j, k)) continue;
        const long long cube = ((long long)(i + inc[0]) * gn + (j + inc[1])) * gl + (k + inc[2]);
        const long long el = 5 * cube + inc[3];
        const int* v = &m.conn[(size_t)el * 4];
        int p = -1;
        for (int q = 0; q < 4; q++) if (v[q] == r) p = q;
        if (p < 0) { bad.store(true); return; }
        adj_elem[a] = (int)el; adj_loc[a] = (uint8_t)p; a++;
      }
    }
  });
  pt.lap("pattern (structured tetrahedra): adjacency, columns");
  // per element (one cache line of the slot map each): local entry (p, q) lands in the row of vertex p at the position of vertex q among the row
  // entries — its 27-neighbourhood slot (dk+1) 9 + (dj+1) 3 + (di+1), counted through the row's mask
  parallel_for(m.ne, nthreads, [&](long long b, long long e, int) {
    for (long long el = b; el < e; el++) {
      const int* v = &m.conn[(size_t)el * 4];
      int c[4][3];
      for (int q = 0; q < 4; q++) node(v[q], c[q][0], c[q][1], c[q][2]);
      for (int p = 0; p < 4; p++) {
        const unsigned mask = maskv[v[p]];
        const long long rs = rowptr[v[p]];
        for (int q = 0; q < 4; q++) {
          const int di = c[q][0] - c[p][0], dj = c[q][1] - c[p][1], dk = c[q][2] - c[p][2];
          if (di < -1 || di > 1 || dj < -1 || dj > 1 || dk < -1 || dk > 1) { bad.store(true); return; }
          const int slot = (dk + 1) * 9 + (dj + 1) * 3 + (di + 1);
          if (!((mask >> slot) & 1)) { bad.store(true); return; }
          slot_nnz[((size_t)el * 4 + p) * 4 + q] = (uint32_t)(rs + popc(mask & ((1u << slot) - 1)));
        }
      }
    }
  });
  if (bad.load()) return "structured tetrahedral grid: an element does not hold the generator's vertices";
  pt.lap("pattern (structured tetrahedra): slot map");
  return "";
}

// ------------------------------------------------------------------------------------------------
namespace {
struct BlobWriter {
  std::vector<uint8_t>& out;
  size_t base;
  explicit BlobWriter(std::vector<uint8_t>& o) : out(o), base(o.size()) {}
  template <class T> void section(const T* data, size_t count) {
    size_t at = out.size(), bytes = count * sizeof(T), padded = align16(bytes);
    out.resize(at + padded, 0);
    if (bytes) memcpy(out.data() + at, data, bytes);
  }
  template <class T> void section(const std::vector<T>& v) { section(v.data(), v.size()); }
  size_t size() const { return out.size() - base; }
};

// vertices (post orientation fix) used by a sorted element list, and the k-major local ids
void tile_vertices(const HostMesh& m, const std::vector<int>& te, std::vector<int>& tvert, std::vector<uint16_t>& tv, std::vector<double>& xy) {
  const int nvl = m.dim + 1, nel = (int)te.size();
  tvert.clear();
  for (int e : te) for (int k = 0; k < nvl; k++) tvert.push_back(m.verts[(size_t)e * nvl + k]);
  std::sort(tvert.begin(), tvert.end());
  tvert.erase(std::unique(tvert.begin(), tvert.end()), tvert.end());
  tv.resize((size_t)nvl * nel);
  for (int le = 0; le < nel; le++)
    for (int k = 0; k < nvl; k++)
      tv[(size_t)k * nel + le] = (uint16_t)(std::lower_bound(tvert.begin(), tvert.end(), m.verts[(size_t)te[le] * nvl + k]) - tvert.begin());
  xy.resize(tvert.size() * m.dim);
  for (size_t i = 0; i < tvert.size(); i++) __builtin_prefetch(&m.coords[(size_t)tvert[i] * m.dim]);
  for (size_t i = 0; i < tvert.size(); i++)
    for (int c = 0; c < m.dim; c++) xy[i * m.dim + c] = m.coords[(size_t)tvert[i] * m.dim + c];
}

struct PartOut {
  std::vector<uint8_t> blob; std::vector<long long> sizes; std::string err;
  int max_rows = 0, max_elems = 0, max_nnz = 0, max_src = 0, max_verts = 0;
  size_t max_head = 0, max_body = 0;
  long long tot_a = 0;
};

// per_tile appends the tile's head and body to P.blob and pushes their two sizes to P.sizes
// A failure in one part (a tile that cannot be built, or too_big() of the part's running maxima — the plan's maxima can only be larger) stops the
// other parts at their next tile.
template <class F> std::string run_parts(int ntiles, int nthreads, std::vector<PartOut>& parts,
                                         const std::function<bool(size_t, size_t, int, int)>& too_big, F per_tile) {
  int nparts = std::max(1, std::min(nthreads, ntiles));
  parts.assign(nparts, PartOut());
  std::vector<long long> pcut(nparts + 1);
  for (int i = 0; i <= nparts; i++) pcut[i] = (long long)ntiles * i / nparts;
  std::atomic<bool> stop(false);
  std::vector<std::thread> th;
  for (int pi = 0; pi < nparts; pi++) th.emplace_back([&, pi] {
    PartOut& P = parts[pi];
    for (long long t = pcut[pi]; t < pcut[pi + 1] && P.err.empty() && !stop.load(std::memory_order_relaxed); t++) {
      per_tile((int)t, P);
      if (P.err.empty() && P.sizes.size() >= 2) {
        P.max_head = std::max(P.max_head, (size_t)P.sizes[P.sizes.size() - 2]);
        P.max_body = std::max(P.max_body, (size_t)P.sizes[P.sizes.size() - 1]);
        if (too_big && too_big(P.max_head, P.max_body, P.max_elems, P.max_nnz)) P.err = "tile too large";
      }
      if (!P.err.empty()) stop.store(true, std::memory_order_relaxed);
    }
  });
  for (auto& t : th) t.join();
  for (auto& P : parts) if (!P.err.empty()) return P.err;
  return "";
}

// concatenate the per-thread outputs: blob, blob_ptr (2 offsets per tile + end), largest head / body
void merge_parts(std::vector<PartOut>& parts, std::vector<long long>& blob_ptr, std::vector<uint8_t>& blob, size_t& max_head, size_t& max_body, int nthreads) {
  blob_ptr.assign(1, 0); max_head = max_body = 0;
  size_t total = 0;
  std::vector<size_t> at;
  for (auto& P : parts) { at.push_back(total); total += P.blob.size(); }
  assign_parallel(blob, total, nthreads);
  for (auto& P : parts)
    for (size_t i = 0; i < P.sizes.size(); i++) {
      blob_ptr.push_back(blob_ptr.back() + P.sizes[i]);
      size_t& mx = (i & 1) ? max_body : max_head;
      mx = std::max(mx, (size_t)P.sizes[i]);
    }
  std::vector<std::thread> th;
  for (size_t i = 0; i < parts.size(); i++)
    th.emplace_back([&, i] { if (!parts[i].blob.empty()) memcpy(blob.data() + at[i], parts[i].blob.data(), parts[i].blob.size()); std::vector<uint8_t>().swap(parts[i].blob); });
  for (auto& x : th) x.join();
}
}  // namespace

std::string FwdTiles::build(const HostMesh& m, const ScalarPattern& pat, int R, int max_tile_elems, int sym_, int nthreads) {
  rows_per_tile = R; sym = sym_;
  const int d = m.d, dd = d * d, n = pat.n;
  const int nslot = sym ? d * (d + 1) / 2 : dd;
  PhaseTimer pt;
  if ((long long)morton_cache.size() != n) {
    Morton mc(m);
    morton_order(n, nthreads, [&](long long i) { double x[3]; m.dof_position((int)i, x); return mc.code(x); }, morton_cache);
  }
  std::vector<int> order(morton_cache);
  pt.lap("fwd tiles: morton order of rows");
  ntiles = (n + R - 1) / R;
  std::vector<int> row_ptr(ntiles + 1);
  for (int t = 0; t <= ntiles; t++) row_ptr[t] = (int)std::min<long long>((long long)t * R, n);
  std::vector<int>& rows = order;
  parallel_for(ntiles, nthreads, [&](long long b, long long e, int) {
    for (long long t = b; t < e; t++) std::sort(rows.begin() + row_ptr[t], rows.begin() + row_ptr[t + 1]);
  });
  // destinations are (row, position) packed into 16 bits when both fit in a byte, else 32 bits
  int max_len = 0;
  for (int r = 0; r < n; r++) max_len = std::max<int>(max_len, (int)(pat.rowptr[r + 1] - pat.rowptr[r]));
  ent32 = (R > 256 || max_len > 256) ? 1 : 0;
  std::vector<int> symidx(dd);
  for (int p = 0; p < d; p++) for (int q = 0; q < d; q++) { int a = std::min(p, q), b = std::max(p, q); symidx[p * d + q] = a * d - a * (a - 1) / 2 + (b - a); }

  pt.lap("fwd tiles: tile sort, max row");
  std::vector<PartOut> parts;
  std::string err = run_parts(ntiles, nthreads, parts, too_big, [&](int t, PartOut& P) {
    std::vector<int> te, tvert, jitem, hkey, hval;
    std::vector<uint16_t> tv, rlen, pool;
    std::vector<double> xy;
    std::vector<uint32_t> rstart;
    struct Item { uint32_t d0, d1; int paired; uint32_t off, cnt; };     // sources: pool[off .. off + cnt)
    std::vector<Item> items;
    const int* trow = rows.data() + row_ptr[t];
    const int nrows = row_ptr[t + 1] - row_ptr[t];
    // The rows of a tile are neighbours in space, not in memory: on a renumbered mesh every row and every element below is a cache miss of its
    // own.  Ask for them a whole tile at a time (the loads are independent) instead of meeting them one after the other.
    for (int i = 0; i < nrows; i++) { __builtin_prefetch(&pat.adj_ptr[trow[i]]); __builtin_prefetch(&pat.rowptr[trow[i]]); }
    for (int i = 0; i < nrows; i++) {
      const int r = trow[i];
      __builtin_prefetch(&pat.adj_elem[pat.adj_ptr[r]]); __builtin_prefetch(&pat.adj_loc[pat.adj_ptr[r]]);
      __builtin_prefetch(&pat.colind[pat.rowptr[r]]); __builtin_prefetch(&pat.colind[pat.rowptr[r + 1] - 1]);
    }
    for (int i = 0; i < nrows; i++) { int r = trow[i]; for (long long a = pat.adj_ptr[r]; a < pat.adj_ptr[r + 1]; a++) te.push_back(pat.adj_elem[a]); }
    std::sort(te.begin(), te.end());
    te.erase(std::unique(te.begin(), te.end()), te.end());
    const int nel = (int)te.size();
    if (nel > max_tile_elems || (long long)nel * (sym ? nslot : dd) > 65535) { P.err = "tile too large"; return; }
    for (int el : te) {
      __builtin_prefetch(&m.verts[(size_t)el * (m.dim + 1)]);
      for (int b = 0; b < dd * 4; b += 64) __builtin_prefetch(reinterpret_cast<const char*>(&pat.slot_nnz[(size_t)el * dd]) + b);
    }
    tile_vertices(m, te, tvert, tv, xy);
    if (tvert.size() > 65535) { P.err = "tile too large"; return; }
    auto code = [&](int lr, int j) { return ent32 ? ((uint32_t)lr | (uint32_t)j << 16) : ((uint32_t)lr | (uint32_t)j << 8); };
    // dof id -> tile row (scalar plans ask it for every CSR entry): open addressing, at most half full
    int hbits = 4;
    while ((1 << hbits) < 2 * nrows) hbits++;
    hkey.assign((size_t)1 << hbits, -1); hval.resize((size_t)1 << hbits);
    const uint32_t hmask = (1u << hbits) - 1;
    auto hslot = [&](int c) { return ((uint32_t)c * 2654435761u) >> (32 - hbits); };
    if (sym)
      for (int lr = 0; lr < nrows; lr++) {
        uint32_t at = hslot(trow[lr]);
        while (hkey[at] >= 0) at = (at + 1) & hmask;
        hkey[at] = trow[lr]; hval[at] = lr;
      }
    auto tile_row_of = [&](int c) {          // -1: not a row of this tile
      for (uint32_t at = hslot(c);; at = (at + 1) & hmask) {
        if (hkey[at] == c) return hval[at];
        if (hkey[at] < 0) return -1;
      }
    };
    size_t nsrc = 0, nnz_t = 0;
    for (int lr = 0; lr < nrows; lr++) {
      const int r = trow[lr];
      const long long rs = pat.rowptr[r];
      const int len = (int)(pat.rowptr[r + 1] - rs), deg = (int)(pat.adj_ptr[r + 1] - pat.adj_ptr[r]);
      rstart.push_back((uint32_t)rs);
      rlen.push_back((uint16_t)len);
      nnz_t += (size_t)len;
      // one item per CSR entry of the row (ascending column) unless its mirror image was emitted from the other row of the pair; an entry
      // receives at most one contribution per incident element (d of them if elements repeat dofs), so item j owns pool[base + j*cap .. + cap)
      jitem.assign(len, -1);
      const size_t base = pool.size();
      const int cap = pat.repeated_dofs ? deg * d : deg;       // a degenerate element can feed an entry from several of its local columns
      pool.resize(base + (size_t)len * cap);
      for (int j = 0; j < len; j++) {
        const int c = pat.colind[rs + j];
        int paired = 0; uint32_t d1 = 0;
        if (sym && c != r) {                       // is the mirrored entry (c, r) produced by this tile too?
          const int lc = tile_row_of(c);
          if (lc >= 0) {
            if (lc < lr) continue;                  // already emitted from the other side
            const int* cb = &pat.colind[pat.rowptr[c]];
            const int* ce = &pat.colind[pat.rowptr[c + 1]];
            int pos = 0;
            if (ce - cb <= 64) { for (const int* k = cb; k < ce; k++) pos += *k < r; }      // short row: branch-free count
            else pos = (int)(std::lower_bound(cb, ce, r) - cb);
            paired = 1; d1 = code(lc, pos);
          }
        }
        jitem[j] = (int)items.size();
        items.push_back(Item{code(lr, j), d1, paired, (uint32_t)(base + (size_t)j * cap), 0u});
      }
      // contributions in ascending (element, local column) order = ascending slot id: the fixed summation order of every entry
      for (long long a = pat.adj_ptr[r]; a < pat.adj_ptr[r + 1]; a++) {
        const int el = pat.adj_elem[a], pl = pat.adj_loc[a];
        const int le = (int)(std::lower_bound(te.begin(), te.end(), el) - te.begin());
        const uint32_t* sn = &pat.slot_nnz[((size_t)el * d + pl) * d];
        for (int q = 0; q < d; q++) {
          const int it = jitem[(size_t)(sn[q] - (uint32_t)rs)];
          if (it < 0) continue;
          Item& I = items[it];
          pool[I.off + I.cnt++] = (uint16_t)(sym ? symidx[pl * d + q] * nel + le : le * dd + pl * d + q);
          nsrc++;
        }
      }
    }
    if (nsrc > 65535 * 4 || nnz_t > 65535) { P.err = "tile too large"; return; }
    // classes of equal (source count, paired), ascending; tile order kept inside a class (coalesced stores)
    std::vector<int> ord(items.size());
    auto key = [&](int a) { return (int)items[a].cnt | items[a].paired << 16; };
    {   // stable counting sort by (paired, source count): a handful of distinct keys for thousands of items
      uint32_t maxc = 0;
      for (const Item& I : items) maxc = std::max(maxc, I.cnt);
      std::vector<int> start(2 * ((size_t)maxc + 1) + 1, 0);
      for (const Item& I : items) start[(size_t)I.paired * (maxc + 1) + I.cnt + 1]++;
      for (size_t b = 1; b < start.size(); b++) start[b] += start[b - 1];
      for (size_t i = 0; i < items.size(); i++) ord[start[(size_t)items[i].paired * (maxc + 1) + items[i].cnt]++] = (int)i;
    }
    std::vector<int> cls;            // {count | paired << 16, items, src offset, dst offset} per class
    std::vector<uint16_t> src, dst16;
    std::vector<uint32_t> dst32;
    auto push_dst = [&](uint32_t v) { if (ent32) dst32.push_back(v); else dst16.push_back((uint16_t)v); };
    for (size_t a = 0; a < ord.size();) {
      size_t b = a;
      const int kk = key(ord[a]), cnt = kk & 0xffff;
      while (b < ord.size() && key(ord[b]) == kk) b++;
      cls.push_back(kk); cls.push_back((int)(b - a)); cls.push_back((int)src.size()); cls.push_back((int)(ent32 ? dst32.size() : dst16.size()));
      for (int k = 0; k < cnt; k++) for (size_t i = a; i < b; i++) src.push_back(pool[items[ord[i]].off + k]);
      for (size_t i = a; i < b; i++) push_dst(items[ord[i]].d0);
      if (kk >> 16) for (size_t i = a; i < b; i++) push_dst(items[ord[i]].d1);
      a = b;
    }
    int hdr[8] = {nrows, nel, (int)tvert.size(), (int)(ent32 ? dst32.size() : dst16.size()), (int)src.size(), (int)cls.size() / 4, ent32, (int)nnz_t};
    const size_t at0 = P.blob.size();
    BlobWriter w(P.blob);
    w.section(hdr, 8); w.section(te);
    const size_t at1 = P.blob.size();
    w.section(rstart);
    if (!sym) w.section(rlen);
    w.section(tv); w.section(xy); w.section(cls);
    if (ent32) w.section(dst32); else w.section(dst16);
    w.section(src);
    P.sizes.push_back((long long)(at1 - at0)); P.sizes.push_back((long long)(P.blob.size() - at1));
    P.max_rows = std::max(P.max_rows, nrows); P.max_elems = std::max(P.max_elems, nel); P.max_nnz = std::max(P.max_nnz, (int)nnz_t);
    P.max_src = std::max(P.max_src, (int)src.size()); P.max_verts = std::max(P.max_verts, (int)tvert.size());
    P.tot_a += nel;
  });
  pt.lap("fwd tiles: per-tile blobs");
  if (!err.empty()) return err;
  max_rows = max_elems = max_nnz = max_src = max_verts = 0;
  long long tot = 0;
  for (auto& P : parts) {
    max_rows = std::max(max_rows, P.max_rows); max_elems = std::max(max_elems, P.max_elems); max_nnz = std::max(max_nnz, P.max_nnz);
    max_src = std::max(max_src, P.max_src); max_verts = std::max(max_verts, P.max_verts);
    tot += P.tot_a;
  }
  merge_parts(parts, blob_ptr, blob, max_head, max_body, nthreads);
  std::vector<int>().swap(morton_cache);
  pt.lap("fwd tiles: merge");
  elem_redundancy = m.ne > 0 ? (double)tot / m.ne : 0;
  return "";
}

// ------------------------------------------------------------------------------------------------
std::string AdjTiles::build(const HostMesh& m, const ScalarPattern& pat, int EPT, int max_tile_nnz, int nthreads) {
  elems_per_tile = EPT;
  const int d = m.d, dd = d * d, nvl = m.dim + 1;
  for (int r = 0; r < pat.n; r++) if (pat.rowptr[r + 1] - pat.rowptr[r] > 255) return "row longer than 255 entries";
  PhaseTimer pt;
  if ((long long)morton_cache.size() != m.ne) {
    Morton mc(m);
    morton_order(m.ne, nthreads, [&](long long e) {
      double c[3] = {0, 0, 0};
      for (int k = 0; k < nvl; k++) for (int a = 0; a < m.dim; a++) c[a] += m.coords[(size_t)m.verts[(size_t)e * nvl + k] * m.dim + a];
      for (int a = 0; a < m.dim; a++) c[a] /= nvl;
      return mc.code(c);
    }, morton_cache);
  }
  std::vector<int> order(morton_cache);
  ntiles = (m.ne + EPT - 1) / EPT;
  std::vector<int> elem_ptr(ntiles + 1);
  for (int t = 0; t <= ntiles; t++) elem_ptr[t] = (int)std::min<long long>((long long)t * EPT, m.ne);
  std::vector<int>& elems = order;
  parallel_for(ntiles, nthreads, [&](long long b, long long e, int) {
    for (long long t = b; t < e; t++) std::sort(elems.begin() + elem_ptr[t], elems.begin() + elem_ptr[t + 1]);
  });
  pt.lap("adj tiles: morton order, tile sort");
  std::vector<PartOut> parts;
  std::string err = run_parts(ntiles, nthreads, parts, too_big, [&](int t, PartOut& P) {
    std::vector<int> tr, te(elems.begin() + elem_ptr[t], elems.begin() + elem_ptr[t + 1]), tvert;
    std::vector<uint16_t> tv, roff, lrow16, td;
    std::vector<uint8_t> lrow8;
    std::vector<uint32_t> gpk, delta;
    std::vector<double> xy;
    std::vector<uint32_t> rstart;
    const int nel = (int)te.size();
    for (int e : te) {          // as in the forward tiles: independent misses, requested together
      __builtin_prefetch(&m.conn[(size_t)e * d]); __builtin_prefetch(&m.verts[(size_t)e * nvl]);
      for (int b = 0; b < dd * 4; b += 64) __builtin_prefetch(reinterpret_cast<const char*>(&pat.slot_nnz[(size_t)e * dd]) + b);
    }
    for (int e : te) { const int* ce = &m.conn[(size_t)e * d]; tr.insert(tr.end(), ce, ce + d); }
    std::sort(tr.begin(), tr.end());
    tr.erase(std::unique(tr.begin(), tr.end()), tr.end());
    const int nrows = (int)tr.size();
    for (int r : tr) __builtin_prefetch(&pat.rowptr[r]);
    const int lrow_wide = nrows > 256;
    long long acc = 0;
    roff.push_back(0);
    for (int lr = 0; lr < nrows; lr++) {
      const long long rs = pat.rowptr[tr[lr]], len = pat.rowptr[tr[lr] + 1] - rs;
      rstart.push_back((uint32_t)rs);
      delta.push_back((uint32_t)rs - (uint32_t)acc);      // global offset of staged entry i of this row = i + delta
      acc += len;
      if (acc > max_tile_nnz || acc > 65535 || nrows > 65535) { P.err = "tile too large"; return; }
      for (long long j = 0; j < len; j++) { if (lrow_wide) lrow16.push_back((uint16_t)lr); else lrow8.push_back((uint8_t)lr); }
      roff.push_back((uint16_t)acc);
    }
    tile_vertices(m, te, tvert, tv, xy);
    const int W = (d + 3) / 4;                    // 32-bit words holding the d 8-bit positions of one local row
    td.resize((size_t)d * nel); gpk.assign((size_t)d * W * nel, 0);
    for (int le = 0; le < nel; le++) {
      const int e = te[le];
      const int* ce = &m.conn[(size_t)e * d];
      for (int p = 0; p < d; p++) {
        td[(size_t)p * nel + le] = (uint16_t)(std::lower_bound(tr.begin(), tr.end(), ce[p]) - tr.begin());
        for (int q = 0; q < d; q++) {
          const uint32_t pos = (uint32_t)((long long)pat.slot_nnz[((size_t)e * d + p) * d + q] - pat.rowptr[ce[p]]);
          gpk[(size_t)(p * W + q / 4) * nel + le] |= pos << (8 * (q % 4));
        }
      }
    }
    // P1: the staged rows are the tile vertices in the same order, so td == tv and is not stored
    const int has_td = !(d == nvl && td == tv);
    int hdr[8] = {nrows, nel, (int)tvert.size(), (int)acc, lrow_wide | has_td << 1, 0, 0, 0};
    const size_t at0 = P.blob.size();
    BlobWriter w(P.blob);
    w.section(hdr, 8); w.section(delta); w.section(roff);
    if (lrow_wide) w.section(lrow16); else w.section(lrow8);
    const size_t at1 = P.blob.size();
    w.section(te); w.section(tv); w.section(xy);
    if (has_td) w.section(td);
    w.section(gpk);
    P.sizes.push_back((long long)(at1 - at0)); P.sizes.push_back((long long)(P.blob.size() - at1));
    P.max_rows = std::max(P.max_rows, nrows); P.max_elems = std::max(P.max_elems, nel); P.max_nnz = std::max(P.max_nnz, (int)acc);
    P.max_verts = std::max(P.max_verts, (int)tvert.size());
    P.tot_a += nrows;
  });
  pt.lap("adj tiles: per-tile blobs");
  if (!err.empty()) return err;
  max_rows = max_elems = max_nnz = max_verts = 0;
  long long tot = 0;
  for (auto& P : parts) {
    max_rows = std::max(max_rows, P.max_rows); max_elems = std::max(max_elems, P.max_elems); max_nnz = std::max(max_nnz, P.max_nnz);
    max_verts = std::max(max_verts, P.max_verts);
    tot += P.tot_a;
  }
  merge_parts(parts, blob_ptr, blob, max_head, max_body, nthreads);
  std::vector<int>().swap(morton_cache);
  pt.lap("adj tiles: merge");
  row_redundancy = pat.n > 0 ? (double)tot / pat.n : 0;
  return "";
}

}  // namespace adfem
