#include "plan.h"

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <thread>
#include <utility>

namespace adfem {

int default_threads() {
  if (const char* s = getenv("ADFEM_HOST_THREADS")) { int v = atoi(s); if (v > 0) return v; }
  unsigned n = std::thread::hardware_concurrency();
  return n == 0 ? 4 : (int)std::min(n, 64u);
}

namespace {

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

typedef std::pair<int, uint64_t> ColSlot;   // (column, global slot id (e*d+p)*d+q)

// all local entries that land in row r, sorted by (col, slot): the fixed summation order of every nnz
inline void row_pairs(const HostMesh& m, const ScalarPattern& pat, int r, std::vector<ColSlot>& out) {
  out.clear();
  const int d = m.d;
  for (long long a = pat.adj_ptr[r]; a < pat.adj_ptr[r + 1]; a++) {
    int e = pat.adj_elem[a], p = pat.adj_loc[a];
    const int* ce = &m.conn[(size_t)e * d];
    uint64_t base = ((uint64_t)e * d + p) * d;
    for (int q = 0; q < d; q++) out.emplace_back(ce[q], base + q);
  }
  std::sort(out.begin(), out.end());
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

// ids sorted by Morton code of their position
void morton_order(long long n, int nthreads, const std::function<uint64_t(long long)>& code, std::vector<int>& order) {
  std::vector<std::pair<uint64_t, int>> keyed(n);
  parallel_for(n, nthreads, [&](long long b, long long e, int) { for (long long i = b; i < e; i++) keyed[i] = {code(i), (int)i}; });
  // sort blocks in parallel, then merge pairwise
  int nb = 1;
  while (nb < nthreads && n / (nb * 2) > 65536) nb *= 2;
  std::vector<long long> cut(nb + 1);
  for (int i = 0; i <= nb; i++) cut[i] = n * i / nb;
  {
    std::vector<std::thread> th;
    for (int i = 0; i < nb; i++) th.emplace_back([&, i] { std::sort(keyed.begin() + cut[i], keyed.begin() + cut[i + 1]); });
    for (auto& t : th) t.join();
  }
  for (int w = 1; w < nb; w *= 2) {
    std::vector<std::thread> th;
    for (int i = 0; i + w < nb; i += 2 * w)
      th.emplace_back([&, i, w] { std::inplace_merge(keyed.begin() + cut[i], keyed.begin() + cut[i + w], keyed.begin() + cut[std::min(nb, i + 2 * w)]); });
    for (auto& t : th) t.join();
  }
  order.resize(n);
  for (long long i = 0; i < n; i++) order[i] = keyed[i].second;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
std::string ScalarPattern::build(const HostMesh& m, int nthreads) {
  n = m.ndof;
  const int d = m.d;
  const long long nslot_rows = (long long)m.ne * d;
  if (nslot_rows > 2147483647LL) return "ne*elem_ndof exceeds 32-bit";
  // dof -> (element, local) adjacency by counting sort (element order preserved inside a row)
  adj_ptr.assign((size_t)n + 1, 0);
  for (long long i = 0; i < nslot_rows; i++) adj_ptr[m.conn[i] + 1]++;
  for (int r = 0; r < n; r++) adj_ptr[r + 1] += adj_ptr[r];
  adj_elem.resize(nslot_rows); adj_loc.resize(nslot_rows);
  {
    std::vector<long long> cur(adj_ptr.begin(), adj_ptr.end() - 1);
    for (int e = 0; e < m.ne; e++)
      for (int p = 0; p < d; p++) { long long at = cur[m.conn[(size_t)e * d + p]]++; adj_elem[at] = e; adj_loc[at] = (uint8_t)p; }
  }
  // pass 1: row lengths
  std::vector<int> rowlen(n);
  parallel_for(n, nthreads, [&](long long b, long long e, int) {
    std::vector<ColSlot> ps;
    for (long long r = b; r < e; r++) {
      row_pairs(m, *this, (int)r, ps);
      int cnt = 0;
      for (size_t i = 0; i < ps.size(); i++) if (i == 0 || ps[i].first != ps[i - 1].first) cnt++;
      rowlen[r] = cnt;
    }
  });
  rowptr.assign((size_t)n + 1, 0);
  for (int r = 0; r < n; r++) rowptr[r + 1] = rowptr[r] + rowlen[r];
  nnz = rowptr[n];
  if (nnz > 4294967295LL) return "scalar nnz exceeds 32-bit slot map";
  colind.resize(nnz);
  slot_nnz.resize((size_t)m.ne * d * d);
  // pass 2: columns + slot map
  parallel_for(n, nthreads, [&](long long b, long long e, int) {
    std::vector<ColSlot> ps;
    for (long long r = b; r < e; r++) {
      row_pairs(m, *this, (int)r, ps);
      long long at = rowptr[r] - 1;
      for (size_t i = 0; i < ps.size(); i++) {
        if (i == 0 || ps[i].first != ps[i - 1].first) colind[++at] = ps[i].first;
        slot_nnz[ps[i].second] = (uint32_t)at;
      }
    }
  });
  return "";
}

// ------------------------------------------------------------------------------------------------
std::string TilePlan::build(const HostMesh& m, const ScalarPattern& pat, int R, int max_tile_elems, int nthreads) {
  rows_per_tile = R;
  const int d = m.d, dd = d * d, n = pat.n;
  Morton mc(m);
  std::vector<int> order;
  morton_order(n, nthreads, [&](long long i) { double x[3]; m.dof_position((int)i, x); return mc.code(x); }, order);
  ntiles = (n + R - 1) / R;
  row_ptr.resize(ntiles + 1);
  for (int t = 0; t <= ntiles; t++) row_ptr[t] = (int)std::min<long long>((long long)t * R, n);
  rows = order;
  parallel_for(ntiles, nthreads, [&](long long b, long long e, int) {
    for (long long t = b; t < e; t++) std::sort(rows.begin() + row_ptr[t], rows.begin() + row_ptr[t + 1]);
  });

  struct Part { std::vector<int> elems, elem_cnt; std::vector<uint16_t> soff, src; std::vector<long long> soff_cnt, src_cnt; std::string err; };
  int nparts = std::max(1, std::min(nthreads, ntiles));
  std::vector<Part> parts(nparts);
  std::vector<long long> pcut(nparts + 1);
  for (int i = 0; i <= nparts; i++) pcut[i] = (long long)ntiles * i / nparts;
  {
    std::vector<std::thread> th;
    for (int pi = 0; pi < nparts; pi++) th.emplace_back([&, pi] {
      Part& P = parts[pi];
      std::vector<ColSlot> ps;
      std::vector<int> te;
      for (long long t = pcut[pi]; t < pcut[pi + 1]; t++) {
        te.clear();
        for (int i = row_ptr[t]; i < row_ptr[t + 1]; i++) { int r = rows[i]; for (long long a = pat.adj_ptr[r]; a < pat.adj_ptr[r + 1]; a++) te.push_back(pat.adj_elem[a]); }
        std::sort(te.begin(), te.end());
        te.erase(std::unique(te.begin(), te.end()), te.end());
        if ((long long)te.size() > max_tile_elems || (long long)te.size() * dd > 65535) { P.err = "tile too large"; return; }
        P.elem_cnt.push_back((int)te.size());
        P.elems.insert(P.elems.end(), te.begin(), te.end());
        size_t s0 = P.src.size(), o0 = P.soff.size();
        for (int i = row_ptr[t]; i < row_ptr[t + 1]; i++) {
          row_pairs(m, pat, rows[i], ps);
          for (size_t k = 0; k < ps.size(); k++) {
            if (k == 0 || ps[k].first != ps[k - 1].first) P.soff.push_back((uint16_t)(P.src.size() - s0));
            int el = (int)(ps[k].second / dd), pq = (int)(ps[k].second % dd);
            int le = (int)(std::lower_bound(te.begin(), te.end(), el) - te.begin());
            P.src.push_back((uint16_t)(le * dd + pq));
          }
        }
        if (P.src.size() - s0 > 65535) { P.err = "tile too large"; return; }
        P.soff.push_back((uint16_t)(P.src.size() - s0));
        P.soff_cnt.push_back((long long)(P.soff.size() - o0));
        P.src_cnt.push_back((long long)(P.src.size() - s0));
      }
    });
    for (auto& t : th) t.join();
  }
  for (auto& P : parts) if (!P.err.empty()) return P.err;
  elem_ptr.assign(1, 0); soff_ptr.assign(1, 0); src_ptr.assign(1, 0);
  elems.clear(); src_off.clear(); src.clear();
  max_rows = max_elems = max_nnz = max_src = 0;
  for (auto& P : parts) {
    for (size_t i = 0; i < P.elem_cnt.size(); i++) {
      elem_ptr.push_back(elem_ptr.back() + P.elem_cnt[i]);
      soff_ptr.push_back(soff_ptr.back() + P.soff_cnt[i]);
      src_ptr.push_back(src_ptr.back() + P.src_cnt[i]);
      max_elems = std::max(max_elems, P.elem_cnt[i]);
      max_nnz = std::max<int>(max_nnz, (int)P.soff_cnt[i] - 1);
      max_src = std::max<int>(max_src, (int)P.src_cnt[i]);
    }
    elems.insert(elems.end(), P.elems.begin(), P.elems.end());
    src_off.insert(src_off.end(), P.soff.begin(), P.soff.end());
    src.insert(src.end(), P.src.begin(), P.src.end());
    P = Part();
  }
  for (int t = 0; t < ntiles; t++) max_rows = std::max(max_rows, row_ptr[t + 1] - row_ptr[t]);
  elem_redundancy = m.ne > 0 ? (double)elems.size() / m.ne : 0;
  return "";
}

// ------------------------------------------------------------------------------------------------
std::string AdjTilePlan::build(const HostMesh& m, const ScalarPattern& pat, int EPT, int max_tile_nnz, int nthreads) {
  elems_per_tile = EPT;
  const int d = m.d, dd = d * d, nvl = m.dim + 1;
  Morton mc(m);
  std::vector<int> order;
  morton_order(m.ne, nthreads, [&](long long e) {
    double c[3] = {0, 0, 0};
    for (int k = 0; k < nvl; k++) for (int a = 0; a < m.dim; a++) c[a] += m.coords[(size_t)m.verts[(size_t)e * nvl + k] * m.dim + a];
    for (int a = 0; a < m.dim; a++) c[a] /= nvl;
    return mc.code(c);
  }, order);
  ntiles = (m.ne + EPT - 1) / EPT;
  elem_ptr.resize(ntiles + 1);
  for (int t = 0; t <= ntiles; t++) elem_ptr[t] = (int)std::min<long long>((long long)t * EPT, m.ne);
  elems = order;
  parallel_for(ntiles, nthreads, [&](long long b, long long e, int) {
    for (long long t = b; t < e; t++) std::sort(elems.begin() + elem_ptr[t], elems.begin() + elem_ptr[t + 1]);
  });
  struct Part { std::vector<int> rows, row_cnt; std::vector<uint16_t> gidx; std::vector<int> nnz_cnt; std::string err; };
  int nparts = std::max(1, std::min(nthreads, ntiles));
  std::vector<Part> parts(nparts);
  std::vector<long long> pcut(nparts + 1);
  for (int i = 0; i <= nparts; i++) pcut[i] = (long long)ntiles * i / nparts;
  {
    std::vector<std::thread> th;
    for (int pi = 0; pi < nparts; pi++) th.emplace_back([&, pi] {
      Part& P = parts[pi];
      std::vector<int> tr;
      std::vector<long long> roff;
      for (long long t = pcut[pi]; t < pcut[pi + 1]; t++) {
        tr.clear();
        for (int i = elem_ptr[t]; i < elem_ptr[t + 1]; i++) { const int* ce = &m.conn[(size_t)elems[i] * d]; tr.insert(tr.end(), ce, ce + d); }
        std::sort(tr.begin(), tr.end());
        tr.erase(std::unique(tr.begin(), tr.end()), tr.end());
        roff.assign(tr.size() + 1, 0);
        for (size_t i = 0; i < tr.size(); i++) roff[i + 1] = roff[i] + (pat.rowptr[tr[i] + 1] - pat.rowptr[tr[i]]);
        if (roff.back() > max_tile_nnz || roff.back() > 65535) { P.err = "tile too large"; return; }
        P.row_cnt.push_back((int)tr.size());
        P.nnz_cnt.push_back((int)roff.back());
        P.rows.insert(P.rows.end(), tr.begin(), tr.end());
        for (int i = elem_ptr[t]; i < elem_ptr[t + 1]; i++) {
          int e = elems[i];
          const int* ce = &m.conn[(size_t)e * d];
          for (int p = 0; p < d; p++) {
            int lr = (int)(std::lower_bound(tr.begin(), tr.end(), ce[p]) - tr.begin());
            for (int q = 0; q < d; q++) {
              long long pos = pat.slot_nnz[((size_t)e * d + p) * d + q] - pat.rowptr[ce[p]];
              P.gidx.push_back((uint16_t)(roff[lr] + pos));
            }
          }
        }
      }
    });
    for (auto& t : th) t.join();
  }
  for (auto& P : parts) if (!P.err.empty()) return P.err;
  row_ptr.assign(1, 0); gidx_ptr.assign(1, 0);
  rows.clear(); gidx.clear();
  max_rows = max_elems = max_nnz = 0;
  int t = 0;
  for (auto& P : parts) {
    for (size_t i = 0; i < P.row_cnt.size(); i++, t++) {
      row_ptr.push_back(row_ptr.back() + P.row_cnt[i]);
      gidx_ptr.push_back(gidx_ptr.back() + (long long)(elem_ptr[t + 1] - elem_ptr[t]) * dd);
      max_rows = std::max(max_rows, P.row_cnt[i]);
      max_nnz = std::max(max_nnz, P.nnz_cnt[i]);
      max_elems = std::max(max_elems, elem_ptr[t + 1] - elem_ptr[t]);
    }
    rows.insert(rows.end(), P.rows.begin(), P.rows.end());
    gidx.insert(gidx.end(), P.gidx.begin(), P.gidx.end());
    P = Part();
  }
  row_redundancy = pat.n > 0 ? (double)rows.size() / pat.n : 0;
  return "";
}

}  // namespace adfem
