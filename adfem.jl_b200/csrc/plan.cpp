#include "plan.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
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
  for (size_t i = 0; i < tvert.size(); i++)
    for (int c = 0; c < m.dim; c++) xy[i * m.dim + c] = m.coords[(size_t)tvert[i] * m.dim + c];
}

struct PartOut { std::vector<uint8_t> blob; std::vector<long long> sizes; std::string err; int max_rows = 0, max_elems = 0, max_nnz = 0, max_src = 0, max_verts = 0; long long tot_a = 0; };

// per_tile appends the tile's head and body to P.blob and pushes their two sizes to P.sizes
template <class F> std::string run_parts(int ntiles, int nthreads, std::vector<PartOut>& parts, F per_tile) {
  int nparts = std::max(1, std::min(nthreads, ntiles));
  parts.assign(nparts, PartOut());
  std::vector<long long> pcut(nparts + 1);
  for (int i = 0; i <= nparts; i++) pcut[i] = (long long)ntiles * i / nparts;
  std::vector<std::thread> th;
  for (int pi = 0; pi < nparts; pi++) th.emplace_back([&, pi] {
    PartOut& P = parts[pi];
    for (long long t = pcut[pi]; t < pcut[pi + 1] && P.err.empty(); t++) per_tile((int)t, P);
  });
  for (auto& t : th) t.join();
  for (auto& P : parts) if (!P.err.empty()) return P.err;
  return "";
}

// concatenate the per-thread outputs: blob, blob_ptr (2 offsets per tile + end), largest head / body
void merge_parts(std::vector<PartOut>& parts, std::vector<long long>& blob_ptr, std::vector<uint8_t>& blob, size_t& max_head, size_t& max_body) {
  blob_ptr.assign(1, 0); blob.clear(); max_head = max_body = 0;
  size_t total = 0;
  for (auto& P : parts) total += P.blob.size();
  blob.reserve(total);
  for (auto& P : parts) {
    for (size_t i = 0; i < P.sizes.size(); i++) {
      blob_ptr.push_back(blob_ptr.back() + P.sizes[i]);
      size_t& mx = (i & 1) ? max_body : max_head;
      mx = std::max(mx, (size_t)P.sizes[i]);
    }
    blob.insert(blob.end(), P.blob.begin(), P.blob.end());
    std::vector<uint8_t>().swap(P.blob);
  }
}
}  // namespace

std::string FwdTiles::build(const HostMesh& m, const ScalarPattern& pat, int R, int max_tile_elems, int sym_, int nthreads) {
  rows_per_tile = R; sym = sym_;
  const int d = m.d, dd = d * d, n = pat.n;
  const int nslot = sym ? d * (d + 1) / 2 : dd;
  Morton mc(m);
  std::vector<int> order;
  morton_order(n, nthreads, [&](long long i) { double x[3]; m.dof_position((int)i, x); return mc.code(x); }, order);
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

  std::vector<PartOut> parts;
  std::string err = run_parts(ntiles, nthreads, parts, [&](int t, PartOut& P) {
    std::vector<ColSlot> ps;
    std::vector<int> te, tvert;
    std::vector<uint16_t> tv, rlen;
    std::vector<double> xy;
    std::vector<uint32_t> rstart;
    struct Item { uint32_t d0, d1; int paired; std::vector<uint16_t> src; };
    std::vector<Item> items;
    const int* trow = rows.data() + row_ptr[t];
    const int nrows = row_ptr[t + 1] - row_ptr[t];
    for (int i = 0; i < nrows; i++) { int r = trow[i]; for (long long a = pat.adj_ptr[r]; a < pat.adj_ptr[r + 1]; a++) te.push_back(pat.adj_elem[a]); }
    std::sort(te.begin(), te.end());
    te.erase(std::unique(te.begin(), te.end()), te.end());
    const int nel = (int)te.size();
    if (nel > max_tile_elems || (long long)nel * (sym ? nslot : dd) > 65535) { P.err = "tile too large"; return; }
    tile_vertices(m, te, tvert, tv, xy);
    if (tvert.size() > 65535) { P.err = "tile too large"; return; }
    auto code = [&](int lr, int j) { return ent32 ? ((uint32_t)lr | (uint32_t)j << 16) : ((uint32_t)lr | (uint32_t)j << 8); };
    size_t nsrc = 0, nnz_t = 0;
    for (int lr = 0; lr < nrows; lr++) {
      const int r = trow[lr];
      rstart.push_back((uint32_t)pat.rowptr[r]);
      rlen.push_back((uint16_t)(pat.rowptr[r + 1] - pat.rowptr[r]));
      nnz_t += rlen.back();
      row_pairs(m, pat, r, ps);
      int j = -1;
      bool skip = false;
      for (size_t k = 0; k < ps.size(); k++) {
        if (k == 0 || ps[k].first != ps[k - 1].first) {
          j++;
          const int c = ps[k].first;
          skip = false;
          int paired = 0; uint32_t d1 = 0;
          if (sym && c != r) {                       // is the mirrored entry (c, r) produced by this tile too?
            const int* it = std::lower_bound(trow, trow + nrows, c);
            if (it != trow + nrows && *it == c) {
              const int lc = (int)(it - trow);
              if (lc < lr) skip = true;               // already emitted from the other side
              else {
                const int* cb = &pat.colind[pat.rowptr[c]];
                const int* ce = &pat.colind[pat.rowptr[c + 1]];
                paired = 1; d1 = code(lc, (int)(std::lower_bound(cb, ce, r) - cb));
              }
            }
          }
          if (!skip) items.push_back(Item{code(lr, j), d1, paired, {}});
        }
        if (skip) continue;
        int el = (int)(ps[k].second / dd), pq = (int)(ps[k].second % dd);
        int le = (int)(std::lower_bound(te.begin(), te.end(), el) - te.begin());
        items.back().src.push_back((uint16_t)(sym ? symidx[pq] * nel + le : le * dd + pq));
        nsrc++;
      }
    }
    if (nsrc > 65535 * 4 || nnz_t > 65535) { P.err = "tile too large"; return; }
    // classes of equal (source count, paired), ascending; tile order kept inside a class (coalesced stores)
    std::vector<int> ord(items.size());
    for (size_t i = 0; i < ord.size(); i++) ord[i] = (int)i;
    auto key = [&](int a) { return (int)items[a].src.size() | items[a].paired << 16; };
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return key(a) < key(b); });
    std::vector<int> cls;            // {count | paired << 16, items, src offset, dst offset} per class
    std::vector<uint16_t> src, dst16;
    std::vector<uint32_t> dst32;
    auto push_dst = [&](uint32_t v) { if (ent32) dst32.push_back(v); else dst16.push_back((uint16_t)v); };
    for (size_t a = 0; a < ord.size();) {
      size_t b = a;
      const int kk = key(ord[a]), cnt = kk & 0xffff;
      while (b < ord.size() && key(ord[b]) == kk) b++;
      cls.push_back(kk); cls.push_back((int)(b - a)); cls.push_back((int)src.size()); cls.push_back((int)(ent32 ? dst32.size() : dst16.size()));
      for (int k = 0; k < cnt; k++) for (size_t i = a; i < b; i++) src.push_back(items[ord[i]].src[k]);
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
  if (!err.empty()) return err;
  max_rows = max_elems = max_nnz = max_src = max_verts = 0;
  long long tot = 0;
  for (auto& P : parts) {
    max_rows = std::max(max_rows, P.max_rows); max_elems = std::max(max_elems, P.max_elems); max_nnz = std::max(max_nnz, P.max_nnz);
    max_src = std::max(max_src, P.max_src); max_verts = std::max(max_verts, P.max_verts);
    tot += P.tot_a;
  }
  merge_parts(parts, blob_ptr, blob, max_head, max_body);
  elem_redundancy = m.ne > 0 ? (double)tot / m.ne : 0;
  return "";
}

// ------------------------------------------------------------------------------------------------
std::string AdjTiles::build(const HostMesh& m, const ScalarPattern& pat, int EPT, int max_tile_nnz, int nthreads) {
  elems_per_tile = EPT;
  const int d = m.d, dd = d * d, nvl = m.dim + 1;
  for (int r = 0; r < pat.n; r++) if (pat.rowptr[r + 1] - pat.rowptr[r] > 255) return "row longer than 255 entries";
  Morton mc(m);
  std::vector<int> order;
  morton_order(m.ne, nthreads, [&](long long e) {
    double c[3] = {0, 0, 0};
    for (int k = 0; k < nvl; k++) for (int a = 0; a < m.dim; a++) c[a] += m.coords[(size_t)m.verts[(size_t)e * nvl + k] * m.dim + a];
    for (int a = 0; a < m.dim; a++) c[a] /= nvl;
    return mc.code(c);
  }, order);
  ntiles = (m.ne + EPT - 1) / EPT;
  std::vector<int> elem_ptr(ntiles + 1);
  for (int t = 0; t <= ntiles; t++) elem_ptr[t] = (int)std::min<long long>((long long)t * EPT, m.ne);
  std::vector<int>& elems = order;
  parallel_for(ntiles, nthreads, [&](long long b, long long e, int) {
    for (long long t = b; t < e; t++) std::sort(elems.begin() + elem_ptr[t], elems.begin() + elem_ptr[t + 1]);
  });
  std::vector<PartOut> parts;
  std::string err = run_parts(ntiles, nthreads, parts, [&](int t, PartOut& P) {
    std::vector<int> tr, te(elems.begin() + elem_ptr[t], elems.begin() + elem_ptr[t + 1]), tvert;
    std::vector<uint16_t> tv, roff, lrow16, td;
    std::vector<uint8_t> lrow8;
    std::vector<uint32_t> gpk, delta;
    std::vector<double> xy;
    std::vector<uint32_t> rstart;
    const int nel = (int)te.size();
    for (int e : te) { const int* ce = &m.conn[(size_t)e * d]; tr.insert(tr.end(), ce, ce + d); }
    std::sort(tr.begin(), tr.end());
    tr.erase(std::unique(tr.begin(), tr.end()), tr.end());
    const int nrows = (int)tr.size();
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
  if (!err.empty()) return err;
  max_rows = max_elems = max_nnz = max_verts = 0;
  long long tot = 0;
  for (auto& P : parts) {
    max_rows = std::max(max_rows, P.max_rows); max_elems = std::max(max_elems, P.max_elems); max_nnz = std::max(max_nnz, P.max_nnz);
    max_verts = std::max(max_verts, P.max_verts);
    tot += P.tot_a;
  }
  merge_parts(parts, blob_ptr, blob, max_head, max_body);
  row_redundancy = pat.n > 0 ? (double)tot / pat.n : 0;
  return "";
}

}  // namespace adfem
