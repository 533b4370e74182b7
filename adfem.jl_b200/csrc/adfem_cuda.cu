// libadfem_cuda.so — mesh handle, symbolic phase, kernel dispatch and the C ABI of include/adfem_cuda.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/adfem_cuda.h"
#include "gauss_ops.h"
#include "host_mesh.h"
#include "internal.h"
#include "kernels.cuh"
#include "plan.h"
#include "tri_grid.cuh"
#include "grid_elast.cuh"
#include "tet_grid.cuh"
#include "tet_node.cuh"
#include "row_gather.cuh"
#include "tet_scalar.cuh"

using namespace adfem;

namespace adfem {
thread_local std::string g_err;
int fail(const std::string& msg) { g_err = msg; return 1; }
}  // namespace adfem

namespace {

template <class T> cudaError_t upload(DevBuf<T>& b, const std::vector<T>& v) {
  cudaError_t e = b.alloc(v.size());
  if (e != cudaSuccess) return e;
  return v.empty() ? cudaSuccess : cudaMemcpy(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
}

struct FwdPlanDev {
  FwdTiles host;   // blob kept on the host only for ADFEM_HOST_ONLY handles (inspection)
  DevBuf<long long> blob_ptr;
  DevBuf<uint8_t> blob;
  DevTiles dev{};
  size_t bytes = 0;
};
struct AdjPlanDev {
  AdjTiles host;
  DevBuf<long long> blob_ptr;
  DevBuf<uint8_t> blob;
  DevTiles dev{};
  size_t bytes = 0;
};

}  // namespace

struct adfem_mesh {
  HostMesh hm;
  bool host_only = false;
  int device = 0;
  DevBuf<double> coords;
  DevBuf<int> verts, conn;
  DevMesh dm{};
  // symbolic
  bool has_pattern = false;
  ScalarPattern pat;
  DevBuf<long long> d_rowptr, d_adj_ptr;
  DevBuf<int> d_colind, d_adj_elem;
  DevBuf<uint8_t> d_adj_loc;
  DevBuf<uint32_t> d_slot_nnz;
  DevPattern dpat{};
  std::map<int, std::unique_ptr<FwdPlanDev>> fwd_plans;   // keyed by S = (nc*d)^2
  std::map<int, std::unique_ptr<AdjPlanDev>> adj_plans;   // keyed by nc
  // options
  int opt_rows_per_tile = 0, opt_elems_per_tile = 0, opt_adjoint_tiled = 1, opt_threads = 0;
  int opt_tile_overlap = 0;                 // scalar tile forward with one barrier per tile (double-buffered local matrices, kernels.cuh k_tile_fwd_ov): 1 = on, -1 = P2 only.
                                            // Off: the gain is not robust.  Measured P2 forward, one GPU: 16 M triangles 5.84 -> 5.15 ms (random numbering), 3.21 -> 2.98 ms
                                            // (Morton elements), 2 M triangles 0.354 -> 0.376 ms; element blocks under torchrun: 8 M per GPU 1.52 -> 2.02 ms, 2 M per GPU
                                            // 0.50 -> 0.63 ms.  P1 forward on 33.5 M triangles 0.730 -> 0.950 ms (smaller tiles, more halo).
  int opt_smem_budget_adj = 0;              // the same for the adjoint tile kernels only (0 = opt_smem_budget / per-operator default)
  int opt_smem_budget = 0;                  // dynamic shared memory per CTA (3 head + 2 body buffers + staging); 0 = per-operator default
  int opt_tile_threads = 0;                 // threads per CTA of the tile kernels; 0 = per-operator default
  int opt_pipeline = 1;                     // 1 = persistent CTAs (software pipeline across tiles), 0 = one CTA per tile
  int opt_grid_limit = 0;                   // > 0: cap the persistent grid (tests: few CTAs walk many tiles)
  int opt_coef_prefetch = 1;                // forward: register prefetch of the next tile's coefficients (P1 scalar operators)
  int opt_row_gather = 0;                   // scalar operators: one-thread-per-row forward (row_gather.cuh) instead of the row-tile kernel (off until measured)
  int rg_rows = 0;                          // rows per CTA of the row-gather forward: 128, 64 or 32, the largest whose CTAs all fit the staging; 0 = none does
  int rge_max_entries = 0;                  // elasticity variant: most CSR entries of any CTA's 64 rows (sizes its dynamic shared memory)
  int opt_coef_presum = 0;                  // P1 elasticity: reduce the g coefficient blocks of an element to one in a streaming pre-pass (and expand the
                                            // adjoint's per-element block afterwards), so the tile kernels move NS*NS instead of g*NS*NS doubles per element
  DevBuf<double> presum_buf;                // ne * NS*NS doubles of scratch for it
  bool adj_untileable = false;              // a CSR row has more than 255 entries: adjoint uses the direct gather kernel
  int num_sms = 0;
  int opt_area_csr = 0, opt_area_coo = 1;   // 2-D weight scale: 0 = det/2, 1 = Heron (reference formula)
  // structured triangulation Mesh(m, n, h) (tri_grid.cuh): detected from the arrays, no mesh-static index data is read
  bool grid_ok = false;
  bool grid_mapped = false;                 // structured connectivity, non-rectilinear node positions: scalar CSR kernels only (MAPPED)
  int grid_m = 0, grid_n = 0, opt_structured = 1, opt_grid_rows = 0, opt_grid_occupancy = 2;
  bool opt_grid_pattern = true;              // symbolic tables of the structured triangulation in closed form (ScalarPattern::build_tri_grid)
  int opt_grid_elast = 1;                   // P1 elasticity on Mesh(m,n,h) / Mesh3(n,n,l,h): index-free kernels of grid_elast.cuh / tet_grid.cuh (measured round 2: 0.63 / 0.24 of roofline against 0.46 / 0.11)
  int opt_tet_node = 1;                     // config 5 forward: 1 = x-fastest Gauss pre-sum + one thread per (node, component) (tet_node.cuh), 0 = one warp per node (tet_grid.cuh),
                                            // 2 = as 1 with the 32-tetrahedron incidence list split over two warps (measured: 5.55 -> 5.75 ms at 10.5 M tetrahedra, no gain)
  int opt_tet_adj_blocks = 3;               // resident CTAs per SM the tetrahedral adjoint kernel is compiled for (register cap 168 / 128 / 96): measured 3.77 / 3.87 / 4.75 ms at 10.5 M tetrahedra
  int opt_tet_chunks = 0;                   // z-chunks of the two-stream forward pipeline (0 or 1 = off: measured no gain)
  cudaStream_t tet_side_stream = nullptr;
  std::vector<cudaEvent_t> tet_events;
  bool tet_const_ready = false;
  DevBuf<double> tet_spacing;              // [1/hx | 1/hy | 1/hz | hx | hy | hz] per cube and axis (tet_node.cuh)
  int opt_tet_scalar = 0;                   // scalar P1 operators on Mesh3(n,n,l,h) through tet_scalar.cuh: measured slower than the tile kernels (0.358 vs 0.151 ms), kept opt-in
  DevBuf<double> grid_xs, grid_ys;
  // structured tetrahedral grid Mesh3(n, n, l, h) (tet_grid.cuh): detected from the arrays; used by the opt-in elasticity forward kernel
  bool tet_ok = false;
  int tet_n = 0, tet_l = 0;
  std::vector<double> tet_axes[3];
  TetGridTables tet_tab;
  DevBuf<double> tet_xs, tet_ys, tet_zs;
  DevBuf<TetGridTables> d_tet_tab;
  // scratch and streams for the host-buffer calls (H2D, kernels, D2H)
  cudaStream_t hs[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t hev[2] = {nullptr, nullptr};
  int opt_host_chunks = 16;
  ~adfem_mesh() {
    for (auto& s : hs) if (s) cudaStreamDestroy(s);
    for (auto& e : hev) if (e) cudaEventDestroy(e);
    for (auto e : tet_events) cudaEventDestroy(e);
    if (tet_side_stream) cudaStreamDestroy(tet_side_stream);
  }
  DevBuf<double> s_in, s_out;
};

namespace {

#define CU_TRY(call)                                                                                   \
  do {                                                                                                 \
    cudaError_t _e = (call);                                                                           \
    if (_e != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(_e));            \
  } while (0)

int need_device(const adfem_mesh* m) {
  if (!m) return fail("null mesh handle");
  if (m->host_only) return fail("mesh was created with ADFEM_HOST_ONLY: no device path (and there is no CPU fallback)");
  cudaError_t e = cudaSetDevice(m->device);
  if (e != cudaSuccess) return fail(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
  return 0;
}

// Is this the reference's structured triangulation (src/MFEM/MFEM.jl:134-146, version 1) on rectilinear coordinates?
// `mapped` = the connectivity is the generator's but the node positions are not rectilinear (a jittered / mapped grid): the scalar CSR
// kernels then read the positions from the coordinate array (tri_grid.cuh, MAPPED), every other structured kernel is not used.
bool detect_tri_grid(const HostMesh& h, int& m, int& n, std::vector<double>& xs, std::vector<double>& ys, bool& mapped) {
  mapped = false;
  if (h.dim != 2 || h.degree != 1 || h.g != 3 || h.ne < 2 || h.nv < 4) return false;
  const int* v = h.verts.data();
  if (v[0] != 0 || v[1] != 1 || v[2] < 2) return false;
  m = v[2] - 1;
  if (h.ne % (2 * m) != 0) return false;
  n = h.ne / (2 * m);
  if ((long long)(m + 1) * (n + 1) != h.nv) return false;
  for (int ci = 0; ci < n; ci++)
    for (int cj = 0; cj < m; cj++) {
      const int a = ci * (m + 1) + cj;
      const int* t = v + 6 * ((size_t)ci * m + cj);
      if (t[0] != a || t[1] != a + 1 || t[2] != a + m + 1 || t[3] != a + m + 1 || t[4] != a + 1 || t[5] != a + m + 2) return false;
    }
  xs.resize(m + 1); ys.resize(n + 1);
  for (int j = 0; j <= m; j++) xs[j] = h.coords[2 * (size_t)j];
  for (int i = 0; i <= n; i++) ys[i] = h.coords[2 * (size_t)i * (m + 1) + 1];
  for (int j = 0; j < m && !mapped; j++) if (!(xs[j] < xs[j + 1])) mapped = true;
  for (int i = 0; i < n && !mapped; i++) if (!(ys[i] < ys[i + 1])) mapped = true;
  for (int i = 0; i <= n && !mapped; i++)
    for (int j = 0; j <= m; j++) {
      const double* c = &h.coords[2 * ((size_t)i * (m + 1) + j)];
      if (c[0] != xs[j] || c[1] != ys[i]) { mapped = true; break; }       // bitwise: the rectilinear kernels recompute nothing, they index xs / ys
    }
  // the vertex order above is the one AFTER the orientation fix, so every triangle is positively oriented in either case
  return true;
}

// the closed-form row pointers of tri_grid.cuh must reproduce the symbolic pattern exactly
// Is this the reference's structured tetrahedral grid (src/MFEM3/MFEM.jl:124-185: 5 tetrahedra per cube, TE1 / TE2 by the parity of the cube) on
// rectilinear coordinates?  The element's vertex SETS are compared (MFEM's orientation fix may have swapped the first two vertices).
bool detect_tet_grid(const HostMesh& h, int& n, int& l, std::vector<double> ax[3]) {
  if (h.dim != 3 || h.degree != 1 || h.g != 4 || h.nv < 8 || h.ne < 5) return false;
  const double* c = h.coords.data();
  long long n1 = 0;
  for (long long t = 1; t < h.nv; t++) if (c[3 * t + 1] != c[1]) { n1 = t; break; }
  if (n1 < 2 || h.nv % (n1 * n1) != 0) return false;
  const long long l1 = h.nv / (n1 * n1);
  n = (int)n1 - 1; l = (int)l1 - 1;
  if (l < 1 || (long long)h.ne != 5LL * n * n * l) return false;
  ax[0].resize(n1); ax[1].resize(n1); ax[2].resize(l1);
  for (long long i = 0; i < n1; i++) { ax[0][i] = c[3 * i]; ax[1][i] = c[3 * (i * n1) + 1]; }
  for (long long k = 0; k < l1; k++) ax[2][k] = c[3 * (k * n1 * n1) + 2];
  for (int d = 0; d < 3; d++)
    for (size_t t = 0; t + 1 < ax[d].size(); t++) if (!(ax[d][t] < ax[d][t + 1])) return false;
  for (long long k = 0; k < l1; k++)
    for (long long j = 0; j < n1; j++)
      for (long long i = 0; i < n1; i++) {
        const double* p = c + 3 * ((k * n1 + j) * n1 + i);
        if (p[0] != ax[0][i] || p[1] != ax[1][j] || p[2] != ax[2][k]) return false;
      }
  // every tetrahedron against the generator's: slabs of cube index ci over the host threads (48 M tetrahedra in config 5)
  std::atomic<bool> same(true);
  const int nn = n, ll = l;
  host_parallel_for(n, 0, [&](long long c0, long long c1) {
    for (int ci = (int)c0; ci < (int)c1 && same.load(std::memory_order_relaxed); ci++)
      for (int cj = 0; cj < nn; cj++)
        for (int ck = 0; ck < ll; ck++) {
          const int (*TE)[4] = tet_grid_split_of(ci, cj, ck);
          const long long cube = ((long long)ci * nn + cj) * ll + ck;
          for (int t = 0; t < 5; t++) {
            int want[4], have[4];
            for (int q = 0; q < 4; q++) {
              const int v = TE[t][q];
              want[q] = (int)(((long long)(ck + (v >> 2)) * n1 + (cj + ((v >> 1) & 1))) * n1 + (ci + (v & 1)));
              have[q] = h.verts[(size_t)(5 * cube + t) * 4 + q];
            }
            std::sort(want, want + 4); std::sort(have, have + 4);
            for (int q = 0; q < 4; q++) if (want[q] != have[q]) { same.store(false, std::memory_order_relaxed); return; }
          }
        }
  });
  return same.load();
}
// the closed-form rows of tet_grid.cuh must be the rows of the symbolic pattern
bool tet_pattern_matches(const ScalarPattern& pat, const TetGridTables& tab, int n, int l) {
  const GridTet gt{n, l, nullptr, nullptr, nullptr, &tab};
  const long long n1 = n + 1;
  if ((long long)pat.n != n1 * n1 * (l + 1)) return false;
  std::atomic<bool> same(true);
  host_parallel_for(l + 1, 0, [&](long long k0, long long k1) {          // node planes over the host threads
    for (int k = (int)k0; k < (int)k1 && same.load(std::memory_order_relaxed); k++)
      for (int j = 0; j <= n; j++)
        for (int i = 0; i <= n; i++) {
          const long long r = ((long long)k * n1 + j) * n1 + i;
          const int mask = tg_row_mask(gt, (i + j + k) & 1, i, j, k);
          if (tg_popc(mask) != (int)(pat.rowptr[r + 1] - pat.rowptr[r])) { same.store(false, std::memory_order_relaxed); return; }
          long long at = pat.rowptr[r];
          for (int s = 0; s < 27; s++)
            if ((mask >> s) & 1) {
              const long long col = r + (long long)(s / 9 - 1) * n1 * n1 + (long long)((s / 3) % 3 - 1) * n1 + (s % 3 - 1);
              if (pat.colind[at++] != col) { same.store(false, std::memory_order_relaxed); return; }
            }
        }
  });
  return same.load();
}

bool grid_pattern_matches(const ScalarPattern& pat, int m, int n) {
  if (pat.n != (m + 1) * (n + 1)) return false;
  for (int i = 0; i <= n; i++)
    for (int j = 0; j <= m; j++) if (pat.rowptr[(size_t)i * (m + 1) + j] != grid_rowptr(i, j, m, n)) return false;
  return pat.nnz == grid_rowptr(n, m + 1, m, n);
}

int nthreads_of(const adfem_mesh* m) { return m->opt_threads > 0 ? m->opt_threads : default_threads(); }

// Tile-kernel launch shape.  P1 scalar operators and 2-D P1 elasticity compile to <= 64 registers: 3 CTAs of 320 threads per SM, 72 KB of
// shared memory each (scripts/gpu_sweep*.sh).  The P2 and the 3-D elasticity kernels need 116-128 registers; at 320 threads only ONE CTA
// fits the register file (seen in ncu: occupancy limit 1, 15 % achieved), so they run 256 threads x 2 CTAs with 110 KB tiles.
bool heavy_kernel(const adfem_mesh* m, int nc) { return m->hm.degree == 2 || (nc > 1 && m->hm.dim == 3); }
int tile_threads_of(const adfem_mesh* m, int nc) { return m->opt_tile_threads > 0 ? m->opt_tile_threads : (heavy_kernel(m, nc) ? 256 : 320); }
// adjoint = true: the P2 scalar adjoint runs one CTA of 200 KB per SM.  Measured (gpurun r2q / r2r / r2t), adjoint of the P2 Laplace operator:
//   2 M triangles, random numbering:        104 KB 0.406 ms, 72 KB 0.329 ms, 200 KB 0.335 ms (the forward is flat over all shapes)
//   16 M triangles, random numbering:       104 KB 3.15 ms,  72 KB 2.95 ms,  200 KB 2.71 ms
//   16 M triangles, Morton element order:   104 KB 2.23 ms,  72 KB 2.61 ms,  148 KB 2.64 ms, 200 KB 2.24 ms
size_t smem_budget_of(const adfem_mesh* m, int nc, bool adjoint = false) {
  if (adjoint && m->opt_smem_budget_adj > 0) return (size_t)m->opt_smem_budget_adj;
  if (m->opt_smem_budget > 0) return (size_t)m->opt_smem_budget;
  if (!heavy_kernel(m, nc)) return 72 * 1024;
  if (m->hm.degree == 2 && nc == 1) return (adjoint ? 200 : 104) * 1024;    // P2 scalar kernels keep up to 5 KB of static moment tables
  return 110 * 1024;
}

// device copy of the slot -> nnz map (SoA [d*d][ne]) for k_csr_adj_gather, built on first use
int ensure_dev_slot_nnz(adfem_mesh* m) {
  if (m->dpat.slot_nnz) return 0;
  const HostMesh& h = m->hm;
  const int dd = h.d * h.d;
  std::vector<uint32_t> soa((size_t)h.ne * dd);
  for (int e = 0; e < h.ne; e++)
    for (int s = 0; s < dd; s++) soa[(size_t)s * h.ne + e] = m->pat.slot_nnz[(size_t)e * dd + s];
  CU_TRY(upload(m->d_slot_nnz, soa));
  m->dpat.slot_nnz = m->d_slot_nnz.p;
  return 0;
}

int ensure_pattern(adfem_mesh* m) {
  if (m->has_pattern) return 0;
  // the structured triangulation (rectilinear or mapped) gets its tables from index arithmetic; option "structured_pattern" = 0: the general build
  std::string err = (m->grid_ok && m->opt_grid_pattern) ? m->pat.build_tri_grid(m->hm, m->grid_m, m->grid_n, nthreads_of(m))
                    : (m->tet_ok && m->opt_grid_pattern) ? m->pat.build_tet_grid(m->hm, m->tet_n, m->tet_l, nthreads_of(m))
                                                         : m->pat.build(m->hm, nthreads_of(m));
  if (!err.empty()) return fail("symbolic: " + err);
  m->has_pattern = true;
  if (m->grid_ok && !grid_pattern_matches(m->pat, m->grid_m, m->grid_n)) m->grid_ok = false;
  if (m->tet_ok && !tet_pattern_matches(m->pat, m->tet_tab, m->tet_n, m->tet_l)) m->tet_ok = false;
  m->rg_rows = 0;
  for (int rows = RG_THREADS; rows >= 32 && !m->rg_rows; rows /= 2) {
    bool fits = true;
    for (long long r0 = 0; r0 < m->pat.n && fits; r0 += rows)
      if (m->pat.rowptr[std::min<long long>(r0 + rows, m->pat.n)] - m->pat.rowptr[r0] > RG_CAP) fits = false;
    if (fits) m->rg_rows = rows;
  }
  m->rge_max_entries = 0;
  for (long long r0 = 0; r0 < m->pat.n; r0 += RGE_THREADS)
    m->rge_max_entries = std::max<long long>(m->rge_max_entries, m->pat.rowptr[std::min<long long>(r0 + RGE_THREADS, m->pat.n)] - m->pat.rowptr[r0]);
  if (!m->host_only) {
    const HostMesh& h = m->hm;
    const int dd = h.d * h.d;
    CU_TRY(upload(m->d_rowptr, m->pat.rowptr));
    CU_TRY(upload(m->d_colind, m->pat.colind));
    CU_TRY(upload(m->d_adj_ptr, m->pat.adj_ptr));
    CU_TRY(upload(m->d_adj_elem, m->pat.adj_elem));
    CU_TRY(upload(m->d_adj_loc, m->pat.adj_loc));
    (void)dd;       // the slot -> nnz map goes to the device only when the direct-gather adjoint needs it (ensure_dev_slot_nnz): 4 d^2 bytes per element
    m->dpat.n = m->pat.n; m->dpat.nnz = m->pat.nnz;
    m->dpat.rowptr = m->d_rowptr.p; m->dpat.colind = m->d_colind.p; m->dpat.slot_nnz = nullptr;
  }
  return 0;
}

// doubles of shared memory per tile element in the forward kernel: the local matrix (packed symmetric for scalar operators) plus, for the
// scalar operators that stage their coefficients asynchronously, two buffers of g coefficients
bool coef_staged(const HostMesh& h) { return h.degree != 1 || h.g > PIPE_GMAX; }       // scalar operators that cannot use the register prefetch
// presum3d: 3-D P1 elasticity with option "coef_presum": the Gauss-summed blocks (ns*ns doubles per element) are small enough to be staged too
bool tile_overlap_of(const adfem_mesh* m) { return m->opt_tile_overlap > 0 || (m->opt_tile_overlap < 0 && m->hm.degree == 2); }
int slots_of(const HostMesh& h, int nc, bool presum3d = false, bool overlap = false) {
  if (nc == 1) return (overlap ? 2 : 1) * (h.d * (h.d + 1) / 2) + (coef_staged(h) ? h.g : 0);     // + a staging buffer of g coefficients; overlap: two local-matrix buffers
  const int ns = h.dim == 2 ? 3 : 6;
  if (presum3d) return (nc * h.d) * (nc * h.d) + ns * ns;
  // 2-D P1 elasticity stages the raw coefficient blocks (ns*ns*g doubles per element) asynchronously
  return (nc * h.d) * (nc * h.d) + ((h.dim == 2 && h.degree == 1) ? ns * ns * h.g : 0);
}

// dynamic shared memory of the tile kernels: 3 head buffers + 2 body buffers + local matrices / 2 staging buffers
size_t fwd_smem_of(size_t max_head, size_t max_body, int max_elems, int slots) { return 3 * align16(max_head) + 2 * align16(max_body) + (size_t)8 * slots * max_elems; }
size_t fwd_smem_bytes(const FwdTiles& tp, int slots) { return fwd_smem_of(tp.max_head, tp.max_body, tp.max_elems, slots); }
// ns2: NS*NS for P1 elasticity (gradient matrices parked in shared memory for the coalesced store), else 0
size_t adj_smem_of(size_t max_head, size_t max_body, int max_elems, int max_nnz, int nc, int ns2) {
  return 3 * align16(max_head) + 2 * align16(max_body) + (size_t)2 * 8 * nc * nc * max_nnz + (size_t)8 * ns2 * max_elems;
}
size_t adj_smem_bytes(const AdjTiles& ap, int nc, int ns2) { return adj_smem_of(ap.max_head, ap.max_body, ap.max_elems, ap.max_nnz, nc, ns2); }
int adj_ns2(const HostMesh& h, int nc) { return (nc > 1 && h.degree == 1) ? (h.dim == 2 ? 9 : 36) : 0; }
constexpr size_t SMEM_LIMIT = 220 * 1024;

bool presum3d_of(const adfem_mesh* m, int nc) { return m->opt_coef_presum && nc > 1 && m->hm.dim == 3 && m->hm.degree == 1 && m->hm.g > 1; }

int ensure_fwd_plan(adfem_mesh* m, int nc, FwdPlanDev** out) {
  if (int rc = ensure_pattern(m)) return rc;
  const HostMesh& h = m->hm;
  const int slots = slots_of(h, nc, presum3d_of(m, nc), nc == 1 && tile_overlap_of(m)), dd = h.d * h.d;
  auto it = m->fwd_plans.find(nc);
  if (it != m->fwd_plans.end()) { *out = it->second.get(); return 0; }
  auto P = std::make_unique<FwdPlanDev>();
  size_t budget = smem_budget_of(m, nc);
  std::string err = "tile too large";
  for (int attempt = 0; attempt < 4 && !err.empty(); attempt++, budget = std::min(SMEM_LIMIT, budget * 2)) {
    // bytes per tile element: local matrix + 3 head copies of its id + 2 body copies of (vertex ids, vertices, sources, destinations)
    const double per_elem = 8.0 * slots + 3 * 4 + 2 * (2 * (h.dim + 1) + 8 * h.dim * 0.6 + 2.0 * dd * (nc == 1 ? 0.7 : 1.1) + 4);
    int max_elems = std::max(4, std::min(65535 / std::max(slots, dd), (int)(budget / per_elem)));
    double elems_per_row = (double)h.ne * h.d / std::max(1, h.ndof);
    int R = m->opt_rows_per_tile > 0 ? m->opt_rows_per_tile : (int)(max_elems / std::max(1.0, elems_per_row) * h.d * 0.8);
    R = std::max(4, std::min(R, 4096));
    // Triangles: the first tries of the shrink sequence below have never fitted (halo elements of a row tile: measured on structured, jittered and
    // renumbered meshes), and each try is a full plan build — 2/3 of the 40 s the 16 M-triangle P2 plan took.  Start where the sequence ends up.
    if (m->opt_rows_per_tile <= 0 && h.dim == 2 && attempt == 0)
      for (int skip = (h.degree == 2 || nc > 1) ? 2 : 1; skip > 0; skip--) R = std::max(4, (int)(R * 0.8));
    // a plan over the limit below is rejected after the build: let the build stop at the first tile that shows it
    const size_t limit = std::max(budget, m->opt_rows_per_tile > 0 ? SMEM_LIMIT : budget);
    P->host.too_big = [=](size_t hd, size_t bd, int me, int) { return fwd_smem_of(hd, bd, me, slots) > limit; };
    for (int tries = 0; tries < 16; tries++) {
      err = P->host.build(h, m->pat, R, max_elems, nc == 1 ? 1 : 0, nthreads_of(m));
      if (getenv("ADFEM_DEBUG_PLAN")) fprintf(stderr, "fwd plan try %d: R=%d max_elems=%d -> '%s' smem %zu (budget %zu) tiles %d\n", tries, R, max_elems, err.c_str(), err.empty() ? fwd_smem_bytes(P->host, slots) : (size_t)0, budget, P->host.ntiles);
      if (err.empty() && fwd_smem_bytes(P->host, slots) > std::max(budget, m->opt_rows_per_tile > 0 ? SMEM_LIMIT : budget)) err = "tile too large";
      if (err.empty() || R <= 4) break;
      R = std::max(4, (int)(R * 0.8));
    }
  }
  if (!err.empty()) return fail("forward tile plan: " + err);
  FwdTiles& tp = P->host;
  P->bytes = tp.blob.size() + 8 * tp.blob_ptr.size();
  if (!m->host_only) {
    CU_TRY(upload(P->blob_ptr, tp.blob_ptr));
    CU_TRY(upload(P->blob, tp.blob));
    P->dev = DevTiles{tp.ntiles, tp.sym, (unsigned)align16(tp.max_head), (unsigned)align16(tp.max_body), tp.max_elems, tp.max_nnz, P->blob_ptr.p, P->blob.p};
    std::vector<uint8_t>().swap(tp.blob);
  }
  *out = P.get();
  m->fwd_plans[nc] = std::move(P);
  return 0;
}

// *out stays null (rc 0) when the mesh has rows the tiled adjoint cannot index (> 255 entries): the caller uses the gather kernel
int ensure_adj_plan(adfem_mesh* m, int nc, AdjPlanDev** out) {
  *out = nullptr;
  if (int rc = ensure_pattern(m)) return rc;
  const HostMesh& h = m->hm;
  auto it = m->adj_plans.find(nc);
  if (it != m->adj_plans.end()) { *out = it->second.get(); return 0; }
  if (m->adj_untileable) return 0;
  auto P = std::make_unique<AdjPlanDev>();
  size_t budget = smem_budget_of(m, nc, true);
  std::string err = "tile too large";
  for (int attempt = 0; attempt < 4 && !err.empty(); attempt++, budget = std::min(SMEM_LIMIT, budget * 2)) {
    const int dd = h.d * h.d;
    double nnz_per_row = (double)m->pat.nnz / std::max(1, m->pat.n), rows_per_elem = (double)m->pat.n / std::max(1, h.ne);
    // bytes per owned element: 2 staging buffers + 3 head copies (row tables) + 2 body copies (ids, vertices, positions)
    const double per_elem = rows_per_elem * 1.3 * (nnz_per_row * (2 * 8.0 * nc * nc + 3 * 1) + 3 * 6) + 2 * (4 + 2 * (h.dim + 1) + 8 * h.dim * rows_per_elem + 1.0 * dd + (h.degree == 2 ? 2 * h.d : 0));
    int EPT = m->opt_elems_per_tile > 0 ? m->opt_elems_per_tile : (int)(budget / per_elem);
    EPT = std::max(4, std::min(EPT, 4096));
    // every thread of the CTA gets the same number of elements
    if (m->opt_elems_per_tile <= 0 && EPT >= tile_threads_of(m, nc)) EPT -= EPT % tile_threads_of(m, nc);
    const int max_nnz = 65535;
    const size_t limit = std::max(budget, m->opt_elems_per_tile > 0 ? SMEM_LIMIT : budget);
    const int ns2 = adj_ns2(h, nc);
    P->host.too_big = [=](size_t hd, size_t bd, int me, int mn) { return adj_smem_of(hd, bd, me, mn, nc, ns2) > limit; };
    for (int tries = 0; tries < 16; tries++) {
      err = P->host.build(h, m->pat, EPT, max_nnz, nthreads_of(m));
      if (getenv("ADFEM_DEBUG_PLAN")) fprintf(stderr, "adj plan try %d: EPT=%d -> '%s' smem %zu (budget %zu) tiles %d\n", tries, EPT, err.c_str(), err.empty() ? adj_smem_bytes(P->host, nc, adj_ns2(h, nc)) : (size_t)0, budget, P->host.ntiles);
      if (err == "row longer than 255 entries") { m->adj_untileable = true; return 0; }
      if (err.empty() && adj_smem_bytes(P->host, nc, adj_ns2(h, nc)) > std::max(budget, m->opt_elems_per_tile > 0 ? SMEM_LIMIT : budget)) err = "tile too large";
      if (err.empty() || EPT <= 4) break;
      EPT = std::max(4, (int)(EPT * 0.8));
    }
  }
  if (!err.empty()) return fail("adjoint tile plan: " + err);
  AdjTiles& ap = P->host;
  P->bytes = ap.blob.size() + 8 * ap.blob_ptr.size();
  if (!m->host_only) {
    CU_TRY(upload(P->blob_ptr, ap.blob_ptr));
    CU_TRY(upload(P->blob, ap.blob));
    P->dev = DevTiles{ap.ntiles, 0, (unsigned)align16(ap.max_head), (unsigned)align16(ap.max_body), ap.max_elems, ap.max_nnz, P->blob_ptr.p, P->blob.p};
    std::vector<uint8_t>().swap(ap.blob);
  }
  *out = P.get();
  m->adj_plans[nc] = std::move(P);
  return 0;
}

// ---- (dim, degree) dispatch ------------------------------------------------------------------------
#define DISPATCH_ELEM(m, CALL)                                             \
  do {                                                                     \
    const int _dim = (m)->hm.dim, _deg = (m)->hm.degree;                   \
    if (_dim == 2 && _deg == 1) { CALL(2, 1); }                            \
    else if (_dim == 2 && _deg == 2) { CALL(2, 2); }                       \
    else if (_dim == 3 && _deg == 1) { CALL(3, 1); }                       \
    else { CALL(3, 2); }                                                   \
  } while (0)

inline unsigned blocks_for(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// grid of a tile kernel: persistent (resident CTAs per SM x SMs, never more than the tiles) or one CTA per tile
template <class K> int tile_grid(adfem_mesh* m, K kern, int threads, size_t smem, int ntiles, int* grid) {
  if (m->opt_pipeline == 0) { *grid = ntiles; return 0; }
  if (m->num_sms == 0) CU_TRY(cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, m->device));
  int per_sm = 0;
  CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) return fail("tile kernel does not fit on an SM (threads / shared memory)");
  *grid = std::max(1, std::min(ntiles, per_sm * m->num_sms));
  if (m->opt_grid_limit > 0) *grid = std::min(*grid, m->opt_grid_limit);
  return 0;
}

DevMesh dev_mesh(const adfem_mesh* m, int heron) { DevMesh d = m->dm; d.heron = heron; return d; }
// the mesh as the tile kernels see it when the coefficients are already summed over the Gauss points (option "coef_presum"): one
// "Gauss point" per element with reference weight 1 (P1: the position is irrelevant)
DevMesh dev_mesh_presum(const adfem_mesh* m, int heron) {
  DevMesh d = dev_mesh(m, heron);
  d.g = 1; d.rule.n = 1; d.rule.w[0] = 1.0;
  return d;
}
bool use_presum(const adfem_mesh* m, int op) { return m->opt_coef_presum && op == ADFEM_OP_STIFFNESS && m->hm.degree == 1 && m->hm.g > 1; }
int ensure_presum_buf(adfem_mesh* m) {
  const size_t n = (size_t)m->hm.ne * (m->hm.dim == 2 ? 9 : 36);
  if (m->presum_buf.n >= n) return 0;
  CU_TRY(m->presum_buf.alloc(n));
  return 0;
}

template <class K>
int launch_fwd_kernel(adfem_mesh* m, K kern, FwdPlanDev* P, size_t smem, int threads, const double* coef, double* vals, cudaStream_t st, bool presum = false) {
  int grid = 0;
  CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (int rc = tile_grid(m, kern, threads, smem, P->dev.ntiles, &grid)) return rc;
  kern<<<grid, threads, smem, st>>>(presum ? dev_mesh_presum(m, m->opt_area_csr) : dev_mesh(m, m->opt_area_csr), m->pat.nnz, P->dev, coef, vals);
  CU_TRY(cudaGetLastError());
  return 0;
}

template <int DIM, int DEG, int OP>
int launch_tile_fwd(adfem_mesh* m, FwdPlanDev* P, const double* coef, double* vals, cudaStream_t st, bool presum = false) {
  constexpr int NC = OP == OP_STIFFNESS ? DIM : 1;
  const bool overlap = NC == 1 && tile_overlap_of(m);
  const size_t smem = fwd_smem_bytes(P->host, slots_of(m->hm, NC, presum && presum3d_of(m, NC), overlap));
  const int threads = tile_threads_of(m, NC);
  if (smem > SMEM_LIMIT) return fail("forward tile needs more shared memory than an SM has");
  if constexpr (OP != OP_STIFFNESS) {
    if (overlap && m->opt_coef_prefetch) {
      if constexpr (DEG == 1) {
        if (m->hm.g <= PIPE_GMAX && P->host.max_elems <= PIPE_EPT * threads) return launch_fwd_kernel(m, k_tile_fwd_ov<DIM, DEG, OP, true>, P, smem, threads, coef, vals, st);
      }
      if (coef_staged(m->hm)) return launch_fwd_kernel(m, k_tile_fwd_ov<DIM, DEG, OP, false>, P, smem, threads, coef, vals, st);
    }
  }
  if constexpr (DEG == 1 && OP != OP_STIFFNESS) {
    // register prefetch of the next tile's coefficients
    if (m->opt_coef_prefetch && m->hm.g <= PIPE_GMAX && P->host.max_elems <= PIPE_EPT * threads)
      return launch_fwd_kernel(m, k_tile_fwd<DIM, DEG, OP, true, false>, P, smem, threads, coef, vals, st);
  }
  if constexpr (OP != OP_STIFFNESS) {
    if (coef_staged(m->hm) && m->opt_coef_prefetch) return launch_fwd_kernel(m, k_tile_fwd<DIM, DEG, OP, false, true>, P, smem, threads, coef, vals, st);
  }
  if constexpr (OP == OP_STIFFNESS && DIM == 2 && DEG == 1) {
    if (m->opt_coef_prefetch) return launch_fwd_kernel(m, k_tile_fwd<DIM, DEG, OP, false, true>, P, smem, threads, coef, vals, st, presum);
  }
  if constexpr (OP == OP_STIFFNESS && DIM == 3 && DEG == 1) {
    // Gauss-summed blocks (option "coef_presum"): 36 doubles per element, staged with the flat coalesced index like the 2-D blocks
    if (presum && presum3d_of(m, NC) && m->opt_coef_prefetch) return launch_fwd_kernel(m, k_tile_fwd<DIM, DEG, OP, false, true>, P, smem, threads, coef, vals, st, presum);
  }
  return launch_fwd_kernel(m, k_tile_fwd<DIM, DEG, OP, false, false>, P, smem, threads, coef, vals, st, presum);
}
template <int DIM, int DEG, int OP>
int launch_adj(adfem_mesh* m, const double* dvals, double* grad, cudaStream_t st, bool presum = false) {
  constexpr int NC = OP == OP_STIFFNESS ? DIM : 1;
  AdjPlanDev* P = nullptr;
  if (m->opt_adjoint_tiled) { if (int rc = ensure_adj_plan(m, NC, &P)) return rc; }
  const size_t smem = P ? adj_smem_bytes(P->host, NC, adj_ns2(m->hm, NC)) : 0;
  if (P && smem <= SMEM_LIMIT) {
    auto kern = k_tile_adj<DIM, DEG, OP>;
    int grid = 0;
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (int rc = tile_grid(m, kern, tile_threads_of(m, NC), smem, P->dev.ntiles, &grid)) return rc;
    kern<<<grid, tile_threads_of(m, NC), smem, st>>>(presum ? dev_mesh_presum(m, m->opt_area_csr) : dev_mesh(m, m->opt_area_csr), m->pat.nnz, P->dev, dvals, grad);
  } else {
    if (int rc = ensure_dev_slot_nnz(m)) return rc;
    k_csr_adj_gather<DIM, DEG, OP><<<blocks_for(m->hm.ne, 128), 128, 0, st>>>(presum ? dev_mesh_presum(m, m->opt_area_csr) : dev_mesh(m, m->opt_area_csr), m->dpat, dvals, grad);
  }
  CU_TRY(cudaGetLastError());
  return 0;
}

int ensure_host_streams(adfem_mesh* m) {
  if (m->hs[0]) return 0;
  for (auto& s : m->hs) CU_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  for (auto& e : m->hev) CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return 0;
}

bool use_grid_any(adfem_mesh* m) { return m->grid_ok && m->opt_structured && !m->host_only; }      // scalar CSR operators and source term: rectilinear or mapped
bool use_grid(adfem_mesh* m) { return use_grid_any(m) && !m->grid_mapped; }                        // rectilinear only (no caller left: every structured kernel has a MAPPED form)

int grid_rows_per_warp(adfem_mesh* m, int strips, int rows) {
  if (m->opt_grid_rows > 0) return m->opt_grid_rows;
  // ~8 full waves of resident CTAs (2 per SM): short enough tails, long enough marches (one extra cell row per chunk is recomputed)
  if (m->num_sms == 0 && cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, m->device) != cudaSuccess) m->num_sms = 148;
  const long long want_warps = 8LL * std::max(2, std::min(3, m->opt_grid_occupancy)) * m->num_sms * GRID_WARPS;
  const int want_chunks = (int)std::max<long long>(1, want_warps / std::max(1, strips));
  return std::max(8, (rows + want_chunks - 1) / want_chunks);
}

template <class K> int launch_grid(K kern, int smem, long long warps, cudaStream_t st, const DevMesh& dm, const GridTri& gt, int r0, int r1, int H, const double* in,
                                   double* out) {
  CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<(unsigned)((warps + GRID_WARPS - 1) / GRID_WARPS), GRID_WARPS * 32, smem, st>>>(dm, gt, r0, r1, H, in, out);
  CU_TRY(cudaGetLastError());
  return 0;
}
// node rows [r0, r1) of the structured forward (r1 < 0: all rows)
template <int OP> int launch_grid_fwd(adfem_mesh* m, const double* coef, double* vals, cudaStream_t st, int r0 = 0, int r1 = -1) {
  GridTri gt{m->grid_m, m->grid_n, m->grid_xs.p, m->grid_ys.p};
  if (r1 < 0) r1 = gt.n + 1;
  const int strips = (gt.m + 1 + GRID_STRIP - 1) / GRID_STRIP, H = grid_rows_per_warp(m, strips, r1 - r0), chunks = (r1 - r0 + H - 1) / H;
  const long long warps = (long long)strips * chunks;
  if (m->grid_mapped) return launch_grid(k_grid_fwd<OP, 2, true>, GRID_FWD_SMEM, warps, st, dev_mesh(m, m->opt_area_csr), gt, r0, r1, H, coef, vals);
  if (m->opt_grid_occupancy >= 3) return launch_grid(k_grid_fwd<OP, 3>, GRID_FWD_SMEM, warps, st, dev_mesh(m, m->opt_area_csr), gt, r0, r1, H, coef, vals);
  return launch_grid(k_grid_fwd<OP, 2>, GRID_FWD_SMEM, warps, st, dev_mesh(m, m->opt_area_csr), gt, r0, r1, H, coef, vals);
}
// cell rows [r0, r1) of the structured adjoint (r1 < 0: all rows)
template <int OP> int launch_grid_adj(adfem_mesh* m, const double* dvals, double* grad, cudaStream_t st, int r0 = 0, int r1 = -1) {
  GridTri gt{m->grid_m, m->grid_n, m->grid_xs.p, m->grid_ys.p};
  if (r1 < 0) r1 = gt.n;
  const int strips = (gt.m + GRID_STRIP - 1) / GRID_STRIP, H = grid_rows_per_warp(m, strips, r1 - r0), chunks = (r1 - r0 + H - 1) / H;
  const long long warps = (long long)strips * chunks;
  if (m->grid_mapped) return launch_grid(k_grid_adj<OP, 2, true>, GRID_ADJ_SMEM, warps, st, dev_mesh(m, m->opt_area_csr), gt, r0, r1, H, dvals, grad);
  if (m->opt_grid_occupancy >= 3) return launch_grid(k_grid_adj<OP, 3>, GRID_ADJ_SMEM, warps, st, dev_mesh(m, m->opt_area_csr), gt, r0, r1, H, dvals, grad);
  return launch_grid(k_grid_adj<OP, 2>, GRID_ADJ_SMEM, warps, st, dev_mesh(m, m->opt_area_csr), gt, r0, r1, H, dvals, grad);
}

// P1 elasticity on the structured triangulation (grid_elast.cuh): ~8 waves of resident warps, like the scalar kernels
bool use_grid_elast(adfem_mesh* m, int op) { return op == ADFEM_OP_STIFFNESS && m->opt_grid_elast && use_grid_any(m) && m->hm.degree == 1 && m->hm.g == GE_G; }
// plane_mode < 0: tangents H in / dH out (in = H or dvals, out = vals or grad_H).  plane_mode = 0 | 1: fused constitutive step — forward
// (in = E, in2 = nu) -> out = vals; adjoint (in = dvals; E, nu) -> out = dE, out2 = dnu.
int launch_grid_elast(adfem_mesh* m, bool adjoint, const double* in, double* out, cudaStream_t st, int plane_mode = -1, const double* E = nullptr,
                      const double* nu = nullptr, double* out2 = nullptr) {
  GridTri gt{m->grid_m, m->grid_n, m->grid_xs.p, m->grid_ys.p};
  if (m->num_sms == 0 && cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, m->device) != cudaSuccess) m->num_sms = 148;
  const int rows = adjoint ? gt.n : gt.n + 1;
  const int strips = adjoint ? (gt.m + GE_COLS - 1) / GE_COLS : (gt.m + 1 + GE_COLS - 1) / GE_COLS;
  const size_t smem = (size_t)GE_WARPS * (adjoint ? GE_ADJ_WARP_DOUBLES : GE_FWD_WARP_DOUBLES) * sizeof(double);
  int rpw = m->opt_grid_rows;
  if (rpw <= 0) {
    const long long want_warps = 8LL * 3 * m->num_sms * GE_WARPS;
    const int want_chunks = (int)std::max<long long>(1, want_warps / std::max(1, strips));
    rpw = std::max(8, (rows + want_chunks - 1) / want_chunks);
  }
  const long long warps = (long long)strips * ((rows + rpw - 1) / rpw);
  const unsigned blocks = (unsigned)((warps + GE_WARPS - 1) / GE_WARPS);
  const DevMesh dm = dev_mesh(m, m->opt_area_csr);
  const bool plane = plane_mode >= 0;
  const bool mp = m->grid_mapped;      // node positions from the coordinate array (structured connectivity on a mapped / jittered grid)
  if (adjoint) {
    auto kern = plane ? (mp ? k_grid_elast_adj<true, true> : k_grid_elast_adj<true, false>) : (mp ? k_grid_elast_adj<false, true> : k_grid_elast_adj<false, false>);
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, GE_WARPS * 32, smem, st>>>(dm, gt, m->pat.nnz, rpw, plane_mode, E, nu, in, out, out2);
  } else {
    auto kern = plane ? (mp ? k_grid_elast_fwd<true, true> : k_grid_elast_fwd<true, false>) : (mp ? k_grid_elast_fwd<false, true> : k_grid_elast_fwd<false, false>);
    CU_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, GE_WARPS * 32, smem, st>>>(dm, gt, m->pat.nnz, rpw, plane_mode, plane ? E : in, nu, out);
  }
  CU_TRY(cudaGetLastError());
  return 0;
}

// P1 tetrahedral elasticity forward on the structured grid Mesh3(n, n, l, h) (tet_grid.cuh): Gauss-sum pre-pass, then one warp per node
bool use_tet_grid(adfem_mesh* m, int op) {
  return op == ADFEM_OP_STIFFNESS && m->opt_grid_elast && m->opt_structured && m->tet_ok && !m->host_only && m->hm.degree == 1;
}
int launch_tet_grid_fwd(adfem_mesh* m, const double* coef, double* vals, cudaStream_t st) {
  if (m->opt_tet_node) {
    const size_t need = tn_scratch_doubles(m->tet_n, m->tet_l);
    if (m->presum_buf.n < need) CU_TRY(m->presum_buf.alloc(need));
    if (!m->tet_const_ready) {
      TetNodeConst C; build_tet_node_const(m->tet_tab, C);
      CU_TRY(cudaMemcpyToSymbol(c_tn, &C, sizeof(C)));
      CU_TRY(cudaFuncSetAttribute(k_tet_node_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM_BYTES));
      CU_TRY(cudaFuncSetAttribute(k_tet_node_fwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM_BYTES2));
      std::vector<double> spc;
      for (int inv = 1; inv >= 0; inv--)
        for (int a = 0; a < 3; a++)
          for (size_t c = 0; c + 1 < m->tet_axes[a].size(); c++) { const double hc = m->tet_axes[a][c + 1] - m->tet_axes[a][c]; spc.push_back(inv ? 1.0 / hc : hc); }
      CU_TRY(upload(m->tet_spacing, spc));
      m->tet_const_ready = true;
    }
    const int n = m->tet_n, l = m->tet_l, nxb = (n + TPX_CUBES - 1) / TPX_CUBES;
    const GridTet gt{n, l, m->tet_xs.p, m->tet_ys.p, m->tet_zs.p, m->d_tet_tab.p};
    const double* q = m->tet_spacing.p;
    const TetSpacing sp{q, q + n, q + 2 * n, q + 2 * n + l, q + 3 * n + l, q + 4 * n + l};
    const long long plane = (long long)(n + 1) * (n + 1);
    // z-chunks: the Gauss pre-sum of cube layers [z0, z1) runs on the caller's stream, the node kernel of node planes [z0, z1) (the last chunk also
    // takes plane l) follows on a side stream as soon as its layers are summed, so the DRAM-bound pass of chunk c + 1 overlaps the
    // instruction-bound pass of chunk c.  One chunk = plain back-to-back launches on the caller's stream.
    int nchunk = m->opt_tet_chunks > 0 ? m->opt_tet_chunks : 1;      // measured (gpurun r2g, 10.5 M tetrahedra): 1 / 4 / 8 / 16 chunks = 5.55 / 5.68 / 5.65 / 5.61 ms: no gain, so off by default
    nchunk = std::max(1, std::min(nchunk, l));
    if (nchunk > 1 && !m->tet_side_stream) {
      CU_TRY(cudaStreamCreateWithFlags(&m->tet_side_stream, cudaStreamNonBlocking));
      m->tet_events.resize(34);
      for (auto& e : m->tet_events) CU_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    nchunk = std::min(nchunk, 32);
    cudaStream_t s2 = nchunk > 1 ? m->tet_side_stream : st;
    if (nchunk > 1) { CU_TRY(cudaEventRecord(m->tet_events[32], st)); CU_TRY(cudaStreamWaitEvent(s2, m->tet_events[32], 0)); }   // the side stream starts after the caller's earlier work (vals may still be read)
    for (int c = 0; c < nchunk; c++) {
      const int z0 = (int)((long long)l * c / nchunk), z1 = (int)((long long)l * (c + 1) / nchunk), nz = z1 - z0;
      const unsigned grid = (unsigned)((size_t)n * nz * nxb * 5);
      if (m->hm.g == 4) k_tet_presum_xg<4><<<grid, TPX_THREADS, 0, st>>>(n, l, z0, nz, m->hm.rule, coef, m->presum_buf.p);
      else k_tet_presum_x<<<grid, TPX_THREADS, 0, st>>>(n, l, z0, nz, m->hm.rule, m->hm.g, coef, m->presum_buf.p);
      if (nchunk > 1) { CU_TRY(cudaEventRecord(m->tet_events[c], st)); CU_TRY(cudaStreamWaitEvent(s2, m->tet_events[c], 0)); }
      // node plane k needs cube layers k - 1 and k: planes [z0, z1) are complete once layers < z1 are summed (plane z1 waits for the next chunk)
      const long long node0 = plane * z0, node1 = (c + 1 == nchunk) ? (long long)m->hm.nv : plane * z1;
      // (measured and dropped, gpurun r2i: one CTA per parity — 3 warps, no barrier between the 32- and the 8-tetrahedron warps — 5.55 -> 6.79 ms at
      // 10.5 M tetrahedra, 6.03 ms as two launches: the two parities of a span share their tangent blocks through L1)
      if (m->opt_tet_node == 2) k_tet_node_fwd<2><<<blocks_for(node1 - node0, TN_NODES), 9 * 32, TN_SMEM_BYTES2, s2>>>(gt, sp, m->pat.nnz, node0, node1, m->d_rowptr.p, m->presum_buf.p, vals);
      else k_tet_node_fwd<1><<<blocks_for(node1 - node0, TN_NODES), 6 * 32, TN_SMEM_BYTES, s2>>>(gt, sp, m->pat.nnz, node0, node1, m->d_rowptr.p, m->presum_buf.p, vals);
    }
    if (nchunk > 1) { CU_TRY(cudaEventRecord(m->tet_events[33], s2)); CU_TRY(cudaStreamWaitEvent(st, m->tet_events[33], 0)); }
    CU_TRY(cudaGetLastError());
    return 0;
  }
  if (int rc = ensure_presum_buf(m)) return rc;
  if (int rc = launch_presum_coef(dev_mesh(m, m->opt_area_csr), 36, coef, m->presum_buf.p, st)) return rc;
  const GridTet gt{m->tet_n, m->tet_l, m->tet_xs.p, m->tet_ys.p, m->tet_zs.p, m->d_tet_tab.p};
  const size_t smem = (size_t)TG_WARPS * TG_WARP_DOUBLES * sizeof(double);
  CU_TRY(cudaFuncSetAttribute(k_tet_grid_elast_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_tet_grid_elast_fwd<<<blocks_for(m->hm.nv, TG_WARPS), TG_WARPS * 32, smem, st>>>(gt, m->pat.nnz, m->d_rowptr.p, m->presum_buf.p, vals);
  CU_TRY(cudaGetLastError());
  return 0;
}

int launch_tet_grid_adj(adfem_mesh* m, const double* dvals, double* grad, cudaStream_t st) {
  const GridTet gt{m->tet_n, m->tet_l, m->tet_xs.p, m->tet_ys.p, m->tet_zs.p, m->d_tet_tab.p};
  const size_t smem = (size_t)TG_ADJ_WARPS * TG_ADJ_WARP_DOUBLES * sizeof(double);
  const long long units = (long long)m->tet_n * m->tet_n * ((5 * m->tet_l + 31) / 32);      // (cube column, chunk of 32 tetrahedra)
  // resident CTAs per SM the kernel is compiled for: 3 = 168 registers, 4 = 128, 5 = 96 (a few spills)
#define ADFEM_TET_ADJ(MINB)                                                                                                                  \
  do {                                                                                                                                       \
    CU_TRY(cudaFuncSetAttribute(k_tet_grid_elast_adj<MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                        \
    k_tet_grid_elast_adj<MINB><<<blocks_for(units, TG_ADJ_WARPS), TG_ADJ_WARPS * 32, smem, st>>>(gt, m->hm.rule, m->hm.g, (long long)m->hm.ne, \
                                                                                             m->pat.nnz, m->d_rowptr.p, dvals, grad);       \
  } while (0)
  if (m->opt_tet_adj_blocks == 4) ADFEM_TET_ADJ(4);
  else if (m->opt_tet_adj_blocks == 5) ADFEM_TET_ADJ(5);
  else ADFEM_TET_ADJ(3);
#undef ADFEM_TET_ADJ
  CU_TRY(cudaGetLastError());
  return 0;
}

// scalar P1 operators on the structured tetrahedral grid (tet_scalar.cuh), same opt-in switch as the elasticity kernels
bool use_tet_scalar(adfem_mesh* m, int op) {
  return op != ADFEM_OP_STIFFNESS && m->opt_tet_scalar && m->opt_structured && m->tet_ok && !m->host_only && m->hm.degree == 1;
}
int launch_tet_scalar(adfem_mesh* m, int op, bool adjoint, const double* in, double* out, cudaStream_t st) {
  const GridTet gt{m->tet_n, m->tet_l, m->tet_xs.p, m->tet_ys.p, m->tet_zs.p, m->d_tet_tab.p};
  if (adjoint) {
    const unsigned nb = blocks_for(m->hm.ne, 128);
    if (op == ADFEM_OP_LAPLACE) k_tet_grid_scalar_adj<OP_LAPLACE><<<nb, 128, 0, st>>>(gt, m->hm.rule, m->hm.g, (long long)m->hm.ne, m->d_rowptr.p, in, out);
    else k_tet_grid_scalar_adj<OP_MASS><<<nb, 128, 0, st>>>(gt, m->hm.rule, m->hm.g, (long long)m->hm.ne, m->d_rowptr.p, in, out);
  } else {
    const unsigned nb = blocks_for(m->hm.nv, RG_THREADS);
    if (op == ADFEM_OP_LAPLACE) k_tet_grid_scalar_fwd<OP_LAPLACE><<<nb, RG_THREADS, 0, st>>>(gt, m->hm.rule, m->hm.g, m->d_rowptr.p, in, out);
    else k_tet_grid_scalar_fwd<OP_MASS><<<nb, RG_THREADS, 0, st>>>(gt, m->hm.rule, m->hm.g, m->d_rowptr.p, in, out);
  }
  CU_TRY(cudaGetLastError());
  return 0;
}

int launch_grid_source(adfem_mesh* m, bool adjoint, const double* in, double* out, cudaStream_t st) {
  GridTri gt{m->grid_m, m->grid_n, m->grid_xs.p, m->grid_ys.p};
  const int rows = adjoint ? gt.n : gt.n + 1;
  const int strips = adjoint ? (gt.m + GRID_STRIP - 1) / GRID_STRIP : (gt.m + 1 + GRID_STRIP - 1) / GRID_STRIP;
  const int H = grid_rows_per_warp(m, strips, rows), chunks = (rows + H - 1) / H;
  const unsigned blocks = (unsigned)(((long long)strips * chunks + GRID_WARPS - 1) / GRID_WARPS);
  if (m->grid_mapped) {
    if (adjoint) k_grid_source_adj<true><<<blocks, GRID_WARPS * 32, 0, st>>>(dev_mesh(m, m->opt_area_coo), gt, 0, rows, H, in, out);
    else k_grid_source_fwd<true><<<blocks, GRID_WARPS * 32, 0, st>>>(dev_mesh(m, m->opt_area_coo), gt, 0, rows, H, in, out);
  } else if (adjoint) k_grid_source_adj<false><<<blocks, GRID_WARPS * 32, 0, st>>>(dev_mesh(m, m->opt_area_coo), gt, 0, rows, H, in, out);
  else k_grid_source_fwd<false><<<blocks, GRID_WARPS * 32, 0, st>>>(dev_mesh(m, m->opt_area_coo), gt, 0, rows, H, in, out);
  CU_TRY(cudaGetLastError());
  return 0;
}

int check_op(const adfem_mesh* m, int op) {
  if (op < 0 || op > 2) return fail("unknown op");
  (void)m;
  return 0;
}

int coef_per_gauss(const adfem_mesh* m, int op) {
  if (op != ADFEM_OP_STIFFNESS) return 1;
  return m->hm.dim == 2 ? 9 : 36;
}

}  // namespace

// ====================================================================================================
// handle API
// ====================================================================================================
extern "C" {

const char* adfem_last_error(void) { return g_err.c_str(); }

int adfem_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int adfem_mesh_create(adfem_mesh** out, int dim, const double* vertices, int vertex_stride, int nv, const int* elems, int ne,
                      int order, int degree, int lorder, int flags) {
  if (!out) return fail("null output");
  *out = nullptr;
  if (order == -1) order = degree == 1 ? 2 : 4;       // src/MFEM/MFEM.jl:71-77, src/MFEM3/MFEM.jl:49-55
  if (lorder == -1) lorder = 6;                       // src/MFEM/MFEM.jl:79-85
  auto m = std::make_unique<adfem_mesh>();
  std::string err = m->hm.build(dim, vertices, vertex_stride, nv, elems, ne, order, degree, lorder);
  if (!err.empty()) return fail(err);
  m->host_only = (flags & ADFEM_HOST_ONLY) != 0;
  std::vector<double> grid_xs, grid_ys;
  m->grid_ok = detect_tri_grid(m->hm, m->grid_m, m->grid_n, grid_xs, grid_ys, m->grid_mapped);
  m->tet_ok = detect_tet_grid(m->hm, m->tet_n, m->tet_l, m->tet_axes);
  if (m->tet_ok) build_tet_grid_tables(m->tet_tab);
  if (!m->host_only) {
    if (adfem_device_count() == 0) return fail("no CUDA device available (libadfem_cuda has no CPU fallback)");
    CU_TRY(cudaGetDevice(&m->device));
    const HostMesh& h = m->hm;
    CU_TRY(upload(m->coords, h.coords));
    const int nvl = h.dim + 1;
    {
      const std::vector<int> vs = soa_copy(h.verts, h.ne, nvl);       // [k][e]: what the kernels read
      CU_TRY(upload(m->verts, vs));
    }
    {
      const std::vector<int> cs = soa_copy(h.conn, h.ne, h.d);
      CU_TRY(upload(m->conn, cs));
    }
    m->dm.dim = h.dim; m->dm.ne = h.ne; m->dm.nv = h.nv; m->dm.d = h.d; m->dm.g = h.g; m->dm.ndof = h.ndof;
    m->dm.coords = m->coords.p; m->dm.verts = m->verts.p; m->dm.conn = m->conn.p; m->dm.rule = h.rule;
    if (m->grid_ok) { CU_TRY(upload(m->grid_xs, grid_xs)); CU_TRY(upload(m->grid_ys, grid_ys)); }
    if (m->tet_ok) {
      CU_TRY(upload(m->tet_xs, m->tet_axes[0])); CU_TRY(upload(m->tet_ys, m->tet_axes[1])); CU_TRY(upload(m->tet_zs, m->tet_axes[2]));
      CU_TRY(upload(m->d_tet_tab, std::vector<TetGridTables>(1, m->tet_tab)));
    }
  }
  *out = m.release();
  return 0;
}

void adfem_mesh_destroy(adfem_mesh* m) { delete m; }

long long adfem_mesh_info(const adfem_mesh* m, int what) {
  if (!m) return -1;
  const HostMesh& h = m->hm;
  switch (what) {
    case ADFEM_INFO_DIM: return h.dim;
    case ADFEM_INFO_NV: return h.nv;
    case ADFEM_INFO_NE: return h.ne;
    case ADFEM_INFO_NDOF: return h.ndof;
    case ADFEM_INFO_NGAUSS: return (long long)h.ne * h.g;
    case ADFEM_INFO_ELEM_NDOF: return h.d;
    case ADFEM_INFO_NEDGES: h.ensure_edges(); return h.nedges;
    case ADFEM_INFO_GAUSS_PER_ELEM: return h.g;
    case ADFEM_INFO_NNZ_SCALAR: return m->has_pattern ? m->pat.nnz : -1;
    case ADFEM_INFO_TILES_FWD: { long long t = 0; for (auto& kv : m->fwd_plans) t = kv.second->host.ntiles; return t; }
    case ADFEM_INFO_TILES_ADJ: { long long t = 0; for (auto& kv : m->adj_plans) t = kv.second->host.ntiles; return t; }
    case ADFEM_INFO_PLAN_BYTES: {
      long long b = 0;
      for (auto& kv : m->fwd_plans) b += (long long)kv.second->bytes;
      for (auto& kv : m->adj_plans) b += (long long)kv.second->bytes;
      return b;
    }
    case ADFEM_INFO_STRUCTURED: return m->grid_ok ? (m->grid_mapped ? 3 : 1) : (m->tet_ok ? 2 : 0);
    default: return -1;
  }
}

int adfem_mesh_edges(const adfem_mesh* m, long long* edges) {
  if (!m) return fail("null mesh handle");
  m->hm.ensure_edges();
  const long long ne = m->hm.nedges;
  for (long long i = 0; i < ne; i++) { edges[i] = m->hm.edge_lo[i] + 1; edges[ne + i] = m->hm.edge_hi[i] + 1; }
  return 0;
}
int adfem_mesh_connectivity(const adfem_mesh* m, long long* conn) {
  if (!m) return fail("null mesh handle");
  const size_t n = (size_t)m->hm.ne * m->hm.d;
  for (size_t i = 0; i < n; i++) conn[i] = m->hm.conn[i] + 1;
  return 0;
}
int adfem_mesh_element_to_vertices(const adfem_mesh* m, long long* elems) {
  if (!m) return fail("null mesh handle");
  const HostMesh& h = m->hm;
  const int nvl = h.dim + 1;
  for (int e = 0; e < h.ne; e++) for (int k = 0; k < nvl; k++) elems[(size_t)k * h.ne + e] = h.verts[(size_t)e * nvl + k] + 1;
  return 0;
}
int adfem_mesh_gauss(const adfem_mesh* m, double* xyz) { if (!m) return fail("null mesh handle"); m->hm.gauss_points(xyz); return 0; }
int adfem_mesh_gauss_weights(const adfem_mesh* m, double* w) { if (!m) return fail("null mesh handle"); m->hm.gauss_weights(w); return 0; }
int adfem_mesh_measure(const adfem_mesh* m, double* a) { if (!m) return fail("null mesh handle"); m->hm.measure(a); return 0; }

int adfem_set_option(adfem_mesh* m, const char* key, long long value) {
  if (!m || !key) return fail("null argument");
  std::string k(key);
  if (k == "rows_per_tile") { m->opt_rows_per_tile = (int)value; m->fwd_plans.clear(); }
  else if (k == "elems_per_tile") { m->opt_elems_per_tile = (int)value; m->adj_plans.clear(); }
  else if (k == "adjoint_tiled") m->opt_adjoint_tiled = (int)value;
  else if (k == "host_threads") m->opt_threads = (int)value;
  else if (k == "smem_budget") { m->opt_smem_budget = (int)value; m->fwd_plans.clear(); m->adj_plans.clear(); }
  else if (k == "smem_budget_adj") { m->opt_smem_budget_adj = (int)value; m->adj_plans.clear(); }
  else if (k == "tile_overlap") { if (m->opt_tile_overlap != (int)value) m->fwd_plans.clear(); m->opt_tile_overlap = (int)value; }      // -1 auto, 0 off, 1 on; the tile size depends on it
  else if (k == "tile_threads" && value == 0) { m->opt_tile_threads = 0; m->adj_plans.clear(); }          // back to the per-operator default
  else if (k == "tile_threads") { if (value < 32 || value > TILE_MAX_THREADS || value % 32) return fail("tile_threads must be a multiple of 32 in [32, 512]"); if (m->opt_tile_threads != (int)value && m->opt_elems_per_tile <= 0) m->adj_plans.clear(); m->opt_tile_threads = (int)value; }
  else if (k == "pipeline") { if (value < 0 || value > 1) return fail("pipeline must be 0 (one CTA per tile) or 1 (persistent, software-pipelined)"); m->opt_pipeline = (int)value; }
  else if (k == "coef_prefetch") m->opt_coef_prefetch = value != 0;
  else if (k == "coef_presum") { if (m->opt_coef_presum != (value != 0) && m->hm.dim == 3) m->fwd_plans.clear(); m->opt_coef_presum = value != 0; }      // 3-D: the tile size depends on it
  else if (k == "grid_limit") m->opt_grid_limit = (int)value;
  else if (k == "structured") m->opt_structured = value != 0;
  else if (k == "structured_pattern") { if (m->has_pattern) return fail("structured_pattern must be set before the symbolic phase has run"); m->opt_grid_pattern = value != 0; }
  else if (k == "structured_elasticity") m->opt_grid_elast = value != 0;
  else if (k == "tet_node") m->opt_tet_node = (int)value;
  else if (k == "tet_chunks") m->opt_tet_chunks = (int)value;
  else if (k == "tet_adj_blocks") m->opt_tet_adj_blocks = (int)value;
  else if (k == "structured_tet_scalar") m->opt_tet_scalar = value != 0;
  else if (k == "row_gather") m->opt_row_gather = value != 0;
  else if (k == "grid_rows") m->opt_grid_rows = (int)value;
  else if (k == "grid_occupancy") m->opt_grid_occupancy = (int)value;
  else if (k == "host_chunks") m->opt_host_chunks = (int)value;
  else if (k == "area_formula_csr") m->opt_area_csr = value != 0;
  else if (k == "area_formula_coo") m->opt_area_coo = value != 0;
  else return fail("unknown option: " + k);
  return 0;
}

int adfem_symbolic(adfem_mesh* m) {
  if (!m) return fail("null mesh handle");
  return ensure_pattern(m);
}

long long adfem_csr_nnz(adfem_mesh* m, int ncomp) {
  if (!m || ensure_pattern(m)) return -1;
  return m->pat.nnz * ncomp * ncomp;
}

int adfem_csr_pattern(adfem_mesh* m, int ncomp, long long* rowptr, int* colind) {
  if (!m) return fail("null mesh handle");
  if (int rc = ensure_pattern(m)) return rc;
  const ScalarPattern& p = m->pat;
  const long long n = p.n;
  if ((long long)ncomp * n > 2147483647LL) return fail("ncomp*ndof exceeds 32-bit column ids");
  host_parallel_for(n, nthreads_of(m), [&](long long r0, long long r1) {         // row blocks over the host threads (the caller's arrays are first touched here)
    for (int a = 0; a < ncomp; a++)
      for (long long r = r0; r < r1; r++) {
        const long long rs = p.rowptr[r], len = p.rowptr[r + 1] - rs, base = ncomp * (a * p.nnz + rs);
        rowptr[a * n + r] = base;
        for (int b = 0; b < ncomp; b++)
          for (long long j = 0; j < len; j++) colind[base + b * len + j] = (int)(p.colind[rs + j] + b * n);
      }
  }, 1 << 16);
  rowptr[ncomp * n] = p.nnz * ncomp * ncomp;
  return 0;
}

int adfem_csr_pattern_device(adfem_mesh* m, const long long** rowptr, const int** colind) {
  if (int rc = need_device(m)) return rc;
  if (int rc = ensure_pattern(m)) return rc;
  if (rowptr) *rowptr = m->d_rowptr.p;
  if (colind) *colind = m->d_colind.p;
  return 0;
}

int adfem_slot_to_nnz(adfem_mesh* m, unsigned int* slot_nnz) {
  if (!m) return fail("null mesh handle");
  if (int rc = ensure_pattern(m)) return rc;
  memcpy(slot_nnz, m->pat.slot_nnz.data(), m->pat.slot_nnz.size() * sizeof(uint32_t));
  return 0;
}

long long adfem_plan_array(adfem_mesh* m, int which_plan, int ncomp, int array_id, void* out) {
  if (!m) { fail("null mesh handle"); return -1; }
  if (!m->host_only) { fail("adfem_plan_array needs an ADFEM_HOST_ONLY handle (plans of device handles live on the device)"); return -1; }
  const std::vector<long long>* ptr = nullptr;
  const std::vector<uint8_t>* blob = nullptr;
  if (which_plan == 0) {
    FwdPlanDev* P = nullptr;
    if (ensure_fwd_plan(m, ncomp, &P)) return -1;
    ptr = &P->host.blob_ptr; blob = &P->host.blob;
  } else {
    AdjPlanDev* P = nullptr;
    if (ensure_adj_plan(m, ncomp, &P)) return -1;
    if (!P) { fail("no adjoint tile plan: the mesh has rows longer than 255 entries (the adjoint uses the gather kernel)"); return -1; }
    ptr = &P->host.blob_ptr; blob = &P->host.blob;
  }
  if (array_id == 0) { if (out) memcpy(out, ptr->data(), ptr->size() * sizeof(long long)); return (long long)ptr->size(); }
  if (array_id == 1) { if (out) memcpy(out, blob->data(), blob->size()); return (long long)blob->size(); }
  fail("unknown plan array");
  return -1;
}

int adfem_assemble_csr(adfem_mesh* m, int op, const double* coef, double* vals, void* stream) {
  if (int rc = need_device(m)) return rc;
  if (int rc = check_op(m, op)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int nc = op == ADFEM_OP_STIFFNESS ? m->hm.dim : 1;
  if ((op != ADFEM_OP_STIFFNESS || m->opt_grid_elast) && m->grid_ok) { if (int rc = ensure_pattern(m)) return rc; }      // validates the closed-form row pointers
  if (use_grid_elast(m, op)) return launch_grid_elast(m, false, coef, vals, st);
  if (m->tet_ok && (m->opt_tet_scalar || (m->opt_grid_elast && op == ADFEM_OP_STIFFNESS))) { if (int rc = ensure_pattern(m)) return rc; }    // validates the closed-form rows
  if (use_tet_grid(m, op)) return launch_tet_grid_fwd(m, coef, vals, st);
  if (use_tet_scalar(m, op)) return launch_tet_scalar(m, op, false, coef, vals, st);
  if (op != ADFEM_OP_STIFFNESS && use_grid_any(m))
    return op == ADFEM_OP_LAPLACE ? launch_grid_fwd<OP_LAPLACE>(m, coef, vals, st) : launch_grid_fwd<OP_MASS>(m, coef, vals, st);
  if (op != ADFEM_OP_STIFFNESS && m->opt_row_gather) {
    if (int rc = ensure_pattern(m)) return rc;
    if (m->rg_rows) {
      const int rgt = m->rg_rows;
      const unsigned nb = blocks_for(m->hm.ndof, rgt);
#define CALL_RG(DIM, DEG)                                                                                                                              \
  if (op == ADFEM_OP_LAPLACE)                                                                                                                          \
    k_row_gather_fwd<DIM, DEG, OP_LAPLACE><<<nb, rgt, 0, st>>>(dev_mesh(m, m->opt_area_csr), m->d_adj_ptr.p, m->d_adj_elem.p, m->d_adj_loc.p,   \
                                                                     m->d_rowptr.p, m->d_colind.p, coef, vals);                                        \
  else                                                                                                                                                 \
    k_row_gather_fwd<DIM, DEG, OP_MASS><<<nb, rgt, 0, st>>>(dev_mesh(m, m->opt_area_csr), m->d_adj_ptr.p, m->d_adj_elem.p, m->d_adj_loc.p,      \
                                                                  m->d_rowptr.p, m->d_colind.p, coef, vals)
      DISPATCH_ELEM(m, CALL_RG);
#undef CALL_RG
      CU_TRY(cudaGetLastError());
      return 0;
    }
  }
  if (op == ADFEM_OP_STIFFNESS && m->opt_row_gather && m->hm.degree == 1 && m->hm.g > 1) {
    // P1 elasticity without a tile plan: Gauss-sum pre-pass, then one thread per scalar row (row_gather.cuh)
    if (int rc = ensure_pattern(m)) return rc;
    const size_t smem = (size_t)nc * nc * m->rge_max_entries * sizeof(double);
    if (smem <= SMEM_LIMIT) {
      if (int rc = ensure_presum_buf(m)) return rc;
      if (int rc = launch_presum_coef(dev_mesh(m, m->opt_area_csr), coef_per_gauss(m, op), coef, m->presum_buf.p, st)) return rc;
      const unsigned nb = blocks_for(m->hm.ndof, RGE_THREADS);
      if (m->hm.dim == 2) {
        CU_TRY(cudaFuncSetAttribute(k_row_gather_elast_fwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_row_gather_elast_fwd<2><<<nb, RGE_THREADS, smem, st>>>(dev_mesh(m, m->opt_area_csr), m->d_adj_ptr.p, m->d_adj_elem.p, m->d_adj_loc.p, m->d_rowptr.p,
                                                                 m->d_colind.p, m->pat.nnz, m->presum_buf.p, vals);
      } else {
        CU_TRY(cudaFuncSetAttribute(k_row_gather_elast_fwd<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_row_gather_elast_fwd<3><<<nb, RGE_THREADS, smem, st>>>(dev_mesh(m, m->opt_area_csr), m->d_adj_ptr.p, m->d_adj_elem.p, m->d_adj_loc.p, m->d_rowptr.p,
                                                                 m->d_colind.p, m->pat.nnz, m->presum_buf.p, vals);
      }
      CU_TRY(cudaGetLastError());
      return 0;
    }
  }
  FwdPlanDev* P = nullptr;
  if (int rc = ensure_fwd_plan(m, nc, &P)) return rc;
  const bool presum = use_presum(m, op);
  if (presum) {
    if (int rc = ensure_presum_buf(m)) return rc;
    if (int rc = launch_presum_coef(dev_mesh(m, m->opt_area_csr), coef_per_gauss(m, op), coef, m->presum_buf.p, st)) return rc;
    coef = m->presum_buf.p;
  }
#define CALL_FWD(DIM, DEG)                                                                                   \
  switch (op) {                                                                                              \
    case ADFEM_OP_LAPLACE: return launch_tile_fwd<DIM, DEG, OP_LAPLACE>(m, P, coef, vals, st);               \
    case ADFEM_OP_MASS: return launch_tile_fwd<DIM, DEG, OP_MASS>(m, P, coef, vals, st);                     \
    default: return launch_tile_fwd<DIM, DEG, OP_STIFFNESS>(m, P, coef, vals, st, presum);                   \
  }
  DISPATCH_ELEM(m, CALL_FWD);
#undef CALL_FWD
  return 0;
}

int adfem_assemble_csr_adjoint(adfem_mesh* m, int op, const double* dvals, double* grad_coef, void* stream) {
  if (int rc = need_device(m)) return rc;
  if (int rc = check_op(m, op)) return rc;
  if (int rc = ensure_pattern(m)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (op != ADFEM_OP_STIFFNESS && use_grid_any(m))
    return op == ADFEM_OP_LAPLACE ? launch_grid_adj<OP_LAPLACE>(m, dvals, grad_coef, st) : launch_grid_adj<OP_MASS>(m, dvals, grad_coef, st);
  if (use_grid_elast(m, op)) return launch_grid_elast(m, true, dvals, grad_coef, st);
  if (use_tet_grid(m, op)) return launch_tet_grid_adj(m, dvals, grad_coef, st);
  if (use_tet_scalar(m, op)) return launch_tet_scalar(m, op, true, dvals, grad_coef, st);
  if (use_presum(m, op)) {
    // one gradient block per element from the tile kernel, expanded to the g Gauss points by a streaming pass
    if (int rc = ensure_presum_buf(m)) return rc;
    int rc = 0;
#define CALL_ADJP(DIM, DEG) rc = launch_adj<DIM, DEG, OP_STIFFNESS>(m, dvals, m->presum_buf.p, st, true)
    DISPATCH_ELEM(m, CALL_ADJP);
#undef CALL_ADJP
    if (rc) return rc;
    return launch_expand_grad(dev_mesh(m, m->opt_area_csr), coef_per_gauss(m, op), m->presum_buf.p, grad_coef, st);
  }
#define CALL_ADJ(DIM, DEG)                                                                                   \
  switch (op) {                                                                                              \
    case ADFEM_OP_LAPLACE: return launch_adj<DIM, DEG, OP_LAPLACE>(m, dvals, grad_coef, st);                 \
    case ADFEM_OP_MASS: return launch_adj<DIM, DEG, OP_MASS>(m, dvals, grad_coef, st);                       \
    default: return launch_adj<DIM, DEG, OP_STIFFNESS>(m, dvals, grad_coef, st);                             \
  }
  DISPATCH_ELEM(m, CALL_ADJ);
#undef CALL_ADJ
  return 0;
}

long long adfem_coo_nslots(const adfem_mesh* m, int op) {
  if (!m) return -1;
  const HostMesh& h = m->hm;
  const long long Dt = (op == ADFEM_OP_STIFFNESS ? h.dim : 1) * h.d;
  if (op == ADFEM_OP_MASS && h.dim == 3) return (long long)h.ne * Dt * Dt;      // quirk Q5
  return (long long)h.ne * h.g * Dt * Dt;
}

int adfem_coo_indices(adfem_mesh* m, int op, long long* indices, void* stream) {
  if (int rc = need_device(m)) return rc;
  if (int rc = check_op(m, op)) return rc;
  const int nc = op == ADFEM_OP_STIFFNESS ? m->hm.dim : 1;
  const int per_gauss = !(op == ADFEM_OP_MASS && m->hm.dim == 3);
  const long long total = adfem_coo_nslots(m, op);
  unsigned nb = (unsigned)std::min<long long>((total + 255) / 256, 148 * 32);
  k_coo_indices<<<std::max(1u, nb), 256, 0, (cudaStream_t)stream>>>(m->dm, nc, per_gauss, indices);
  CU_TRY(cudaGetLastError());
  return 0;
}

int adfem_assemble_coo(adfem_mesh* m, int op, const double* coef, double* vv, void* stream) {
  if (int rc = need_device(m)) return rc;
  if (int rc = check_op(m, op)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const HostMesh& h = m->hm;
  const long long G = (long long)h.ne * h.g;
  if (op == ADFEM_OP_MASS && h.dim == 3) {
    const long long n = (long long)h.ne * h.d;
    if (h.degree == 1) k_coo_mass3_fwd<1><<<blocks_for(n, 128), 128, 0, st>>>(dev_mesh(m, m->opt_area_coo), coef, vv);
    else k_coo_mass3_fwd<2><<<blocks_for(n, 128), 128, 0, st>>>(dev_mesh(m, m->opt_area_coo), coef, vv);
    CU_TRY(cudaGetLastError());
    return 0;
  }
#define CALL_COO(DIM, DEG)                                                                                   \
  switch (op) {                                                                                              \
    case ADFEM_OP_LAPLACE: k_coo_scalar_fwd<DIM, DEG, OP_LAPLACE><<<blocks_for(G, 128), 128, 0, st>>>(dev_mesh(m, m->opt_area_coo), coef, vv); break; \
    case ADFEM_OP_MASS: k_coo_scalar_fwd<DIM, DEG, OP_MASS><<<blocks_for(G, 128), 128, 0, st>>>(dev_mesh(m, m->opt_area_coo), coef, vv); break;       \
    default: k_coo_stiff_fwd<DIM, DEG><<<blocks_for(G, 128), 128, 0, st>>>(dev_mesh(m, m->opt_area_coo), coef, vv); break;          \
  }
  DISPATCH_ELEM(m, CALL_COO);
#undef CALL_COO
  CU_TRY(cudaGetLastError());
  return 0;
}

int adfem_assemble_coo_adjoint(adfem_mesh* m, int op, const double* grad_vv, double* grad_coef, void* stream) {
  if (int rc = need_device(m)) return rc;
  if (int rc = check_op(m, op)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const HostMesh& h = m->hm;
  const long long G = (long long)h.ne * h.g;
  if (op == ADFEM_OP_MASS && h.dim == 3) {
    if (h.degree == 1) k_coo_mass3_bwd<1><<<blocks_for(G, 128), 128, 0, st>>>(dev_mesh(m, m->opt_area_coo), grad_vv, grad_coef);
    else k_coo_mass3_bwd<2><<<blocks_for(G, 128), 128, 0, st>>>(dev_mesh(m, m->opt_area_coo), grad_vv, grad_coef);
    CU_TRY(cudaGetLastError());
    return 0;
  }
#define CALL_COOB(DIM, DEG)                                                                                  \
  switch (op) {                                                                                              \
    case ADFEM_OP_LAPLACE: k_coo_scalar_bwd<DIM, DEG, OP_LAPLACE><<<blocks_for(G, 128), 128, 0, st>>>(dev_mesh(m, m->opt_area_coo), grad_vv, grad_coef); break; \
    case ADFEM_OP_MASS: k_coo_scalar_bwd<DIM, DEG, OP_MASS><<<blocks_for(G, 128), 128, 0, st>>>(dev_mesh(m, m->opt_area_coo), grad_vv, grad_coef); break;       \
    default: k_coo_stiff_bwd<DIM, DEG><<<blocks_for(G, 128), 128, 0, st>>>(dev_mesh(m, m->opt_area_coo), grad_vv, grad_coef); break; \
  }
  DISPATCH_ELEM(m, CALL_COOB);
#undef CALL_COOB
  CU_TRY(cudaGetLastError());
  return 0;
}

int adfem_pcl_laplace_jacobian(adfem_mesh* m, double* H, void* stream) {
  if (int rc = need_device(m)) return rc;
  const long long G = (long long)m->hm.ne * m->hm.g;
#define CALL_PCL(DIM, DEG) k_pcl_laplace<DIM, DEG><<<blocks_for(G, 128), 128, 0, (cudaStream_t)stream>>>(dev_mesh(m, m->opt_area_coo), H)
  DISPATCH_ELEM(m, CALL_PCL);
#undef CALL_PCL
  CU_TRY(cudaGetLastError());
  return 0;
}

int adfem_source(adfem_mesh* m, const double* f, double* rhs, void* stream) {
  if (int rc = need_device(m)) return rc;
  if (int rc = ensure_pattern(m)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (use_grid_any(m)) return launch_grid_source(m, false, f, rhs, st);
#define CALL_SRC(DIM, DEG) \
  k_source_fwd<DIM, DEG><<<blocks_for(m->hm.ndof, 128), 128, 0, st>>>(dev_mesh(m, m->opt_area_coo), m->d_adj_ptr.p, m->d_adj_elem.p, m->d_adj_loc.p, f, rhs)
  DISPATCH_ELEM(m, CALL_SRC);
#undef CALL_SRC
  CU_TRY(cudaGetLastError());
  return 0;
}

int adfem_source_adjoint(adfem_mesh* m, const double* grad_rhs, double* grad_f, void* stream) {
  if (int rc = need_device(m)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (m->grid_ok) { if (int rc = ensure_pattern(m)) return rc; }
  if (use_grid_any(m)) return launch_grid_source(m, true, grad_rhs, grad_f, st);
#define CALL_SRCB(DIM, DEG) k_source_bwd<DIM, DEG><<<blocks_for(m->hm.ne, 128), 128, 0, st>>>(dev_mesh(m, m->opt_area_coo), grad_rhs, grad_f)
  DISPATCH_ELEM(m, CALL_SRCB);
#undef CALL_SRCB
  CU_TRY(cudaGetLastError());
  return 0;
}

// ---- Gauss-point operators (gauss_ops.cu) ---------------------------------------------------------------------
namespace {
bool use_tet_gauss(const adfem_mesh* m) { return m->tet_ok && m->opt_structured && !m->host_only && m->hm.degree == 1; }
GridTet grid_tet_of(const adfem_mesh* m) { return GridTet{m->tet_n, m->tet_l, m->tet_xs.p, m->tet_ys.p, m->tet_zs.p, m->d_tet_tab.p}; }
struct GpKind { int basis; bool weighted, scatter_fwd; };     // forward = gather (dof -> Gauss points) unless scatter_fwd
bool gp_kind(int kind, GpKind& k) {
  switch (kind) {
    case ADFEM_GP_FEM_TO_GAUSS: k = {0 /* GB_P1SHAPE */, false, false}; return true;
    case ADFEM_GP_DOF_TO_GAUSS: k = {1 /* GB_SHAPE */, false, false}; return true;
    case ADFEM_GP_GRAD: k = {2 /* GB_GRAD */, false, false}; return true;
    case ADFEM_GP_STRAIN: k = {3 /* GB_STRAIN */, false, false}; return true;
    case ADFEM_GP_STRAIN_ENERGY: k = {3, true, true}; return true;
    default: return false;
  }
}
DofAdjacency dof_adjacency(const adfem_mesh* m) { return DofAdjacency{m->d_adj_ptr.p, m->d_adj_elem.p, m->d_adj_loc.p}; }
// one direction of a Gauss-point operator: to_gauss = dof values -> Gauss points (gather), else the transposed scatter
int gp_apply(adfem_mesh* m, const GpKind& k, bool to_gauss, const double* in, double* out, cudaStream_t st) {
  if (to_gauss) return launch_gp_gather(dev_mesh(m, m->opt_area_coo), m->hm.degree, k.basis, k.weighted, in, out, st);
  if (int rc = ensure_pattern(m)) return rc;
  if (use_tet_gauss(m)) return launch_tet_gp_scatter(dev_mesh(m, m->opt_area_coo), grid_tet_of(m), k.basis, k.weighted, in, out, st);
  if (use_grid_any(m) && m->hm.degree == 1)      // structured triangulation (rectilinear or mapped): index arithmetic instead of the adjacency (grid_gauss.cuh)
    return launch_grid_gp_scatter(dev_mesh(m, m->opt_area_coo), GridTri{m->grid_m, m->grid_n, m->grid_xs.p, m->grid_ys.p}, k.basis, k.weighted, in, out, st,
                                  m->grid_mapped ? m->coords.p : nullptr);
  return launch_gp_scatter(dev_mesh(m, m->opt_area_coo), m->hm.degree, dof_adjacency(m), k.basis, k.weighted, in, out, st);
}
}  // namespace

long long adfem_gauss_op_len(const adfem_mesh* m, int kind, int output) {
  GpKind k;
  if (!m || !gp_kind(kind, k)) { fail(m ? "unknown Gauss-point operator" : "null mesh handle"); return -1; }
  const long long G = (long long)m->hm.ne * m->hm.g, dim = m->hm.dim, ns = dim == 2 ? 3 : 6;
  long long ndofs, ngauss;
  switch (kind) {
    case ADFEM_GP_FEM_TO_GAUSS: ndofs = m->hm.nv; ngauss = G; break;
    case ADFEM_GP_DOF_TO_GAUSS: ndofs = m->hm.ndof; ngauss = G; break;
    case ADFEM_GP_GRAD: ndofs = m->hm.ndof; ngauss = G * dim; break;
    default: ndofs = dim * m->hm.ndof; ngauss = G * ns; break;
  }
  return (output != 0) == k.scatter_fwd ? ndofs : ngauss;
}

int adfem_gauss_op(adfem_mesh* m, int kind, const double* in, double* out, void* stream) {
  if (int rc = need_device(m)) return rc;
  GpKind k;
  if (!gp_kind(kind, k)) return fail("unknown Gauss-point operator");
  return gp_apply(m, k, !k.scatter_fwd, in, out, (cudaStream_t)stream);
}

int adfem_gauss_op_adjoint(adfem_mesh* m, int kind, const double* grad_out, double* grad_in, void* stream) {
  if (int rc = need_device(m)) return rc;
  GpKind k;
  if (!gp_kind(kind, k)) return fail("unknown Gauss-point operator");
  return gp_apply(m, k, k.scatter_fwd, grad_out, grad_in, (cudaStream_t)stream);
}

int adfem_laplace_term(adfem_mesh* m, const double* nu, const double* u, double* out, void* stream) {
  if (int rc = need_device(m)) return rc;
  if (int rc = ensure_pattern(m)) return rc;
  if (use_tet_gauss(m)) return launch_tet_laplace_term(dev_mesh(m, m->opt_area_coo), grid_tet_of(m), nu, u, out, (cudaStream_t)stream);
  if (use_grid_any(m) && m->hm.degree == 1)
    return launch_grid_laplace_term(dev_mesh(m, m->opt_area_coo), GridTri{m->grid_m, m->grid_n, m->grid_xs.p, m->grid_ys.p}, nu, u, out, (cudaStream_t)stream,
                                    m->grid_mapped ? m->coords.p : nullptr);
  return launch_laplace_term(dev_mesh(m, m->opt_area_coo), m->hm.degree, dof_adjacency(m), nu, u, out, (cudaStream_t)stream);
}

int adfem_laplace_term_adjoint(adfem_mesh* m, const double* nu, const double* u, const double* grad_out, double* grad_nu, double* grad_u,
                               void* stream) {
  if (int rc = need_device(m)) return rc;
  if (int rc = ensure_pattern(m)) return rc;
  const DevMesh dm = dev_mesh(m, m->opt_area_coo);
  if (grad_nu) { if (int rc = launch_laplace_term_grad_nu(dm, m->hm.degree, u, grad_out, grad_nu, (cudaStream_t)stream)) return rc; }
  // the term is symmetric in (u, v): d/du of grad_out . K(nu) u is K(nu) grad_out
  if (grad_u) {
    if (use_tet_gauss(m)) {
      if (int rc = launch_tet_laplace_term(dm, grid_tet_of(m), nu, grad_out, grad_u, (cudaStream_t)stream)) return rc;
    } else if (use_grid_any(m) && m->hm.degree == 1) {
      if (int rc = launch_grid_laplace_term(dm, GridTri{m->grid_m, m->grid_n, m->grid_xs.p, m->grid_ys.p}, nu, grad_out, grad_u, (cudaStream_t)stream,
                                            m->grid_mapped ? m->coords.p : nullptr)) return rc;
    } else {
      if (int rc = launch_laplace_term(dm, m->hm.degree, dof_adjacency(m), nu, grad_out, grad_u, (cudaStream_t)stream)) return rc;
    }
  }
  return 0;
}

int adfem_plane_matrix(int mode, long long n, const double* E, const double* nu, double* H, void* stream) {
  if (adfem_device_count() <= 0) return fail("no usable CUDA device (there is no CPU fallback)");
  return launch_plane_matrix(mode, n, E, nu, H, (cudaStream_t)stream);
}
int adfem_plane_matrix_grad(int mode, long long n, const double* E, const double* nu, const double* grad_H, double* grad_E, double* grad_nu,
                            void* stream) {
  if (adfem_device_count() <= 0) return fail("no usable CUDA device (there is no CPU fallback)");
  return launch_plane_matrix_grad(mode, n, E, nu, grad_H, grad_E, grad_nu, (cudaStream_t)stream);
}

// Fused constitutive pre-step + elasticity assembly for P1 triangles: the per-Gauss-point tangent H(E, nu) is never
// written to memory — a streaming pass forms sum_k w_k H_k per element straight from the moduli and the row-tile kernel takes it from there.
int adfem_assemble_csr_plane(adfem_mesh* m, int mode, const double* E, const double* nu, double* vals, void* stream) {
  if (int rc = need_device(m)) return rc;
  if (m->hm.dim != 2 || m->hm.degree != 1) return fail("adfem_assemble_csr_plane: P1 triangles only (use adfem_plane_matrix + adfem_assemble_csr otherwise)");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode != 0 && mode != 1) return fail("plane matrix: mode must be 0 (PlaneStrainMatrix) or 1 (PlaneStressMatrix)");
  if (m->opt_grid_elast && m->grid_ok) { if (int rc = ensure_pattern(m)) return rc; }
  if (use_grid_elast(m, ADFEM_OP_STIFFNESS)) return launch_grid_elast(m, false, nullptr, vals, st, mode, E, nu);
  FwdPlanDev* P = nullptr;
  if (int rc = ensure_fwd_plan(m, 2, &P)) return rc;
  if (int rc = ensure_presum_buf(m)) return rc;
  if (int rc = launch_presum_plane(dev_mesh(m, m->opt_area_csr), mode, E, nu, m->presum_buf.p, st)) return rc;
  return launch_tile_fwd<2, 1, OP_STIFFNESS>(m, P, m->presum_buf.p, vals, st, true);
}
int adfem_assemble_csr_plane_adjoint(adfem_mesh* m, int mode, const double* E, const double* nu, const double* dvals, double* grad_E, double* grad_nu,
                                     void* stream) {
  if (int rc = need_device(m)) return rc;
  if (m->hm.dim != 2 || m->hm.degree != 1) return fail("adfem_assemble_csr_plane_adjoint: P1 triangles only");
  if (int rc = ensure_pattern(m)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode != 0 && mode != 1) return fail("plane matrix: mode must be 0 (PlaneStrainMatrix) or 1 (PlaneStressMatrix)");
  if (use_grid_elast(m, ADFEM_OP_STIFFNESS)) return launch_grid_elast(m, true, dvals, grad_E, st, mode, E, nu, grad_nu);
  if (int rc = ensure_presum_buf(m)) return rc;
  if (int rc = launch_adj<2, 1, OP_STIFFNESS>(m, dvals, m->presum_buf.p, st, true)) return rc;
  return launch_expand_plane_grad(dev_mesh(m, m->opt_area_csr), mode, E, nu, m->presum_buf.p, grad_E, grad_nu, st);
}

int adfem_assemble_csr_host(adfem_mesh* m, int op, const double* coef_host, double* vals_host) {
  if (int rc = need_device(m)) return rc;
  if (int rc = check_op(m, op)) return rc;
  const long long G = (long long)m->hm.ne * m->hm.g, nin = G * coef_per_gauss(m, op);
  const int nc = op == ADFEM_OP_STIFFNESS ? m->hm.dim : 1;
  const long long nout = adfem_csr_nnz(m, nc);
  if (nout < 0) return 1;
  if (m->s_in.n < (size_t)nin) CU_TRY(m->s_in.alloc(nin));
  if (m->s_out.n < (size_t)nout) CU_TRY(m->s_out.alloc(nout));
  if (op != ADFEM_OP_STIFFNESS && use_grid_any(m) && m->opt_host_chunks > 1) {
    // structured mesh: node-row chunks flow through three streams, so the H2D copy of chunk c+1, the kernel of chunk c and the
    // D2H copy of chunk c-1 overlap (PCIe is full duplex).  Chunk c needs the coefficients of cell rows < R[c+1] and writes the
    // contiguous CSR range of node rows [R[c], R[c+1]).
    if (int rc = ensure_host_streams(m)) return rc;
    const int gm = m->grid_m, gn = m->grid_n, C = std::min(m->opt_host_chunks, gn + 1);
    for (int c = 0; c < C; c++) {
      const int ra = (int)((long long)(gn + 1) * c / C), rb = (int)((long long)(gn + 1) * (c + 1) / C);
      const long long k0 = 6LL * gm * std::min(ra, gn), k1 = 6LL * gm * std::min(rb, gn);
      if (k1 > k0) CU_TRY(cudaMemcpyAsync(m->s_in.p + k0, coef_host + k0, (k1 - k0) * sizeof(double), cudaMemcpyHostToDevice, m->hs[0]));
      CU_TRY(cudaEventRecord(m->hev[0], m->hs[0]));
      CU_TRY(cudaStreamWaitEvent(m->hs[1], m->hev[0], 0));
      if (int rc = op == ADFEM_OP_LAPLACE ? launch_grid_fwd<OP_LAPLACE>(m, m->s_in.p, m->s_out.p, m->hs[1], ra, rb)
                                          : launch_grid_fwd<OP_MASS>(m, m->s_in.p, m->s_out.p, m->hs[1], ra, rb)) return rc;
      CU_TRY(cudaEventRecord(m->hev[1], m->hs[1]));
      CU_TRY(cudaStreamWaitEvent(m->hs[2], m->hev[1], 0));
      const long long v0 = grid_rowptr(ra, 0, gm, gn), v1 = rb > gn ? nout : grid_rowptr(rb, 0, gm, gn);
      CU_TRY(cudaMemcpyAsync(vals_host + v0, m->s_out.p + v0, (v1 - v0) * sizeof(double), cudaMemcpyDeviceToHost, m->hs[2]));
    }
    CU_TRY(cudaStreamSynchronize(m->hs[2]));
    return 0;
  }
  CU_TRY(cudaMemcpyAsync(m->s_in.p, coef_host, nin * sizeof(double), cudaMemcpyHostToDevice, 0));
  if (int rc = adfem_assemble_csr(m, op, m->s_in.p, m->s_out.p, nullptr)) return rc;
  CU_TRY(cudaMemcpyAsync(vals_host, m->s_out.p, nout * sizeof(double), cudaMemcpyDeviceToHost, 0));
  CU_TRY(cudaStreamSynchronize(0));
  return 0;
}

int adfem_assemble_csr_adjoint_host(adfem_mesh* m, int op, const double* dvals_host, double* grad_coef_host) {
  if (int rc = need_device(m)) return rc;
  if (int rc = check_op(m, op)) return rc;
  const long long G = (long long)m->hm.ne * m->hm.g, nout = G * coef_per_gauss(m, op);
  const int nc = op == ADFEM_OP_STIFFNESS ? m->hm.dim : 1;
  const long long nin = adfem_csr_nnz(m, nc);
  if (nin < 0) return 1;
  if (m->s_out.n < (size_t)nin) CU_TRY(m->s_out.alloc(nin));
  if (m->s_in.n < (size_t)nout) CU_TRY(m->s_in.alloc(nout));
  if (op != ADFEM_OP_STIFFNESS && use_grid_any(m) && m->opt_host_chunks > 1) {
    // cell-row chunks [C[c], C[c+1]) need the upstream CSR rows of node rows <= C[c+1] and write a contiguous gradient range
    if (int rc = ensure_host_streams(m)) return rc;
    const int gm = m->grid_m, gn = m->grid_n, C = std::min(m->opt_host_chunks, gn);
    long long sent = 0;
    for (int c = 0; c < C; c++) {
      const int ca = (int)((long long)gn * c / C), cb = (int)((long long)gn * (c + 1) / C);
      const long long v1 = cb + 1 > gn ? nin : grid_rowptr(cb + 1, 0, gm, gn);
      if (v1 > sent) CU_TRY(cudaMemcpyAsync(m->s_out.p + sent, dvals_host + sent, (v1 - sent) * sizeof(double), cudaMemcpyHostToDevice, m->hs[0]));
      sent = std::max(sent, v1);
      CU_TRY(cudaEventRecord(m->hev[0], m->hs[0]));
      CU_TRY(cudaStreamWaitEvent(m->hs[1], m->hev[0], 0));
      if (int rc = op == ADFEM_OP_LAPLACE ? launch_grid_adj<OP_LAPLACE>(m, m->s_out.p, m->s_in.p, m->hs[1], ca, cb)
                                          : launch_grid_adj<OP_MASS>(m, m->s_out.p, m->s_in.p, m->hs[1], ca, cb)) return rc;
      CU_TRY(cudaEventRecord(m->hev[1], m->hs[1]));
      CU_TRY(cudaStreamWaitEvent(m->hs[2], m->hev[1], 0));
      const long long g0 = 6LL * gm * ca, g1 = 6LL * gm * cb;
      CU_TRY(cudaMemcpyAsync(grad_coef_host + g0, m->s_in.p + g0, (g1 - g0) * sizeof(double), cudaMemcpyDeviceToHost, m->hs[2]));
    }
    CU_TRY(cudaStreamSynchronize(m->hs[2]));
    return 0;
  }
  CU_TRY(cudaMemcpyAsync(m->s_out.p, dvals_host, nin * sizeof(double), cudaMemcpyHostToDevice, 0));
  if (int rc = adfem_assemble_csr_adjoint(m, op, m->s_out.p, m->s_in.p, nullptr)) return rc;
  CU_TRY(cudaMemcpyAsync(grad_coef_host, m->s_in.p, nout * sizeof(double), cudaMemcpyDeviceToHost, 0));
  CU_TRY(cudaStreamSynchronize(0));
  return 0;
}

}  // extern "C"

// ====================================================================================================
// legacy symbols: global singletons + host pointers (deps/MFEM/API.cpp, deps/MFEM3/API.cpp and the
// *_forward_Julia entry points).  `void` returns like the reference; failures print to stderr.
// ====================================================================================================
namespace {
adfem_mesh* g_mesh2 = nullptr;
adfem_mesh* g_mesh3 = nullptr;

void legacy_fail(const char* where) { fprintf(stderr, "libadfem_cuda: %s failed: %s\n", where, adfem_last_error()); }

long long* legacy_init(adfem_mesh*& slot, int dim, double* vertices, int nv, int* elems, int ne, int order, int lorder, int degree,
                       long long* nedges) {
  if (slot) { printf("WARNING: Internal mesh is being overwritten!\n"); adfem_mesh_destroy(slot); slot = nullptr; }   // API.cpp:6-10
  if (adfem_mesh_create(&slot, dim, vertices, 3, nv, elems, ne, order, degree, lorder, 0)) { legacy_fail("init_nnfem_mesh"); return nullptr; }
  slot->hm.ensure_edges();
  *nedges = slot->hm.nedges;
  long long* edges = (long long*)malloc(sizeof(long long) * 2 * (size_t)std::max<long long>(1, slot->hm.nedges));
  adfem_mesh_edges(slot, edges);
  return edges;
}

// host-pointer COO forward: indices + values
void legacy_coo_fwd(adfem_mesh* m, int op, long long* indices, double* vv, const double* coef, const char* name) {
  if (!m) { fprintf(stderr, "libadfem_cuda: %s called before the mesh was initialised\n", name); return; }
  const long long N = adfem_coo_nslots(m, op), nin = (long long)m->hm.ne * m->hm.g * coef_per_gauss(m, op);
  DevBuf<double> dc, dv; DevBuf<long long> di;
  bool ok = need_device(m) == 0 && dc.alloc(nin) == cudaSuccess && dv.alloc(N) == cudaSuccess &&
            cudaMemcpy(dc.p, coef, nin * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
            adfem_assemble_coo(m, op, dc.p, dv.p, nullptr) == 0 &&
            cudaMemcpy(vv, dv.p, N * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
  if (ok && indices)
    ok = di.alloc(2 * N) == cudaSuccess && adfem_coo_indices(m, op, di.p, nullptr) == 0 &&
         cudaMemcpy(indices, di.p, 2 * N * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess;
  if (!ok) { if (g_err.empty()) g_err = cudaGetErrorString(cudaGetLastError()); legacy_fail(name); }
}
void legacy_coo_bwd(adfem_mesh* m, int op, double* grad_coef, const double* grad_vv, const char* name) {
  if (!m) { fprintf(stderr, "libadfem_cuda: %s called before the mesh was initialised\n", name); return; }
  const long long N = adfem_coo_nslots(m, op), nout = (long long)m->hm.ne * m->hm.g * coef_per_gauss(m, op);
  DevBuf<double> dg, dv;
  bool ok = need_device(m) == 0 && dg.alloc(nout) == cudaSuccess && dv.alloc(N) == cudaSuccess &&
            cudaMemcpy(dv.p, grad_vv, N * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
            adfem_assemble_coo_adjoint(m, op, dv.p, dg.p, nullptr) == 0 &&
            cudaMemcpy(grad_coef, dg.p, nout * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
  if (!ok) { if (g_err.empty()) g_err = cudaGetErrorString(cudaGetLastError()); legacy_fail(name); }
}
void legacy_source_fwd(adfem_mesh* m, double* rhs, const double* f, const char* name) {
  if (!m) { fprintf(stderr, "libadfem_cuda: %s called before the mesh was initialised\n", name); return; }
  const long long G = (long long)m->hm.ne * m->hm.g, n = m->hm.ndof;
  DevBuf<double> df, dr;
  std::vector<double> tmp(n);
  bool ok = need_device(m) == 0 && df.alloc(G) == cudaSuccess && dr.alloc(n) == cudaSuccess &&
            cudaMemcpy(df.p, f, G * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess && adfem_source(m, df.p, dr.p, nullptr) == 0 &&
            cudaMemcpy(tmp.data(), dr.p, n * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
  if (!ok) { if (g_err.empty()) g_err = cudaGetErrorString(cudaGetLastError()); legacy_fail(name); return; }
  for (long long i = 0; i < n; i++) rhs[i] += tmp[i];      // the reference accumulates into a caller-zeroed rhs
}
void legacy_source_bwd(adfem_mesh* m, double* grad_f, const double* grad_rhs, const char* name) {
  if (!m) { fprintf(stderr, "libadfem_cuda: %s called before the mesh was initialised\n", name); return; }
  const long long G = (long long)m->hm.ne * m->hm.g, n = m->hm.ndof;
  DevBuf<double> df, dr;
  bool ok = need_device(m) == 0 && df.alloc(G) == cudaSuccess && dr.alloc(n) == cudaSuccess &&
            cudaMemcpy(dr.p, grad_rhs, n * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
            adfem_source_adjoint(m, dr.p, df.p, nullptr) == 0 &&
            cudaMemcpy(grad_f, df.p, G * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
  if (!ok) { if (g_err.empty()) g_err = cudaGetErrorString(cudaGetLastError()); legacy_fail(name); }
}
// Host-pointer wrapper of a device call with one or two inputs and up to two outputs: copies in, runs `call`, copies out.  `accumulate`
// adds the result into the host array (the reference's `+=` into a caller-zeroed buffer), otherwise it overwrites.
struct HostArg { const double* in; long long n; };
struct HostOut { double* out; long long n; bool accumulate; };
template <class F> void legacy_call(adfem_mesh* m, const char* name, std::initializer_list<HostArg> ins, std::initializer_list<HostOut> outs, F call) {
  if (!m) { fprintf(stderr, "libadfem_cuda: %s called before the mesh was initialised\n", name); return; }
  if (need_device(m)) { legacy_fail(name); return; }
  std::vector<DevBuf<double>> din(ins.size()), dout(outs.size());
  std::vector<const double*> pi; std::vector<double*> po;
  bool ok = true;
  size_t i = 0;
  for (const HostArg& a : ins) {
    ok = ok && din[i].alloc(std::max<long long>(a.n, 1)) == cudaSuccess &&
         (a.n == 0 || cudaMemcpy(din[i].p, a.in, a.n * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess);
    pi.push_back(din[i].p); i++;
  }
  i = 0;
  for (const HostOut& o : outs) { ok = ok && (o.out == nullptr || dout[i].alloc(std::max<long long>(o.n, 1)) == cudaSuccess); po.push_back(o.out ? dout[i].p : nullptr); i++; }
  ok = ok && call(pi, po) == 0 && cudaStreamSynchronize(nullptr) == cudaSuccess;
  i = 0;
  std::vector<double> tmp;
  for (const HostOut& o : outs) {
    if (ok && o.out && o.n > 0) {
      if (o.accumulate) {
        tmp.resize(o.n);
        ok = cudaMemcpy(tmp.data(), dout[i].p, o.n * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
        if (ok) for (long long j = 0; j < o.n; j++) o.out[j] += tmp[j];
      } else {
        ok = cudaMemcpy(o.out, dout[i].p, o.n * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
      }
    }
    i++;
  }
  if (!ok) { if (g_err.empty()) g_err = cudaGetErrorString(cudaGetLastError()); legacy_fail(name); }
}
void legacy_gp(adfem_mesh* m, const char* name, int kind, bool adjoint, const double* in, double* out, bool accumulate) {
  if (!m) { fprintf(stderr, "libadfem_cuda: %s called before the mesh was initialised\n", name); return; }
  const long long nin = adfem_gauss_op_len(m, kind, adjoint ? 1 : 0), nout = adfem_gauss_op_len(m, kind, adjoint ? 0 : 1);
  legacy_call(m, name, {{in, nin}}, {{out, nout, accumulate}}, [&](const std::vector<const double*>& pi, const std::vector<double*>& po) {
    return adjoint ? adfem_gauss_op_adjoint(m, kind, pi[0], po[0], nullptr) : adfem_gauss_op(m, kind, pi[0], po[0], nullptr);
  });
}
void legacy_laplace_term(adfem_mesh* m, const char* name, double* out, const double* nu, const double* u) {
  if (!m) { fprintf(stderr, "libadfem_cuda: %s called before the mesh was initialised\n", name); return; }
  const long long G = (long long)m->hm.ne * m->hm.g, n = m->hm.ndof;
  legacy_call(m, name, {{nu, G}, {u, n}}, {{out, n, true}}, [&](const std::vector<const double*>& pi, const std::vector<double*>& po) {
    return adfem_laplace_term(m, pi[0], pi[1], po[0], nullptr);
  });
}
void legacy_laplace_term_bwd(adfem_mesh* m, const char* name, double* grad_nu, double* grad_u, const double* grad_out, const double* nu, const double* u) {
  if (!m) { fprintf(stderr, "libadfem_cuda: %s called before the mesh was initialised\n", name); return; }
  const long long G = (long long)m->hm.ne * m->hm.g, n = m->hm.ndof;
  legacy_call(m, name, {{nu, G}, {u, n}, {grad_out, n}}, {{grad_nu, G, true}, {grad_u, n, true}},
              [&](const std::vector<const double*>& pi, const std::vector<double*>& po) {
                return adfem_laplace_term_adjoint(m, pi[0], pi[1], pi[2], po[0], po[1], nullptr);
              });
}
void legacy_plane(const char* name, int mode, bool backward, double* o1, double* o2, const double* g, const double* E, const double* nu, int N) {
  DevBuf<double> dE, dnu, dH, d1, d2;
  const size_t n = (size_t)std::max(N, 0);
  bool ok = adfem_device_count() > 0 && dE.alloc(n + 1) == cudaSuccess && dnu.alloc(n + 1) == cudaSuccess && dH.alloc(9 * n + 1) == cudaSuccess &&
            cudaMemcpy(dE.p, E, n * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(dnu.p, nu, n * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess;
  if (ok && !backward) {
    ok = adfem_plane_matrix(mode, N, dE.p, dnu.p, dH.p, nullptr) == 0 && cudaMemcpy(o1, dH.p, 9 * n * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
  } else if (ok) {
    ok = d1.alloc(n + 1) == cudaSuccess && d2.alloc(n + 1) == cudaSuccess && cudaMemcpy(dH.p, g, 9 * n * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
         adfem_plane_matrix_grad(mode, N, dE.p, dnu.p, dH.p, d1.p, d2.p, nullptr) == 0 &&
         cudaMemcpy(o1, d1.p, n * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess &&      // o1 = grad_E
         cudaMemcpy(o2, d2.p, n * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;        // o2 = grad_nu
  }
  if (!ok) { if (g_err.empty()) g_err = cudaGetErrorString(cudaGetLastError()); if (adfem_device_count() <= 0) g_err = "no usable CUDA device (there is no CPU fallback)"; legacy_fail(name); }
}
void legacy_gauss(adfem_mesh* m, double** out) {
  if (!m) return;
  const size_t G = (size_t)m->hm.ne * m->hm.g;
  std::vector<double> xyz(G * m->hm.dim);
  m->hm.gauss_points(xyz.data());
  for (int c = 0; c < m->hm.dim; c++) memcpy(out[c], xyz.data() + c * G, G * sizeof(double));
}
}  // namespace

extern "C" {

long long* init_nnfem_mesh(double* vertices, int num_vertices, int* element_indices, int num_elements, int order, int lorder,
                           int degree, long long* nedges) {
  if (degree == -1) { fprintf(stderr, "libadfem_cuda: BDM1 meshes (degree=-1) are outside the assembly path\n"); return nullptr; }
  return legacy_init(g_mesh2, 2, vertices, num_vertices, element_indices, num_elements, order, lorder, degree, nedges);
}
int mfem_get_ngauss(void) { return g_mesh2 ? g_mesh2->hm.ne * g_mesh2->hm.g : 0; }
void mfem_get_gauss(double* x, double* y) { double* o[2] = {x, y}; legacy_gauss(g_mesh2, o); }
void mfem_get_gauss_weights(double* w) { if (g_mesh2) g_mesh2->hm.gauss_weights(w); }
void mfem_get_area(double* a) { if (g_mesh2) g_mesh2->hm.measure(a); }
int mfem_get_elem_ndof(void) { return g_mesh2 ? g_mesh2->hm.d : 0; }
int mfem_get_ndof(void) { return g_mesh2 ? g_mesh2->hm.ndof : 0; }
void mfem_get_connectivity(long long* conn) { if (g_mesh2) adfem_mesh_connectivity(g_mesh2, conn); }
void mfem_get_element_to_vertices(long long* elems) { if (g_mesh2) adfem_mesh_element_to_vertices(g_mesh2, elems); }
int get_LineIntegralN(void) { double p[64], w[64]; return segment_rule(g_mesh2 ? g_mesh2->hm.lorder : 6, p, w); }
void get_LineIntegralPnW(double* p, double* w) { segment_rule(g_mesh2 ? g_mesh2->hm.lorder : 6, p, w); }

long long* init_nnfem_mesh3(double* vertices, int num_vertices, int* element_indices, int num_elements, int order, int degree,
                            long long* nedges) {
  return legacy_init(g_mesh3, 3, vertices, num_vertices, element_indices, num_elements, order, -1, degree, nedges);
}
int mfem_get_ngauss3(void) { return g_mesh3 ? g_mesh3->hm.ne * g_mesh3->hm.g : 0; }
void mfem_get_gauss3(double* x, double* y, double* z) { double* o[3] = {x, y, z}; legacy_gauss(g_mesh3, o); }
void mfem_get_gauss_weights3(double* w) { if (g_mesh3) g_mesh3->hm.gauss_weights(w); }
void mfem_get_volume3(double* v) { if (g_mesh3) g_mesh3->hm.measure(v); }
int mfem_get_elem_ndof3(void) { return g_mesh3 ? g_mesh3->hm.d : 0; }
int mfem_get_ndof3(void) { return g_mesh3 ? g_mesh3->hm.ndof : 0; }
void mfem_get_connectivity3(long long* conn) { if (g_mesh3) adfem_mesh_connectivity(g_mesh3, conn); }
void mfem_get_element_to_vertices3(long long* elems) { if (g_mesh3) adfem_mesh_element_to_vertices(g_mesh3, elems); }

void FemLaplaceScalar_forward(long long* indices, double* vv, const double* kappa) { legacy_coo_fwd(g_mesh2, ADFEM_OP_LAPLACE, indices, vv, kappa, "FemLaplaceScalar_forward"); }
void FemLaplaceScalar_forward_Julia(long long* indices, double* vv, const double* kappa) { FemLaplaceScalar_forward(indices, vv, kappa); }
// Dense G x (G*d*d) column-major Jacobian into a caller-zeroed HOST array: only the structural entries are written, from the
// COO values of kappa == 1 (FemLaplaceScalar.h:65-92 evaluates exactly N = D D^T w).
void pcl_FemLaplaceScalar_Jacobian(double* H) {
  adfem_mesh* m = g_mesh2;
  if (!m) { fprintf(stderr, "libadfem_cuda: pcl_FemLaplaceScalar_Jacobian called before the mesh was initialised\n"); return; }
  const long long G = (long long)m->hm.ne * m->hm.g, dd = (long long)m->hm.d * m->hm.d, N = G * dd;
  std::vector<double> ones((size_t)G, 1.0), vv((size_t)N);
  DevBuf<double> dc, dv;
  const bool ok = need_device(m) == 0 && dc.alloc(G) == cudaSuccess && dv.alloc(N) == cudaSuccess &&
                  cudaMemcpy(dc.p, ones.data(), G * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
                  adfem_assemble_coo(m, ADFEM_OP_LAPLACE, dc.p, dv.p, nullptr) == 0 &&
                  cudaMemcpy(vv.data(), dv.p, N * sizeof(double), cudaMemcpyDeviceToHost) == cudaSuccess;
  if (!ok) { if (g_err.empty()) g_err = cudaGetErrorString(cudaGetLastError()); legacy_fail("pcl_FemLaplaceScalar_Jacobian"); return; }
  for (long long nz = 0; nz < N; nz++) H[nz / dd + nz * G] = vv[nz];
}
void FemLaplaceScalar_backward(double* grad_kappa, const double* grad_vv, const long long*, const double*, const double*) {
  legacy_coo_bwd(g_mesh2, ADFEM_OP_LAPLACE, grad_kappa, grad_vv, "FemLaplaceScalar_backward");
}
void ComputeFemMassMatrix1_forward(long long* indices, double* vv, const double* rho) { legacy_coo_fwd(g_mesh2, ADFEM_OP_MASS, indices, vv, rho, "ComputeFemMassMatrix1_forward"); }
void ComputeFemMassMatrix1_backward(double* grad_rho, const double* grad_vv, const double*, const double*) {
  legacy_coo_bwd(g_mesh2, ADFEM_OP_MASS, grad_rho, grad_vv, "ComputeFemMassMatrix1_backward");
}
void ComputeFemStiffnessMatrixMfem_forward(long long* indices, double* vv, const double* hmat) {
  legacy_coo_fwd(g_mesh2, ADFEM_OP_STIFFNESS, indices, vv, hmat, "ComputeFemStiffnessMatrixMfem_forward");
}
void ComputeFemStiffnessMatrixMfem_forward_Julia(long long* indices, double* vv, const double* hmat) { ComputeFemStiffnessMatrixMfem_forward(indices, vv, hmat); }
void ComputeFemStiffnessMatrixMfem_backward(double* grad_hmat, const double* grad_vv) {
  legacy_coo_bwd(g_mesh2, ADFEM_OP_STIFFNESS, grad_hmat, grad_vv, "ComputeFemStiffnessMatrixMfem_backward");
}
void FemSourceScalar_forward(double* rhs, const double* f) { legacy_source_fwd(g_mesh2, rhs, f, "FemSourceScalar_forward"); }
void FemSourceScalar_forward_Julia(double* rhs, const double* f) { FemSourceScalar_forward(rhs, f); }
void FemSourceScalar_backward(double* grad_f, const double* grad_rhs, const double*, const double*) { legacy_source_bwd(g_mesh2, grad_f, grad_rhs, "FemSourceScalar_backward"); }

void FemLaplaceScalarT_forward(long long* indices, double* vv, const double* kappa) { legacy_coo_fwd(g_mesh3, ADFEM_OP_LAPLACE, indices, vv, kappa, "FemLaplaceScalarT_forward"); }
void FemLaplaceScalarT_forward_Julia(long long* indices, double* vv, const double* kappa) { FemLaplaceScalarT_forward(indices, vv, kappa); }
void FemLaplaceScalarT_backward(double* grad_kappa, const double* grad_vv, const long long*, const double*, const double*) {
  legacy_coo_bwd(g_mesh3, ADFEM_OP_LAPLACE, grad_kappa, grad_vv, "FemLaplaceScalarT_backward");
}
void ComputeFemMassMatrixMfemT_forward(long long* indices, double* vv, const double* rho) { legacy_coo_fwd(g_mesh3, ADFEM_OP_MASS, indices, vv, rho, "ComputeFemMassMatrixMfemT_forward"); }
void ComputeFemMassMatrixMfemT_backward(double* grad_rho, const double* grad_vv) { legacy_coo_bwd(g_mesh3, ADFEM_OP_MASS, grad_rho, grad_vv, "ComputeFemMassMatrixMfemT_backward"); }
void FemSourceScalarT_forward(double* rhs, const double* f) { legacy_source_fwd(g_mesh3, rhs, f, "FemSourceScalarT_forward"); }
void FemSourceScalarT_forward_Julia(double* rhs, const double* f) { FemSourceScalarT_forward(rhs, f); }
void FemSourceScalarT_backward(double* grad_f, const double* grad_rhs, const double*, const double*) { legacy_source_bwd(g_mesh3, grad_f, grad_rhs, "FemSourceScalarT_backward"); }

// Gauss-point operators and matrix-free terms (gauss_ops.cu)
void FemToGaussPointsMfem_forward(double* out, const double* u) { legacy_gp(g_mesh2, "FemToGaussPointsMfem_forward", ADFEM_GP_FEM_TO_GAUSS, false, u, out, false); }
void FemToGaussPointsMfem_Julia(double* out, const double* u) { FemToGaussPointsMfem_forward(out, u); }
void FemToGaussPointsMfem_backward(double* grad_u, const double* grad_out, const double*, const double*) {
  legacy_gp(g_mesh2, "FemToGaussPointsMfem_backward", ADFEM_GP_FEM_TO_GAUSS, true, grad_out, grad_u, true);
}
void DofToGaussPointsMfem_forward(double* out, const double* u) { legacy_gp(g_mesh2, "DofToGaussPointsMfem_forward", ADFEM_GP_DOF_TO_GAUSS, false, u, out, false); }
void DofToGaussPointsMfem_forward_Julia(double* out, const double* u) { DofToGaussPointsMfem_forward(out, u); }
void DofToGaussPointsMfem_backward(double* grad_u, const double* grad_out, const double*, const double*) {
  legacy_gp(g_mesh2, "DofToGaussPointsMfem_backward", ADFEM_GP_DOF_TO_GAUSS, true, grad_out, grad_u, true);
}
void FemGradMfem_forward(double* out, const double* u) { legacy_gp(g_mesh2, "FemGradMfem_forward", ADFEM_GP_GRAD, false, u, out, false); }
void FemGradMfem_backward(double* grad_u, const double* grad_out, const double*, const double*) {
  legacy_gp(g_mesh2, "FemGradMfem_backward", ADFEM_GP_GRAD, true, grad_out, grad_u, true);
}
void EvalStrainOnGaussPts_forward(double* epsilon, const double* u) { legacy_gp(g_mesh2, "EvalStrainOnGaussPts_forward", ADFEM_GP_STRAIN, false, u, epsilon, true); }
void EvalStrainOnGaussPts_forward_Julia(double* epsilon, const double* u) { EvalStrainOnGaussPts_forward(epsilon, u); }
void EvalStrainOnGaussPts_backward(double* grad_u, const double* grad_epsilon) {
  legacy_gp(g_mesh2, "EvalStrainOnGaussPts_backward", ADFEM_GP_STRAIN, true, grad_epsilon, grad_u, true);
}
void ComputeStrainEnergyTermMfem_forward(double* out, const double* sigma) {
  legacy_gp(g_mesh2, "ComputeStrainEnergyTermMfem_forward", ADFEM_GP_STRAIN_ENERGY, false, sigma, out, true);
}
void ComputeStrainEnergyTermMfem_forward_Julia(double* out, const double* sigma) { ComputeStrainEnergyTermMfem_forward(out, sigma); }
void ComputeStrainEnergyTermMfem_backward(double* grad_sigma, const double* grad_out) {
  legacy_gp(g_mesh2, "ComputeStrainEnergyTermMfem_backward", ADFEM_GP_STRAIN_ENERGY, true, grad_out, grad_sigma, true);
}
void ComputeLaplaceTermMfem_forward(double* out, const double* nu, const double* u) { legacy_laplace_term(g_mesh2, "ComputeLaplaceTermMfem_forward", out, nu, u); }
void ComputeLaplaceTermMfem_forward_Julia(double* out, const double* nu, const double* u) { ComputeLaplaceTermMfem_forward(out, nu, u); }
void ComputeLaplaceTermMfem_backward(double* grad_nu, double* grad_u, const double* grad_out, const double*, const double* nu, const double* u) {
  legacy_laplace_term_bwd(g_mesh2, "ComputeLaplaceTermMfem_backward", grad_nu, grad_u, grad_out, nu, u);
}
void ComputeLaplaceTermMfemT_forward(double* out, const double* nu, const double* u) { legacy_laplace_term(g_mesh3, "ComputeLaplaceTermMfemT_forward", out, nu, u); }
void ComputeLaplaceTermMfem3_forward_Julia(double* out, const double* nu, const double* u) { ComputeLaplaceTermMfemT_forward(out, nu, u); }
void ComputeLaplaceTermMfemT_backward(double* grad_nu, double* grad_u, const double* grad_out, const double*, const double* nu, const double* u) {
  legacy_laplace_term_bwd(g_mesh3, "ComputeLaplaceTermMfemT_backward", grad_nu, grad_u, grad_out, nu, u);
}
void PlaneStrainMatrix_forward(double* out, const double* E, const double* nu, int N) { legacy_plane("PlaneStrainMatrix_forward", 0, false, out, nullptr, nullptr, E, nu, N); }
void PlaneStrainMatrix_backward(double* grad_nu, double* grad_E, const double* grad_out, const double* E, const double* nu, int N) {
  legacy_plane("PlaneStrainMatrix_backward", 0, true, grad_E, grad_nu, grad_out, E, nu, N);
}
void PlaneStressMatrix_forward(double* out, const double* E, const double* nu, int N) { legacy_plane("PlaneStressMatrix_forward", 1, false, out, nullptr, nullptr, E, nu, N); }
void PlaneStressMatrix_backward(double* grad_nu, double* grad_E, const double* grad_out, const double* E, const double* nu, int N) {
  legacy_plane("PlaneStressMatrix_backward", 1, true, grad_E, grad_nu, grad_out, E, nu, N);
}

}  // extern "C"
