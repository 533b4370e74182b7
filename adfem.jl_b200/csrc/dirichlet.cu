// Algebraic Dirichlet boundary conditions on a COO matrix — deps/MFEM/ImposeDirichlet/ImposeDirichlet.h:27-93.
//
// The reference walks the COO slots with two std::map lookups per slot, pushes the kept triplets into
// std::vectors and appends one (b, b, 1.0) per boundary dof in ascending dof order.  Here: a dof -> bd-index
// table replaces the map (last duplicate wins, like map assignment), the kept slots and the boundary dofs are
// compacted with exclusive prefix sums (stable: input order kept, ascending dof order for the diagonal, quirk
// Q11).  The right-hand-side correction rhs[i] -= v * u_B[j] and the boundary-value gradient are SEGMENTED REDUCTIONS without atomics
// (SURVEY K9): the slots that couple a free row to a boundary column are compacted, stably radix-sorted by their target (row i, or
// boundary index j), and one thread per segment accumulates in ascending slot order — the reference's own summation order, so the
// result is bit-reproducible run to run.
#include <cuda_runtime.h>

#include <algorithm>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/adfem_cuda.h"
#include "internal.h"

using namespace adfem;

namespace {

#define CU_TRY(call)                                                                                   \
  do {                                                                                                 \
    cudaError_t _e = (call);                                                                           \
    if (_e != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(_e));            \
  } while (0)

__global__ void k_fill_i32(int* p, long long n, int v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
// bmap[dof] = largest index i with bd[i]-1 == dof  (std::map assignment: the last duplicate wins)
__global__ void k_bd_map(const long long* __restrict__ bd, long long bdN, long long N, int* __restrict__ bmap, int* __restrict__ err) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= bdN) return;
  const long long dof = bd[i] - 1;                                   // bd is 1-based (ImposeDirichlet.h:32)
  if (dof < 0 || dof >= N) { *err = 1; return; }
  atomicMax(&bmap[dof], (int)i);
}
__global__ void k_flags(const long long* __restrict__ indices, long long sN, long long N, const int* __restrict__ bmap,
                        int* __restrict__ keep, int* __restrict__ cpl, int* __restrict__ isbd, int* __restrict__ err) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < sN) {
    const long long i = indices[2 * k], j = indices[2 * k + 1];
    if (i < 0 || i >= N || j < 0 || j >= N) { *err = 2; keep[k] = 0; cpl[k] = 0; }
    else { keep[k] = (bmap[i] < 0 && bmap[j] < 0) ? 1 : 0; cpl[k] = (bmap[i] < 0 && bmap[j] >= 0) ? 1 : 0; }     // free row, boundary column
  }
  if (k < N) isbd[k] = bmap[k] >= 0 ? 1 : 0;
}
__global__ void k_copy(const double* __restrict__ a, double* __restrict__ b, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) b[i] = a[i];
}
__global__ void k_dirichlet_fwd(const long long* __restrict__ indices, const double* __restrict__ vv, long long sN, long long N,
                                const int* __restrict__ bmap, const double* __restrict__ bdval, const int* __restrict__ kpos,
                                const int* __restrict__ bpos, long long nkeep, long long* __restrict__ oindices, double* __restrict__ ov,
                                double* __restrict__ orhs) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < sN) {
    const long long i = indices[2 * k], j = indices[2 * k + 1];
    const int bi = bmap[i], bj = bmap[j];
    if (bi < 0 && bj < 0) { const long long z = kpos[k]; oindices[2 * z] = i; oindices[2 * z + 1] = j; ov[z] = vv[k]; }     // :35-39
  }                                                                                                                        // :41-43 -> k_seg_rhs
  if (k < N && bmap[k] >= 0) {                                                                                             // :45-50
    const long long z = nkeep + bpos[k];
    oindices[2 * z] = k; oindices[2 * z + 1] = k; ov[z] = 1.0;
  }
}
__global__ void k_dirichlet_rhs_bd(long long N, const int* __restrict__ bmap, const double* __restrict__ bdval, double* __restrict__ orhs) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < N && bmap[k] >= 0) orhs[k] = bdval[bmap[k]];
}
__global__ void k_dirichlet_bwd(const long long* __restrict__ indices, const double* __restrict__ vv, long long sN, long long N,
                                const int* __restrict__ bmap, const double* __restrict__ bdval, const int* __restrict__ kpos,
                                const double* __restrict__ grad_ov, const double* __restrict__ grad_orhs, double* __restrict__ grad_vv,
                                double* __restrict__ grad_rhs, double* __restrict__ grad_bdval) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < sN) {
    const long long i = indices[2 * k], j = indices[2 * k + 1];
    const int bi = bmap[i], bj = bmap[j];
    double g = 0.0;
    if (bi < 0 && bj < 0) g = grad_ov[kpos[k]];                                                                            // :73-75
    else if (bi < 0 && bj >= 0) g = -bdval[bj] * grad_orhs[i];                                                             // :77-81 (grad_bdval part -> k_seg_gbd)
    grad_vv[k] = g;
  }
  if (k < N) grad_rhs[k] = bmap[k] < 0 ? grad_orhs[k] : 0.0;                                                               // :83-87
}
// coupling slots, compacted in slot order: key = target of the reduction (row i for the rhs, boundary index for grad_bdval), value = slot
__global__ void k_cpl_keys(const long long* __restrict__ indices, long long sN, const int* __restrict__ bmap, const int* __restrict__ cpl,
                           const int* __restrict__ cpos, int by_bd, int* __restrict__ key, int* __restrict__ slot) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= sN || !cpl[k]) return;
  const int z = cpos[k];
  key[z] = by_bd ? bmap[indices[2 * k + 1]] : (int)indices[2 * k];
  slot[z] = (int)k;
}
// one thread per segment of equal keys (stable sort: ascending slot order inside a segment)
__global__ void k_seg_rhs(const int* __restrict__ key, const int* __restrict__ slot, int n, const long long* __restrict__ indices,
                          const double* __restrict__ vv, const int* __restrict__ bmap, const double* __restrict__ bdval, double* __restrict__ orhs) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n || (t > 0 && key[t - 1] == key[t])) return;
  const int i = key[t];
  double acc = orhs[i];                                                         // = rhs[i] (copied before)
  for (int u = t; u < n && key[u] == i; u++) { const int k = slot[u]; acc -= vv[k] * bdval[bmap[indices[2 * (long long)k + 1]]]; }   // :41-43
  orhs[i] = acc;
}
// grad_bdval[b] = sum over coupling slots with boundary index b of -vv * grad_orhs[row]  (slot order)  +  grad_orhs[dof of b]   (:77-81, :88-90)
__global__ void k_gbd_init(const long long* __restrict__ bd, long long bdN, const int* __restrict__ bmap, const double* __restrict__ grad_orhs,
                           double* __restrict__ grad_bdval) {
  const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (b >= bdN) return;
  const long long dof = bd[b] - 1;
  grad_bdval[b] = bmap[dof] == (int)b ? grad_orhs[dof] : 0.0;                   // duplicates in bd: only the last index is the map's value
}
__global__ void k_seg_gbd(const int* __restrict__ key, const int* __restrict__ slot, int n, const long long* __restrict__ indices,
                          const double* __restrict__ vv, const double* __restrict__ grad_orhs, double* __restrict__ grad_bdval) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n || (t > 0 && key[t - 1] == key[t])) return;
  const int b = key[t];
  double acc = 0.0;
  for (int u = t; u < n && key[u] == b; u++) { const int k = slot[u]; acc -= vv[k] * grad_orhs[indices[2 * (long long)k]]; }
  grad_bdval[b] = acc + grad_bdval[b];
}

struct Work {
  int *bmap = nullptr, *keep = nullptr, *isbd = nullptr, *kpos = nullptr, *bpos = nullptr, *err = nullptr, *cpl = nullptr, *cpos = nullptr;
  int *ckey = nullptr, *cslot = nullptr, *ckey2 = nullptr, *cslot2 = nullptr;      // coupling slots before / after the stable sort
  long long ncpl = 0;
  void *tmp = nullptr, *tmp2 = nullptr;
  cudaStream_t st;
  explicit Work(cudaStream_t s) : st(s) {}
  ~Work() {
    for (void* p : {(void*)bmap, (void*)keep, (void*)isbd, (void*)kpos, (void*)bpos, (void*)err, (void*)cpl, (void*)cpos, (void*)ckey, (void*)cslot, (void*)ckey2,
                    (void*)cslot2, tmp, tmp2})
      if (p) cudaFreeAsync(p, st);
  }
};

inline unsigned nblk(long long n) { return (unsigned)((n > 0 ? n : 1) + 255) / 256; }

// builds bmap, keep/isbd flags and their exclusive scans; returns counts
int prepare(Work& W, const long long* indices, long long sN, const long long* bd, long long bdN, long long N, long long* nkeep, long long* nbd) {
  if (sN > 2147483647LL || N > 2147483647LL || bdN > 2147483647LL) return fail("ImposeDirichlet: sizes exceed 32-bit");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) { cudaGetLastError(); return fail("no CUDA device available (libadfem_cuda has no CPU fallback)"); }
  cudaStream_t st = W.st;
  const long long M = sN > N ? sN : N;
  CU_TRY(cudaMallocAsync((void**)&W.bmap, sizeof(int) * (N + 1), st));
  CU_TRY(cudaMallocAsync((void**)&W.keep, sizeof(int) * (sN + 1), st));
  CU_TRY(cudaMallocAsync((void**)&W.isbd, sizeof(int) * (N + 1), st));
  CU_TRY(cudaMallocAsync((void**)&W.kpos, sizeof(int) * (sN + 1), st));
  CU_TRY(cudaMallocAsync((void**)&W.bpos, sizeof(int) * (N + 1), st));
  CU_TRY(cudaMallocAsync((void**)&W.cpl, sizeof(int) * (sN + 1), st));
  CU_TRY(cudaMallocAsync((void**)&W.cpos, sizeof(int) * (sN + 1), st));
  CU_TRY(cudaMallocAsync((void**)&W.err, sizeof(int), st));
  CU_TRY(cudaMemsetAsync(W.err, 0, sizeof(int), st));
  k_fill_i32<<<std::min(nblk(N), 4096u), 256, 0, st>>>(W.bmap, N, -1);
  if (bdN > 0) k_bd_map<<<nblk(bdN), 256, 0, st>>>(bd, bdN, N, W.bmap, W.err);
  CU_TRY(cudaMemsetAsync(W.keep + sN, 0, sizeof(int), st));
  CU_TRY(cudaMemsetAsync(W.cpl + sN, 0, sizeof(int), st));
  CU_TRY(cudaMemsetAsync(W.isbd + N, 0, sizeof(int), st));
  k_flags<<<nblk(M), 256, 0, st>>>(indices, sN, N, W.bmap, W.keep, W.cpl, W.isbd, W.err);
  size_t b1 = 0, b2 = 0;
  CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, b1, W.keep, W.kpos, (int)(sN + 1), st));
  CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, b2, W.isbd, W.bpos, (int)(N + 1), st));
  size_t tb = b1 > b2 ? b1 : b2;
  CU_TRY(cudaMallocAsync(&W.tmp, tb > 0 ? tb : 16, st));
  CU_TRY(cub::DeviceScan::ExclusiveSum(W.tmp, b1, W.keep, W.kpos, (int)(sN + 1), st));
  CU_TRY(cub::DeviceScan::ExclusiveSum(W.tmp, b2, W.isbd, W.bpos, (int)(N + 1), st));
  CU_TRY(cub::DeviceScan::ExclusiveSum(W.tmp, b1, W.cpl, W.cpos, (int)(sN + 1), st));
  int h[3] = {0, 0, 0}, hc = 0;
  CU_TRY(cudaMemcpyAsync(&hc, W.cpos + sN, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(&h[0], W.kpos + sN, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(&h[1], W.bpos + N, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(&h[2], W.err, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  if (h[2] == 1) return fail("ImposeDirichlet: boundary dof out of range");
  if (h[2] == 2) return fail("ImposeDirichlet: COO index out of range");
  *nkeep = h[0]; *nbd = h[1];
  W.ncpl = hc;
  return 0;
}

// compacts the coupling slots and sorts them (stable LSD radix sort) by row (by_bd = 0) or by boundary index (by_bd = 1)
int sort_coupling(Work& W, const long long* indices, long long sN, int by_bd, long long key_range) {
  if (W.ncpl == 0) return 0;
  cudaStream_t st = W.st;
  const int n = (int)W.ncpl;
  for (int** p : {&W.ckey, &W.cslot, &W.ckey2, &W.cslot2})
    if (!*p) CU_TRY(cudaMallocAsync((void**)p, sizeof(int) * n, st));
  k_cpl_keys<<<nblk(sN), 256, 0, st>>>(indices, sN, W.bmap, W.cpl, W.cpos, by_bd, W.ckey, W.cslot);
  int bits = 1;
  while (bits < 31 && (1LL << bits) < key_range) bits++;
  size_t tb = 0;
  CU_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, W.ckey, W.ckey2, W.cslot, W.cslot2, n, 0, bits, st));
  if (W.tmp2) { cudaFreeAsync(W.tmp2, st); W.tmp2 = nullptr; }
  CU_TRY(cudaMallocAsync(&W.tmp2, tb > 0 ? tb : 16, st));
  CU_TRY(cub::DeviceRadixSort::SortPairs(W.tmp2, tb, W.ckey, W.ckey2, W.cslot, W.cslot2, n, 0, bits, st));
  return 0;
}

// J[k + kpos[k]*sN] = 1 for every kept slot k (pcl_ImposeDirichlet, ImposeDirichlet.h:98-112)
__global__ void k_pcl_dirichlet(long long sN, const int* __restrict__ keep, const int* __restrict__ kpos, double* __restrict__ J) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k < sN && keep[k]) J[k + (long long)kpos[k] * sN] = 1.0;
}

}  // namespace

extern "C" {

long long adfem_impose_dirichlet_count(const long long* indices, long long sN, const long long* bd, long long bdN, long long N, void* stream) {
  Work W((cudaStream_t)stream);
  long long nkeep = 0, nbd = 0;
  if (prepare(W, indices, sN, bd, bdN, N, &nkeep, &nbd)) return -1;
  return nkeep + nbd;
}

int adfem_impose_dirichlet(const long long* indices, const double* vv, long long sN, const long long* bd, const double* bdval, long long bdN,
                           const double* rhs, long long N, long long* oindices, double* ov, double* orhs, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  Work W(st);
  long long nkeep = 0, nbd = 0;
  if (int rc = prepare(W, indices, sN, bd, bdN, N, &nkeep, &nbd)) return rc;
  const long long M = sN > N ? sN : N;
  if (N > 0) k_copy<<<nblk(N), 256, 0, st>>>(rhs, orhs, N);
  k_dirichlet_fwd<<<nblk(M), 256, 0, st>>>(indices, vv, sN, N, W.bmap, bdval, W.kpos, W.bpos, nkeep, oindices, ov, orhs);
  if (int rc = sort_coupling(W, indices, sN, 0, N)) return rc;
  if (W.ncpl > 0) k_seg_rhs<<<nblk(W.ncpl), 256, 0, st>>>(W.ckey2, W.cslot2, (int)W.ncpl, indices, vv, W.bmap, bdval, orhs);
  k_dirichlet_rhs_bd<<<nblk(N), 256, 0, st>>>(N, W.bmap, bdval, orhs);
  CU_TRY(cudaGetLastError());
  return 0;
}

int adfem_impose_dirichlet_grad(const double* grad_ov, const double* grad_orhs, const long long* indices, const double* vv, long long sN,
                                const long long* bd, const double* bdval, long long bdN, long long N, double* grad_vv, double* grad_rhs,
                                double* grad_bdval, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  Work W(st);
  long long nkeep = 0, nbd = 0;
  if (int rc = prepare(W, indices, sN, bd, bdN, N, &nkeep, &nbd)) return rc;
  const long long M = sN > N ? sN : N;
  k_dirichlet_bwd<<<nblk(M), 256, 0, st>>>(indices, vv, sN, N, W.bmap, bdval, W.kpos, grad_ov, grad_orhs, grad_vv, grad_rhs, grad_bdval);
  if (bdN > 0) k_gbd_init<<<nblk(bdN), 256, 0, st>>>(bd, bdN, W.bmap, grad_orhs, grad_bdval);
  if (int rc = sort_coupling(W, indices, sN, 1, bdN)) return rc;
  if (W.ncpl > 0) k_seg_gbd<<<nblk(W.ncpl), 256, 0, st>>>(W.ckey2, W.cslot2, (int)W.ncpl, indices, vv, grad_orhs, grad_bdval);
  CU_TRY(cudaGetLastError());
  return 0;
}

int adfem_pcl_impose_dirichlet(const long long* indices, long long sN, const long long* bd, long long bdN, long long N, double* J, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  Work W(st);
  long long nkeep = 0, nbd = 0;
  if (int rc = prepare(W, indices, sN, bd, bdN, N, &nkeep, &nbd)) return rc;
  if (sN > 0) k_pcl_dirichlet<<<nblk(sN), 256, 0, st>>>(sN, W.keep, W.kpos, J);
  CU_TRY(cudaGetLastError());
  return 0;
}

// Legacy symbol (host pointers): J is sN x outdof column-major and caller-zeroed; indices are 1-BASED and column-major
// (indices[k] = row, indices[k + sN] = col), bd 1-based — exactly the reference's arguments (ImposeDirichlet.h:98-112, src/pcl.jl:15-22).
void pcl_ImposeDirichlet(double* J, const long long* indices, const long long* bd, int bdN, int sN) {
  long long N = 0;
  std::vector<long long> ind((size_t)2 * sN);
  for (int k = 0; k < sN; k++) {
    ind[2 * (size_t)k] = indices[k] - 1; ind[2 * (size_t)k + 1] = indices[k + (size_t)sN] - 1;
    N = std::max(N, std::max(indices[k], indices[k + (size_t)sN]));
  }
  for (int i = 0; i < bdN; i++) N = std::max(N, bd[i]);
  long long *d_ind = nullptr, *d_bd = nullptr;
  std::vector<int> keep((size_t)sN + 1), kpos((size_t)sN + 1);
  bool ok = cudaMalloc((void**)&d_ind, sizeof(long long) * std::max(1, 2 * sN)) == cudaSuccess &&
            cudaMalloc((void**)&d_bd, sizeof(long long) * std::max(1, bdN)) == cudaSuccess &&
            cudaMemcpy(d_ind, ind.data(), sizeof(long long) * 2 * sN, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(d_bd, bd, sizeof(long long) * bdN, cudaMemcpyHostToDevice) == cudaSuccess;
  if (ok) {
    Work W(nullptr);
    long long nkeep = 0, nbd = 0;
    ok = prepare(W, d_ind, sN, d_bd, bdN, N, &nkeep, &nbd) == 0 &&
         cudaMemcpy(keep.data(), W.keep, sizeof(int) * sN, cudaMemcpyDeviceToHost) == cudaSuccess &&
         cudaMemcpy(kpos.data(), W.kpos, sizeof(int) * sN, cudaMemcpyDeviceToHost) == cudaSuccess;
  }
  if (d_ind) cudaFree(d_ind);
  if (d_bd) cudaFree(d_bd);
  if (!ok) { fprintf(stderr, "libadfem_cuda: pcl_ImposeDirichlet failed: %s\n", adfem_last_error()); return; }
  for (int k = 0; k < sN; k++) if (keep[k]) J[k + (size_t)kpos[k] * sN] = 1.0;
}

}  // extern "C"
