// DirichletBd — deps/DirichletBd/DirichletBd.h:8-60 (op "dirichlet_bd", called by fem_impose_Dirichlet_boundary_condition_experimental and
// fem_impose_coupled_Dirichlet_boundary_condition, src/InvCore.jl:6-23): splits the COO triplets of a two-component structured-grid
// operator (dofs u then v, (m+1)(n+1) each) into the free-free part A1 — plus one (b, b, 1.0) per boundary dof — and the free-boundary
// coupling A2 whose column is the position of the boundary dof in the list [bd, bd + (m+1)(n+1)].
//
// The reference walks the slots with std::set / std::map lookups and push_back.  Here: a dof -> list-position table (last duplicate wins, like
// map assignment), two stable stream compactions (exclusive prefix sums over keep flags: input order is kept) and an ascending-dof
// compaction for the unit diagonal (std::set iteration order).  No atomics in values; every output element is written exactly once.
#include <cuda_runtime.h>

#include <algorithm>
#include <cub/device/device_scan.cuh>
#include <string>

#include "../../include/adfem_cuda.h"
#include "internal.h"

using namespace adfem;

namespace {

#define CU_TRY(call)                                                                                   \
  do {                                                                                                 \
    cudaError_t _e = (call);                                                                           \
    if (_e != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(_e));            \
  } while (0)

inline unsigned nblk(long long n) { return (unsigned)((n > 0 ? n : 1) + 255) / 256; }

__global__ void k_bdb_fill(int* p, long long n, int v) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// pos[dof] = 1-based position in bdvec = [bd, bd + off] (DirichletBd.h:27-34); the largest position wins for duplicates
__global__ void k_bdb_map(const int* __restrict__ bd, int bdn, long long off, long long ndof, int* __restrict__ pos, int* __restrict__ err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * bdn) return;
  const long long dof = i < bdn ? (long long)bd[i] : (long long)bd[i - bdn] + off;
  if (dof < 0 || dof >= ndof) { *err = 1; return; }
  atomicMax(&pos[dof], i + 1);
}
__global__ void k_bdb_flags(const long long* __restrict__ ii, const long long* __restrict__ jj, long long N, long long ndof, const int* __restrict__ pos,
                            int* __restrict__ k1, int* __restrict__ k2, int* __restrict__ isbd, int* __restrict__ err) {
  const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s < N) {
    const long long i = ii[s], j = jj[s];
    if (i < 0 || i >= ndof || j < 0 || j >= ndof) { *err = 2; k1[s] = 0; k2[s] = 0; }
    else { const bool bi = pos[i] > 0, bj = pos[j] > 0; k1[s] = (!bi && !bj) ? 1 : 0; k2[s] = (!bi && bj) ? 1 : 0; }      // :36-43
  }
  if (s < ndof) isbd[s] = pos[s] > 0 ? 1 : 0;
}
__global__ void k_bdb_fwd(const long long* __restrict__ ii, const long long* __restrict__ jj, const double* __restrict__ vv, long long N, long long ndof,
                          const int* __restrict__ pos, const int* __restrict__ k1, const int* __restrict__ p1, const int* __restrict__ k2,
                          const int* __restrict__ p2, const int* __restrict__ isbd, const int* __restrict__ pb, long long n1keep,
                          long long* __restrict__ ii1, long long* __restrict__ jj1, double* __restrict__ vv1, long long* __restrict__ ii2,
                          long long* __restrict__ jj2, double* __restrict__ vv2) {
  const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s < N) {
    if (k1[s]) { const long long z = p1[s]; ii1[z] = ii[s]; jj1[z] = jj[s]; vv1[z] = vv[s]; }
    else if (k2[s]) { const long long z = p2[s]; ii2[z] = ii[s]; jj2[z] = pos[jj[s]]; vv2[z] = vv[s]; }
  }
  if (s < ndof && isbd[s]) { const long long z = n1keep + pb[s]; ii1[z] = s; jj1[z] = s; vv1[z] = 1.0; }                  // :44-46, ascending dof
}
__global__ void k_bdb_bwd(long long N, const int* __restrict__ k1, const int* __restrict__ p1, const int* __restrict__ k2, const int* __restrict__ p2,
                          const double* __restrict__ g1, const double* __restrict__ g2, double* __restrict__ grad_vv) {
  const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (s < N) grad_vv[s] = k1[s] ? g1[p1[s]] : (k2[s] ? g2[p2[s]] : 0.0);                                                  // backward(), :96-112
}

struct BdWork {
  int *pos = nullptr, *k1 = nullptr, *k2 = nullptr, *isbd = nullptr, *p1 = nullptr, *p2 = nullptr, *pb = nullptr, *err = nullptr;
  void* tmp = nullptr;
  long long n1keep = 0, n2 = 0, nbd = 0, ndof = 0;
  cudaStream_t st;
  explicit BdWork(cudaStream_t s) : st(s) {}
  ~BdWork() { for (void* p : {(void*)pos, (void*)k1, (void*)k2, (void*)isbd, (void*)p1, (void*)p2, (void*)pb, (void*)err, tmp}) if (p) cudaFreeAsync(p, st); }
};

int bd_prepare(BdWork& W, const long long* ii, const long long* jj, long long N, const int* bd, int bdn, int m, int n) {
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) { cudaGetLastError(); return fail("no CUDA device available (libadfem_cuda has no CPU fallback)"); }
  if (m < 1 || n < 1 || bdn < 0 || N < 0) return fail("DirichletBd: bad sizes");
  const long long off = (long long)(m + 1) * (n + 1);
  // dof space of the callers: 2 (m+1)(n+1) displacement dofs (+ m n pressure dofs in the coupled variant), 0- or 1-based ids
  const long long ndof = 2 * off + (long long)m * n + 2;
  if (N > 2147483647LL - 1 || ndof > 2147483647LL - 1) return fail("DirichletBd: sizes exceed 32-bit");
  W.ndof = ndof;
  cudaStream_t st = W.st;
  for (int** p : {&W.pos, &W.isbd, &W.pb}) CU_TRY(cudaMallocAsync((void**)p, sizeof(int) * (ndof + 1), st));
  for (int** p : {&W.k1, &W.k2, &W.p1, &W.p2}) CU_TRY(cudaMallocAsync((void**)p, sizeof(int) * (N + 1), st));
  CU_TRY(cudaMallocAsync((void**)&W.err, sizeof(int), st));
  CU_TRY(cudaMemsetAsync(W.err, 0, sizeof(int), st));
  k_bdb_fill<<<nblk(ndof + 1), 256, 0, st>>>(W.pos, ndof + 1, 0);
  if (bdn > 0) k_bdb_map<<<nblk(2LL * bdn), 256, 0, st>>>(bd, bdn, off, ndof, W.pos, W.err);
  CU_TRY(cudaMemsetAsync(W.k1 + N, 0, sizeof(int), st));
  CU_TRY(cudaMemsetAsync(W.k2 + N, 0, sizeof(int), st));
  CU_TRY(cudaMemsetAsync(W.isbd + ndof, 0, sizeof(int), st));
  k_bdb_flags<<<nblk(std::max(N, ndof)), 256, 0, st>>>(ii, jj, N, ndof, W.pos, W.k1, W.k2, W.isbd, W.err);
  size_t b1 = 0, b2 = 0;
  CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, b1, W.k1, W.p1, (int)(N + 1), st));
  CU_TRY(cub::DeviceScan::ExclusiveSum(nullptr, b2, W.isbd, W.pb, (int)(ndof + 1), st));
  const size_t tb = std::max(b1, b2);
  CU_TRY(cudaMallocAsync(&W.tmp, tb > 0 ? tb : 16, st));
  CU_TRY(cub::DeviceScan::ExclusiveSum(W.tmp, b1, W.k1, W.p1, (int)(N + 1), st));
  CU_TRY(cub::DeviceScan::ExclusiveSum(W.tmp, b1, W.k2, W.p2, (int)(N + 1), st));
  CU_TRY(cub::DeviceScan::ExclusiveSum(W.tmp, b2, W.isbd, W.pb, (int)(ndof + 1), st));
  int h[4] = {0, 0, 0, 0};
  CU_TRY(cudaMemcpyAsync(&h[0], W.p1 + N, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(&h[1], W.p2 + N, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(&h[2], W.pb + ndof, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaMemcpyAsync(&h[3], W.err, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU_TRY(cudaStreamSynchronize(st));
  if (h[3] == 1) return fail("DirichletBd: boundary dof out of range");
  if (h[3] == 2) return fail("DirichletBd: COO index out of range");
  W.n1keep = h[0]; W.n2 = h[1]; W.nbd = h[2];
  return 0;
}

}  // namespace

extern "C" {

int adfem_dirichlet_bd_count(const long long* ii, const long long* jj, long long N, const int* bd, int bdn, int m, int n, long long* n1, long long* n2,
                             void* stream) {
  BdWork W((cudaStream_t)stream);
  if (int rc = bd_prepare(W, ii, jj, N, bd, bdn, m, n)) return rc;
  *n1 = W.n1keep + W.nbd; *n2 = W.n2;
  return 0;
}

int adfem_dirichlet_bd(const long long* ii, const long long* jj, const double* vv, long long N, const int* bd, int bdn, int m, int n, long long* ii1,
                       long long* jj1, double* vv1, long long* ii2, long long* jj2, double* vv2, void* stream) {
  BdWork W((cudaStream_t)stream);
  if (int rc = bd_prepare(W, ii, jj, N, bd, bdn, m, n)) return rc;
  k_bdb_fwd<<<nblk(std::max(N, W.ndof)), 256, 0, W.st>>>(ii, jj, vv, N, W.ndof, W.pos, W.k1, W.p1, W.k2, W.p2, W.isbd, W.pb, W.n1keep, ii1, jj1, vv1, ii2, jj2, vv2);
  CU_TRY(cudaGetLastError());
  return 0;
}

int adfem_dirichlet_bd_grad(const long long* ii, const long long* jj, long long N, const int* bd, int bdn, int m, int n, const double* grad_vv1,
                            const double* grad_vv2, double* grad_vv, void* stream) {
  BdWork W((cudaStream_t)stream);
  if (int rc = bd_prepare(W, ii, jj, N, bd, bdn, m, n)) return rc;
  if (N > 0) k_bdb_bwd<<<nblk(N), 256, 0, W.st>>>(N, W.k1, W.p1, W.k2, W.p2, grad_vv1, grad_vv2, grad_vv);
  CU_TRY(cudaGetLastError());
  return 0;
}

}  // extern "C"
