// Multi-GPU interface exchange of the assembly path, inside the library: one process per GPU, element-block partition,
// a dof row touched by several ranks is owned by one of them.  The reference has no parallelism at all, so there
// is no reference interface to mirror; the boundary is the C ABI of include/adfem_cuda.h group (3): the caller (Julia with NCCL.jl / MPI.jl,
// Python with torch.distributed, C++) hands over an `ncclComm_t` (or lets the library create one from an ncclUniqueId) and the mesh-static
// send / receive lists of CSR entry positions; every exchange is then three launches on the caller's stream
//     pack kernel  ->  ONE ncclGroup of ncclSend / ncclRecv over NVLink  ->  unpack kernel
// with no atomics: an owner entry that receives contributions from several ranks sums them in ascending source-rank order in one thread
// (bit-reproducible run to run), unlike an index_add_ scatter.
//
//   adfem_dist_reduce     forward:  partial CSR row segments of non-owned rows travel to their owners and are added there; entries whose
//                                   column the owner does not hold locally land in a small ghost block (PETSc MPIAIJ's off-diagonal part)
//   adfem_dist_replicate  adjoint:  owners send d loss / d K of those entries back, so that the element-level adjoint stays rank-local
//
// Block operators (elasticity, ncomp = dim): the ncomp^2 values of a scalar entry travel together; scalar entry `pos` of row r (start rs,
// length len) holds block element (a, b) at ncomp*(a*nnz + rs) + b*len + (pos - rs), the layout of adfem_assemble_csr.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2, the copy the host process already loaded if any), so single-GPU users of
// libadfem_cuda.so carry no NCCL dependency.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/adfem_cuda.h"
#include "internal.h"

using namespace adfem;

#define CU_TRY(call)                                                                                   \
  do {                                                                                                 \
    cudaError_t _e = (call);                                                                           \
    if (_e != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(_e));            \
  } while (0)

namespace {

// ---- the few NCCL entry points used, bound by name ---------------------------------------------------------------------------------
typedef void* nccl_comm_t;
struct NcclUniqueId { char internal[128]; };
struct Nccl {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(nccl_comm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string err;
};
constexpr int NCCL_FLOAT64 = 8;      // ncclDataType_t::ncclFloat64 (nccl.h)

Nccl* nccl() {
  static Nccl N;
  if (N.lib || !N.err.empty()) return &N;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    N.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (N.lib) break;
  }
  if (!N.lib) { N.err = std::string("NCCL not found (dlopen libnccl.so.2): ") + dlerror(); return &N; }
  auto sym = [&](const char* s) { void* p = dlsym(N.lib, s); if (!p && N.err.empty()) N.err = std::string("NCCL symbol missing: ") + s; return p; };
  N.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
  N.CommInitRank = (int (*)(nccl_comm_t*, int, NcclUniqueId, int))sym("ncclCommInitRank");
  N.CommDestroy = (int (*)(nccl_comm_t))sym("ncclCommDestroy");
  N.GroupStart = (int (*)())sym("ncclGroupStart");
  N.GroupEnd = (int (*)())sym("ncclGroupEnd");
  N.Send = (int (*)(const void*, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclSend");
  N.Recv = (int (*)(void*, size_t, int, int, nccl_comm_t, cudaStream_t))sym("ncclRecv");
  N.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
  return &N;
}
int nccl_fail(const char* what, int rc) {
  Nccl* N = nccl();
  return fail(std::string(what) + ": " + (N->GetErrorString ? N->GetErrorString(rc) : "NCCL error"));
}
#define NCCL_TRY(call)                                   \
  do {                                                   \
    int _r = (call);                                     \
    if (_r != 0) return nccl_fail(#call, _r);            \
  } while (0)

// one exchanged scalar entry: where its block lives in the CSR value array
struct EntryRef { long long rs, pos; int len; };

__device__ __forceinline__ long long block_addr(const EntryRef& e, int nc, long long nnz, int a, int b) {
  return nc * ((long long)a * nnz + e.rs) + (long long)b * e.len + (e.pos - e.rs);
}

// buf[k*nc2 + ab] = vals[entry k, block element ab]
__global__ void k_dist_pack(long long n, int nc, long long nnz, const EntryRef* __restrict__ ent, const double* __restrict__ vals,
                            const double* __restrict__ ghost, const long long* __restrict__ ghost_of, double* __restrict__ buf) {
  const int nc2 = nc * nc;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n * nc2) return;
  const long long k = t / nc2;
  const int ab = (int)(t - k * nc2);
  const EntryRef e = ent[k];
  if (e.pos >= 0) buf[t] = vals[block_addr(e, nc, nnz, ab / nc, ab % nc)];
  else buf[t] = ghost ? ghost[ghost_of[k] * nc2 + ab] : 0.0;          // replicate(): a ghost-column entry returns its own gradient block
}
// vals[entry k, ab] = buf[k*nc2 + ab]   (replicate: every entry position occurs once in the list)
__global__ void k_dist_unpack_copy(long long n, int nc, long long nnz, const EntryRef* __restrict__ ent, const double* __restrict__ buf,
                                   double* __restrict__ vals) {
  const int nc2 = nc * nc;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n * nc2) return;
  const long long k = t / nc2;
  const int ab = (int)(t - k * nc2);
  vals[block_addr(ent[k], nc, nnz, ab / nc, ab % nc)] = buf[t];
}
// reduce: destination d sums the received blocks src[ptr[d] .. ptr[d+1]) in list order (ascending source rank) and adds the sum to its entry
__global__ void k_dist_unpack_sum(long long ndest, int nc, long long nnz, const EntryRef* __restrict__ dest, const long long* __restrict__ ptr,
                                  const long long* __restrict__ src, const double* __restrict__ buf, double* __restrict__ vals) {
  const int nc2 = nc * nc;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= ndest * nc2) return;
  const long long d = t / nc2;
  const int ab = (int)(t - d * nc2);
  double s = 0.0;
  for (long long i = ptr[d]; i < ptr[d + 1]; i++) s += buf[src[i] * nc2 + ab];
  const long long at = block_addr(dest[d], nc, nnz, ab / nc, ab % nc);
  vals[at] += s;
}
// ghost block: one value block per received entry without a local column
__global__ void k_dist_unpack_ghost(long long ng, int nc2, const long long* __restrict__ src, const double* __restrict__ buf, double* __restrict__ ghost) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= ng * nc2) return;
  ghost[t] = buf[src[t / nc2] * nc2 + t % nc2];
}

template <class T> cudaError_t upload_vec(DevBuf<T>& d, const std::vector<T>& h) {
  cudaError_t e = d.alloc(h.size());
  if (e != cudaSuccess || h.empty()) return e;
  return cudaMemcpy(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}
inline unsigned nblk(long long n) { return (unsigned)((n + 255) / 256); }

}  // namespace

struct adfem_dist {
  nccl_comm_t comm = nullptr;
  int rank = 0, world = 1, max_nc = 1;
  long long nnz = 0, nsend = 0, nrecv = 0, ndest = 0, nghost = 0;
  std::vector<long long> send_counts, recv_counts;     // scalar entries per peer
  DevBuf<EntryRef> d_send, d_recv, d_dest;
  DevBuf<long long> d_dest_ptr, d_dest_src, d_ghost_src, d_ghost_of;
  DevBuf<double> sendbuf, recvbuf;                      // (nsend | nrecv) * max_nc^2 doubles each, both directions reuse them
};

extern "C" {

int adfem_dist_nccl_unique_id(void* id128) {
  Nccl* N = nccl();
  if (!N->err.empty()) return fail(N->err);
  NcclUniqueId id;
  NCCL_TRY(N->GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int adfem_dist_comm_create(void** comm, const void* id128, int rank, int world) {
  Nccl* N = nccl();
  if (!N->err.empty()) return fail(N->err);
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  nccl_comm_t c = nullptr;
  NCCL_TRY(N->CommInitRank(&c, world, id, rank));
  *comm = c;
  return 0;
}

int adfem_dist_comm_destroy(void* comm) {
  Nccl* N = nccl();
  if (!N->err.empty()) return fail(N->err);
  if (comm) NCCL_TRY(N->CommDestroy((nccl_comm_t)comm));
  return 0;
}

int adfem_dist_create(adfem_dist** out, adfem_mesh* m, void* nccl_comm, int rank, int world, int max_ncomp, const long long* send_counts,
                      const long long* send_pos, const long long* recv_counts, const long long* recv_pos) {
  if (!out || !m) return fail("adfem_dist_create: null argument");
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return fail("adfem_dist_create: bad rank / world");
  if (world > 1 && !nccl_comm) return fail("adfem_dist_create: world > 1 needs an ncclComm_t (adfem_dist_comm_create makes one)");
  if (max_ncomp < 1 || max_ncomp > 3) return fail("adfem_dist_create: max_ncomp must be 1, 2 or 3");
  const long long nnz = adfem_csr_nnz(m, 1);
  if (nnz < 0) return 1;
  const long long n = adfem_mesh_info(m, ADFEM_INFO_NDOF);
  std::vector<long long> rowptr((size_t)n + 1);
  std::vector<int> colind((size_t)nnz);
  if (int rc = adfem_csr_pattern(m, 1, rowptr.data(), colind.data())) return rc;
  std::vector<int>().swap(colind);
  auto D = std::make_unique<adfem_dist>();
  D->comm = (nccl_comm_t)nccl_comm; D->rank = rank; D->world = world; D->max_nc = max_ncomp; D->nnz = nnz;
  D->send_counts.assign(send_counts, send_counts + world);
  D->recv_counts.assign(recv_counts, recv_counts + world);
  D->nsend = std::accumulate(D->send_counts.begin(), D->send_counts.end(), 0LL);
  D->nrecv = std::accumulate(D->recv_counts.begin(), D->recv_counts.end(), 0LL);
  if (D->send_counts[rank] != 0 || D->recv_counts[rank] != 0) return fail("adfem_dist_create: a rank does not exchange with itself");
  auto ref_of = [&](long long pos, EntryRef& e) -> bool {
    if (pos < 0 || pos >= nnz) return false;
    const long long r = (long long)(std::upper_bound(rowptr.begin(), rowptr.end(), pos) - rowptr.begin()) - 1;
    e.rs = rowptr[r]; e.len = (int)(rowptr[r + 1] - rowptr[r]); e.pos = pos;
    return true;
  };
  std::vector<EntryRef> hs((size_t)D->nsend), hr((size_t)D->nrecv);
  for (long long k = 0; k < D->nsend; k++)
    if (!ref_of(send_pos[k], hs[k])) return fail("adfem_dist_create: send position out of range");
  // received entries: matched ones grouped by destination position (stable: ascending receive index = ascending source rank)
  std::vector<long long> ghost_src, ghost_of((size_t)D->nrecv, -1), order;
  for (long long k = 0; k < D->nrecv; k++) {
    if (recv_pos[k] < 0) { hr[k] = EntryRef{0, -1, 0}; ghost_of[k] = (long long)ghost_src.size(); ghost_src.push_back(k); continue; }
    if (!ref_of(recv_pos[k], hr[k])) return fail("adfem_dist_create: receive position out of range");
    order.push_back(k);
  }
  std::stable_sort(order.begin(), order.end(), [&](long long a, long long b) { return recv_pos[a] < recv_pos[b]; });
  std::vector<EntryRef> hd;
  std::vector<long long> dptr(1, 0);
  for (size_t i = 0; i < order.size(); i++) {
    if (i == 0 || recv_pos[order[i]] != recv_pos[order[i - 1]]) { if (i) dptr.push_back((long long)i); hd.push_back(hr[order[i]]); }
  }
  if (!order.empty()) dptr.push_back((long long)order.size());
  D->ndest = (long long)hd.size();
  D->nghost = (long long)ghost_src.size();
  CU_TRY(upload_vec(D->d_send, hs));
  CU_TRY(upload_vec(D->d_recv, hr));
  CU_TRY(upload_vec(D->d_dest, hd));
  CU_TRY(upload_vec(D->d_dest_ptr, dptr));
  CU_TRY(upload_vec(D->d_dest_src, order));
  CU_TRY(upload_vec(D->d_ghost_src, ghost_src));
  CU_TRY(upload_vec(D->d_ghost_of, ghost_of));
  const size_t nc2 = (size_t)max_ncomp * max_ncomp;
  CU_TRY(D->sendbuf.alloc(std::max<size_t>(1, (size_t)std::max(D->nsend, D->nrecv) * nc2)));
  CU_TRY(D->recvbuf.alloc(std::max<size_t>(1, (size_t)std::max(D->nsend, D->nrecv) * nc2)));
  *out = D.release();
  return 0;
}

void adfem_dist_destroy(adfem_dist* D) { delete D; }

long long adfem_dist_info(const adfem_dist* D, int what) {
  if (!D) return -1;
  switch (what) {
    case 0: return D->nsend;
    case 1: return D->nrecv;
    case 2: return D->nghost;
    case 3: return D->ndest;
    case 4: return 8 * (D->nsend + D->nrecv);       // interface bytes per exchange and rank, scalar operator
    default: return -1;
  }
}

// one grouped exchange: this rank sends `scount[q]*w` doubles to q and receives `rcount[q]*w` from q
static int exchange(adfem_dist* D, const std::vector<long long>& scount, const std::vector<long long>& rcount, int w, const double* sbuf, double* rbuf,
                    cudaStream_t st) {
  Nccl* N = nccl();
  if (!N->err.empty()) return fail(N->err);
  NCCL_TRY(N->GroupStart());
  long long so = 0, ro = 0;
  for (int q = 0; q < D->world; q++) {
    if (scount[q]) NCCL_TRY(N->Send(sbuf + so * w, (size_t)scount[q] * w, NCCL_FLOAT64, q, D->comm, st));
    if (rcount[q]) NCCL_TRY(N->Recv(rbuf + ro * w, (size_t)rcount[q] * w, NCCL_FLOAT64, q, D->comm, st));
    so += scount[q]; ro += rcount[q];
  }
  NCCL_TRY(N->GroupEnd());
  return 0;
}

int adfem_dist_reduce(adfem_dist* D, int ncomp, double* vals, double* ghost_vals, void* stream) {
  if (!D) return fail("null exchange handle");
  if (ncomp < 1 || ncomp > D->max_nc) return fail("adfem_dist_reduce: ncomp exceeds the max_ncomp the handle was created with");
  if (D->world == 1) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int nc2 = ncomp * ncomp;
  if (D->nsend) k_dist_pack<<<nblk(D->nsend * nc2), 256, 0, st>>>(D->nsend, ncomp, D->nnz, D->d_send.p, vals, nullptr, nullptr, D->sendbuf.p);
  if (int rc = exchange(D, D->send_counts, D->recv_counts, nc2, D->sendbuf.p, D->recvbuf.p, st)) return rc;
  if (D->ndest) k_dist_unpack_sum<<<nblk(D->ndest * nc2), 256, 0, st>>>(D->ndest, ncomp, D->nnz, D->d_dest.p, D->d_dest_ptr.p, D->d_dest_src.p, D->recvbuf.p, vals);
  if (D->nghost && ghost_vals) k_dist_unpack_ghost<<<nblk(D->nghost * nc2), 256, 0, st>>>(D->nghost, nc2, D->d_ghost_src.p, D->recvbuf.p, ghost_vals);
  CU_TRY(cudaGetLastError());
  return 0;
}

int adfem_dist_replicate(adfem_dist* D, int ncomp, double* dvals, const double* dghost, void* stream) {
  if (!D) return fail("null exchange handle");
  if (ncomp < 1 || ncomp > D->max_nc) return fail("adfem_dist_replicate: ncomp exceeds the max_ncomp the handle was created with");
  if (D->world == 1) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int nc2 = ncomp * ncomp;
  // owners pack what they received positions for (recvbuf side), contributors get it back in their send order
  if (D->nrecv) k_dist_pack<<<nblk(D->nrecv * nc2), 256, 0, st>>>(D->nrecv, ncomp, D->nnz, D->d_recv.p, dvals, dghost, D->d_ghost_of.p, D->recvbuf.p);
  if (int rc = exchange(D, D->recv_counts, D->send_counts, nc2, D->recvbuf.p, D->sendbuf.p, st)) return rc;
  if (D->nsend) k_dist_unpack_copy<<<nblk(D->nsend * nc2), 256, 0, st>>>(D->nsend, ncomp, D->nnz, D->d_send.p, D->sendbuf.p, dvals);
  CU_TRY(cudaGetLastError());
  return 0;
}

}  // extern "C"
