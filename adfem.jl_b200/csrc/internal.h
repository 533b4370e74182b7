// Small shared internals of libadfem_cuda.
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace adfem {

extern thread_local std::string g_err;
int fail(const std::string& msg);   // records the message for adfem_last_error() and returns 1

// owning device buffer
template <class T> struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
  cudaError_t alloc(size_t count) {
    release();
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
    if (e == cudaSuccess) n = count; else p = nullptr;
    return e;
  }
};

}  // namespace adfem
