// tests/host_emul/host_check.cpp — TEST INFRASTRUCTURE ONLY.  Host helpers of adfem.jl_b200/csrc/host_mesh.cpp that the library only calls on its
// device path (adfem_mesh_create with a GPU), exported so that tests/test_host_plan.py can check them on a machine without one.  Built by the test
// with g++ together with host_mesh.cpp; never linked into libadfem_cuda.so.
#include <cstring>

#include "host_mesh.h"

extern "C" int check_soa_copy(const int* aos, long long ne, int kcount, int* out) {
  const std::vector<int> in(aos, aos + (size_t)ne * kcount);
  const std::vector<int> got = adfem::soa_copy(in, ne, kcount);
  if (got.size() != (size_t)ne * kcount) return 1;
  if (!got.empty()) memcpy(out, got.data(), got.size() * sizeof(int));
  return 0;
}
