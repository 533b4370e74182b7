// Index arithmetic of the reference's structured triangulation `Mesh(m, n, h)` version 1 (src/MFEM/MFEM.jl:134-146): node (i, j) = i*(m+1)+j,
// cell (ci, cj) = elements 2*(ci*m+cj) = T0 = [BL, BR, TL] and 2*(ci*m+cj)+1 = T1 = [TL, BR, TR]; closed-form CSR row pointers of the
// 7-point pattern.  Shared by tri_grid.cuh (scalar operators) and grid_elast.cuh (elasticity); host + device.
#pragma once

namespace adfem {

struct GridTri {
  int m, n;                    // cells in x and y
  const double* xs;            // m+1 node abscissae
  const double* ys;            // n+1 node ordinates
};

// CSR row pointer of node (i, j), j in [0, m+1] (j = m+1: end of node row i), closed form for the 7-point pattern
//   row = [ (i-1,j), (i-1,j+1), (i,j-1), (i,j), (i,j+1), (i+1,j-1), (i+1,j) ]  restricted to existing nodes
__host__ __device__ __forceinline__ long long grid_row_prefix(int j, int m, int A, int B) {
  const int jm = j < m ? j : m, j1 = j > 0 ? j - 1 : 0;
  return (long long)j * (1 + A + B) + (long long)(A + 1) * jm + (long long)(1 + B) * j1;
}
__host__ __device__ __forceinline__ long long grid_rowptr(int i, int j, int m, int n) {
  const int A = i > 0, B = i < n;
  long long before = 0;
  if (i > 0) {
    before = grid_row_prefix(m + 1, m, 0, n > 0);                                   // node row 0
    if (i > 1) before += (long long)(i - 1) * grid_row_prefix(m + 1, m, 1, 1);      // node rows 1 .. i-1 (all have a row above and below)
  }
  return before + grid_row_prefix(j, m, A, B);
}

}  // namespace adfem
