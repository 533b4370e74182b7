// Mesh-static host tables: what the reference builds in NNFEM_Mesh::init (deps/MFEM/Common.cpp:20-142)
// and NNFEM_Mesh3::init (deps/MFEM3/Common.cpp:9-148), kept as flat arrays instead of one heap object
// per element.  Everything here is computed once per mesh; the per-Gauss-point shape tables (h, hx, hy,
// w) are NOT stored — the kernels recompute them from the vertex coordinates.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "quadrature.h"

namespace adfem {

// Ask for transparent huge pages on freshly reserved, not yet touched storage of a large host array (this image runs THP in "madvise" mode): the
// symbolic phase walks multi-GB arrays at random addresses on renumbered meshes, where 4 KB pages mean a TLB miss per access.  Advice only.
void advise_huge_pages(void* p, size_t bytes);

// fn(begin, end) on equal blocks of [0, n) over `nthreads` host threads (0: ADFEM_HOST_THREADS, else the hardware's); fn(0, n) when n is small
void host_parallel_for(long long n, int nthreads, const std::function<void(long long, long long)>& fn, long long serial_below = 16);

// Struct-of-arrays copy of an element table for the device: out[k * ne + e] = aos[e * kcount + k] (the kernels read index k of 32 consecutive
// elements with one coalesced load).  Element blocks over the host threads, pages of the copy first touched by them.
std::vector<int> soa_copy(const std::vector<int>& aos, long long ne, int kcount);

struct KeyId {
  uint64_t key; int id;
  bool operator<(const KeyId& o) const { return key != o.key ? key < o.key : id < o.id; }
};
// Stable LSD radix sort by key, 11 bits per pass, blocks of the input over the threads (per-thread histograms, exclusive offsets per
// (digit, thread)); passes whose digit is the same for every key are skipped.  With the input in ascending id order equal keys stay in
// ascending id order: the result is the (key, id)-lexicographic order a comparison sort of the pairs gives.
void radix_sort_by_key(std::vector<KeyId>& a, int nthreads);

struct HostMesh {
  int dim = 0;          // 2 (triangles) or 3 (tetrahedra)
  int nv = 0;           // vertices
  int ne = 0;           // elements
  int order = 0, degree = 0, lorder = 0;
  int d = 0;            // dofs per element: 3/6 (tri P1/P2), 4/10 (tet P1/P2)
  int g = 0;            // Gauss points per element
  int ndof = 0;         // nv (P1) or nv + nedges (P2)
  mutable long long nedges = 0;   // P1 meshes: filled by ensure_edges() (the edge list is only an output there, not needed by any kernel)
  mutable bool edges_built = false;
  QuadRule rule;
  std::vector<double> coords;     // nv x dim, packed
  std::vector<int> verts;         // ne x (dim+1) after the orientation fix (det<0 => swap local 0,1)
  std::vector<int> conn;          // ne x d, 0-based dofs: vertices then (P2) nv + edge id, geometry edge order
  mutable std::vector<int> edge_lo, edge_hi;   // edge i joins edge_lo[i] < edge_hi[i] (first-appearance numbering)
  size_t expected_edges() const;     // reservation guess for the edge numbering
  // P1: numbers the edges on first use (nedges, edge_lo, edge_hi); P2 meshes have them from build().  Not thread-safe (setup-time getter).
  void ensure_edges() const;

  // builds all tables; returns "" or an error message
  std::string build(int dim, const double* vertices, int vstride, int nv, const int* elems, int ne, int order,
                    int degree, int lorder);

  // physical position of a dof (vertex, or edge midpoint for P2 edge dofs)
  void dof_position(int dof, double* x) const;

  // setup-time getters (host arithmetic, same formulas as the kernels)
  void gauss_points(double* xyz) const;     // dim blocks of ne*g values (column-major), deps/MFEM/Common.cpp:110-111
  void gauss_weights(double* w) const;      // element-major, deps/MFEM/API.cpp:26-34
  void measure(double* a) const;            // Heron area (Common.cpp:9-15) or tet volume (MFEM3/Common.cpp:88)
};

}  // namespace adfem
