// Mesh-static "symbolic" phase: everything that depends only on the connectivity and is reused by
// every assembly / adjoint call of an optimiser loop.
//
//  * ScalarPattern — the CSR pattern of the scalar operators (set of (dof[p], dof[q]) over elements;
//    this is what Julia's sparse(i, j, v) / TF's sparse ops produce downstream of the reference's COO,
//    src/MFEM/MCore.jl:118-119) plus slot -> nnz map.  Vector (elasticity) operators reuse it: with the
//    reference's component-blocked dof layout (deps/MFEM/ComputeFemStiffnessMatrixMfem/
//    ComputeFemStiffnessMatrixMfem.h:31-34) entry (r + a*n, c + b*n) lives at
//        nc * (a * nnz + rowptr[r]) + b * rowlen(r) + j,        j = position of c in row r.
//  * TilePlan — partition of the rows into spatially compact tiles (Morton order of the dof positions);
//    a CTA computes the local matrices of every element touching its rows into shared memory and then
//    each CSR entry gathers its contributions in a fixed order: no atomics, no global intermediate.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "host_mesh.h"

namespace adfem {

struct ScalarPattern {
  int n = 0;                        // rows
  long long nnz = 0;
  std::vector<long long> rowptr;    // n+1
  std::vector<int> colind;          // nnz, ascending within a row
  std::vector<uint32_t> slot_nnz;   // ne*d*d: CSR position of local entry (e,p,q)
  std::vector<long long> adj_ptr;   // n+1: dof -> incident (element, local index) pairs
  std::vector<int> adj_elem;        // ne*d
  std::vector<uint8_t> adj_loc;     // ne*d
  std::string build(const HostMesh& m, int nthreads);
};

struct TilePlan {
  int ntiles = 0;
  int rows_per_tile = 0;
  int max_rows = 0, max_elems = 0, max_nnz = 0, max_src = 0;
  std::vector<int> row_ptr, rows;           // ntiles+1 ; tile rows (global dof ids, ascending inside a tile)
  std::vector<int> elem_ptr, elems;         // ntiles+1 ; elements a tile evaluates (ascending)
  std::vector<long long> soff_ptr;          // ntiles+1 ; offsets into src_off (tile nnz + 1 entries per tile)
  std::vector<uint16_t> src_off;            // per tile-nnz start into the tile's source list
  std::vector<long long> src_ptr;           // ntiles+1 ; offsets into src
  std::vector<uint16_t> src;                // local_elem * d*d + p*d + q
  double elem_redundancy = 0;               // sum(tile elems) / ne
  std::string build(const HostMesh& m, const ScalarPattern& pat, int rows_per_tile, int max_tile_elems, int nthreads);
};

// adjoint tiles: a CTA owns a compact set of ELEMENTS, stages every CSR row they touch in shared memory
// (coalesced), and each element gathers its d*d upstream gradients from there.
struct AdjTilePlan {
  int ntiles = 0;
  int elems_per_tile = 0;
  int max_rows = 0, max_elems = 0, max_nnz = 0;
  std::vector<int> elem_ptr, elems;         // owned elements
  std::vector<int> row_ptr, rows;           // rows staged by the tile (ascending)
  std::vector<long long> gidx_ptr;          // ntiles+1 ; offsets into gidx (d*d per owned element)
  std::vector<uint16_t> gidx;               // position of slot (e,p,q) inside the tile's staged nnz
  double row_redundancy = 0;
  std::string build(const HostMesh& m, const ScalarPattern& pat, int elems_per_tile, int max_tile_nnz, int nthreads);
};

int default_threads();

}  // namespace adfem
