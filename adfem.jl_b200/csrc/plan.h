// Mesh-static "symbolic" phase: everything that depends only on the connectivity and is reused by
// every assembly / adjoint call of an optimiser loop.
//
//  * ScalarPattern — the CSR pattern of the scalar operators (set of (dof[p], dof[q]) over elements;
//    this is what Julia's sparse(i, j, v) / TF's sparse ops produce downstream of the reference's COO,
//    src/MFEM/MCore.jl:118-119) plus slot -> nnz map.  Vector (elasticity) operators reuse it: with the
//    reference's component-blocked dof layout (deps/MFEM/ComputeFemStiffnessMatrixMfem/
//    ComputeFemStiffnessMatrixMfem.h:31-34) entry (r + a*n, c + b*n) lives at
//        nc * (a * nnz + rowptr[r]) + b * rowlen(r) + j,        j = position of c in row r.
//  * FwdTiles — partition of the rows into spatially compact tiles (Morton order of the dof positions).
//    A CTA evaluates the local matrices of every element touching its rows into shared memory, then
//    each CSR entry gathers its contributions in a fixed order: no atomics, no global intermediate.
//  * AdjTiles — partition of the ELEMENTS into compact tiles; a CTA stages the CSR rows its elements
//    touch (coalesced) and each element gathers its upstream gradients from shared memory.
//
// Every tile's mesh-static data (index lists AND the coordinates of the vertices it needs) is packed
// into one contiguous, 16-byte aligned blob that TMA bulk copies bring on chip (two copies: head, body).
#pragma once
#include <cstdint>
#include <functional>
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
  bool repeated_dofs = false;       // some element lists a dof twice (a degenerate element): a CSR entry can then receive more than one
                                    // contribution per incident (element, local dof) pair, which the forward tile builder sizes its lists for
  std::string build(const HostMesh& m, int nthreads);
  // The same tables for the structured triangulation Mesh(gm, gn, h), version 1, P1 (connectivity as detect_tri_grid of adfem_cuda.cu verifies
  // it: cell (ci, cj) with first node a = ci (gm+1) + cj holds the triangles (a, a+1, a+gm+1) and (a+gm+1, a+1, a+gm+2)), from index arithmetic:
  // no counting sort, no per-row sets, no searches in the connectivity.  Byte-identical to build() (tests/test_structured.py).
  std::string build_tri_grid(const HostMesh& m, int gm, int gn, int nthreads);
  // The same for the structured tetrahedral grid Mesh3(gn, gn, gl, h), P1 (element 5 * cube + t holds the vertex SET of tetrahedron t of the cube's
  // splitting, as detect_tet_grid verifies; the local position of a vertex inside its element is read from the mesh, the orientation fix may have
  // swapped the first two): incident tetrahedra and row entries of a node from the two parity neighbourhoods of tet_grid_tables.h.
  std::string build_tet_grid(const HostMesh& m, int gn, int gl, int nthreads);
};

inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// Every tile blob is split into a small HEAD (what the kernel needs one pipeline stage early: the element ids whose
// coefficients are prefetched / the row segments whose upstream gradients are staged) and a BODY (everything else).
// Head and body of tile t are contiguous: blob_ptr[2t] = head offset, blob_ptr[2t+1] = body offset, blob_ptr[2t+2] = end.
// Every section starts on a 16-byte boundary (TMA bulk copy alignment rule).
//
// ---- forward tile blob -----------------------------------------------------------------------------
// HEAD
//   hdr    int32[8]            {nrows, nel, nvt, ndst, nsrc, ncls, flags (bit0: 32-bit destinations), nnz}
//   elems  int32[nel]          global element ids evaluated by the tile (ascending) — coefficient index
// BODY
//   rstart uint32[nrows]       CSR offset of the first entry of each tile row (scalar pattern)
//   rlen   uint16[nrows]       row lengths — vector plans (sym = 0) only
//   tv     uint16[nvl*nel]     tile-local vertex ids, k-major (tv[k*nel+le]), post orientation fix
//   xy     double[dim*nvt]     coordinates of the tile-local vertices
//   cls    int32[4*ncls]       gather classes {source count | paired << 16, items, first source, first destination}: items
//                              with the same number of contributions are processed together so that every lane of a warp
//                              runs the same fully unrolled gather (no divergence).  A PAIRED item (scalar plans only) is an
//                              off-diagonal pair (r,c),(c,r) with both rows in the tile: the local matrices are symmetric, so
//                              the sum is formed once and stored to both destinations (dst[first+i], dst[first+items+i]).
//   dst    uint16|uint32[...]  destinations: tile row | position-in-row << 8 (or << 16)
//   src    uint16[nsrc]        class-major, k-major inside a class: src[first + k*items + i].  Scalar plans (sym=1):
//                              shared-memory index sym(p,q)*nel + le of a local-matrix value (packed upper triangle);
//                              vector plans: le*d*d + p*d + q
struct FwdTiles {
  int ntiles = 0, rows_per_tile = 0, sym = 1, ent32 = 0;
  int max_rows = 0, max_elems = 0, max_nnz = 0, max_src = 0, max_verts = 0;
  size_t max_head = 0, max_body = 0;
  std::vector<long long> blob_ptr;    // 2*ntiles+1 byte offsets
  std::vector<uint8_t> blob;
  double elem_redundancy = 0;
  // optional, set by the caller before build(): true when a plan whose largest head / body / element count / entry count are these cannot be
  // launched (shared-memory budget).  build() then gives up with "tile too large" at the first tile that shows it instead of finishing a plan
  // the caller is going to reject (the caller's tile-size search shrinks the tiles and builds again).
  std::function<bool(size_t max_head, size_t max_body, int max_elems, int max_nnz)> too_big;
  std::vector<int> morton_cache;      // rows in Morton order: kept from a rejected build for the next try of the tile-size search, dropped on success
  std::string build(const HostMesh& m, const ScalarPattern& pat, int rows_per_tile, int max_tile_elems, int sym, int nthreads);
};

// ---- adjoint tile blob -----------------------------------------------------------------------------
// HEAD
//   hdr    int32[8]            {nrows, nel, nvt, nnz, flags (bit0: 16-bit lrow, bit1: td present), 0, 0, 0}
//   delta  uint32[nrows]       CSR offset of each staged row minus its offset in the staging buffer (mod 2^32): staged entry i
//                              of that row is global entry i + delta
//   roff   uint16[nrows+1]     offset of each staged row inside the staging buffer
//   lrow   uint8|uint16[nnz]   staged row of every staged entry
// BODY
//   elems  int32[nel]          owned elements (ascending)
//   tv     uint16[nvl*nel]
//   xy     double[dim*nvt]
//   td     uint16[d*nel]       staged row of each local dof, k-major — only when it differs from tv (P2); for P1 the staged
//                              rows ARE the tile vertices in the same (ascending) order
//   gpk    uint32[d*W*nel]     W = ceil(d/4); word (p*W + q/4)*nel + le, byte q%4: position of slot (le,p,q) inside its CSR row;
//                              the staged entry is roff[td_p] + position
struct AdjTiles {
  int ntiles = 0, elems_per_tile = 0;
  int max_rows = 0, max_elems = 0, max_nnz = 0, max_verts = 0;
  size_t max_head = 0, max_body = 0;
  std::vector<long long> blob_ptr;    // 2*ntiles+1
  std::vector<uint8_t> blob;
  double row_redundancy = 0;
  std::function<bool(size_t max_head, size_t max_body, int max_elems, int max_nnz)> too_big;      // as in FwdTiles
  std::vector<int> morton_cache;      // elements in Morton order, as in FwdTiles
  std::string build(const HostMesh& m, const ScalarPattern& pat, int elems_per_tile, int max_tile_nnz, int nthreads);
};

int default_threads();

}  // namespace adfem
