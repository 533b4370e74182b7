// Mesh-independent tables of the reference's structured tetrahedral grid `Mesh3(n, n, l, h)` (src/MFEM3/MFEM.jl:124-185): every cube is cut into
// 5 tetrahedra, `TE1` for cubes whose 1-based index sum is even and `TE2` otherwise (MFEM.jl:131-144), so a node sees one of two neighbourhoods,
// selected by the parity of i + j + k.  The tables list, for each parity, the tetrahedra around a node in ascending element order, where the node
// and its neighbours sit in them, and which tetrahedra feed each of the 27 possible row entries.  Built once on the host (and by the host
// emulation harness); a few KB.
#pragma once
#include <cstring>

namespace adfem {

struct TetGridTables {
  int ninc[2];              // incident tetrahedra of an interior node of parity 0 / 1
  int inc[2][32][5];        // cube origin relative to the node (oi, oj, ok in {-1, 0}), tetrahedron of the cube (0..4), local index of the node
  int voff[2][32][4][3];    // offsets (di, dj, dk) of the four vertices relative to the node
  int vslot[2][32][4];      // their row slots: slot = (dk+1)*9 + (dj+1)*3 + (di+1), i.e. ascending node id
  int present[2];           // 27-bit masks of the slots that occur for the parity
  int te[2][5][4];          // local cube vertices of the 5 tetrahedra: te[1] = TE1 (1-based cube index sum even), te[0] = TE2
  int nsrc[2][27];          // sources of a slot: (incident tetrahedron, local vertex) pairs in ascending element order
  int src[2][27][32][2];
};

// local vertices 0..7 of a cube at offsets (v & 1, (v >> 1) & 1, v >> 2); tetrahedra of the two splittings, 0-based (MFEM.jl:131-144)
inline const int (*tet_grid_split(int even_one_based))[4] {
  static const int TE1[5][4] = {{0, 1, 2, 4}, {1, 2, 3, 7}, {2, 4, 6, 7}, {1, 2, 4, 7}, {1, 4, 5, 7}};
  static const int TE2[5][4] = {{0, 1, 3, 5}, {0, 4, 5, 6}, {3, 5, 6, 7}, {0, 3, 5, 6}, {0, 2, 3, 6}};
  return even_one_based ? TE1 : TE2;
}
// splitting of the cube with 0-based origin (ci, cj, ck): the reference tests (ii + jj + kk) % 2 == 0 on 1-based indices
inline const int (*tet_grid_split_of(int ci, int cj, int ck))[4] { return tet_grid_split(((ci + cj + ck + 3) & 1) == 0); }

inline void build_tet_grid_tables(TetGridTables& T) {
  std::memset(&T, 0, sizeof(T));
  for (int ev = 0; ev < 2; ev++)
    for (int t = 0; t < 5; t++)
      for (int q = 0; q < 4; q++) T.te[ev][t][q] = tet_grid_split(ev)[t][q];
  for (int par = 0; par < 2; par++) {
    const int i = 4 + par, j = 4, k = 4;                 // an interior node with (i + j + k) & 1 == par
    int n = 0;
    // cubes in ascending element order: ci slowest, ck fastest
    for (int oi = -1; oi <= 0; oi++)
      for (int oj = -1; oj <= 0; oj++)
        for (int ok = -1; ok <= 0; ok++) {
          const int ci = i + oi, cj = j + oj, ck = k + ok;
          const int (*TE)[4] = tet_grid_split_of(ci, cj, ck);
          const int me = (-oi) | ((-oj) << 1) | ((-ok) << 2);     // local vertex of the node in this cube
          for (int t = 0; t < 5; t++) {
            int p = -1;
            for (int q = 0; q < 4; q++) if (TE[t][q] == me) p = q;
            if (p < 0) continue;
            T.inc[par][n][0] = oi; T.inc[par][n][1] = oj; T.inc[par][n][2] = ok; T.inc[par][n][3] = t; T.inc[par][n][4] = p;
            for (int q = 0; q < 4; q++) {
              const int v = TE[t][q], di = oi + (v & 1), dj = oj + ((v >> 1) & 1), dk = ok + (v >> 2);
              T.voff[par][n][q][0] = di; T.voff[par][n][q][1] = dj; T.voff[par][n][q][2] = dk;
              const int slot = (dk + 1) * 9 + (dj + 1) * 3 + (di + 1);
              T.vslot[par][n][q] = slot;
              T.present[par] |= 1 << slot;
              int& c = T.nsrc[par][slot];
              T.src[par][slot][c][0] = n; T.src[par][slot][c][1] = q; c++;
            }
            n++;
          }
        }
    T.ninc[par] = n;
  }
}

}  // namespace adfem
