"""Structured mesh generators mirroring the reference's Julia constructors (host logic, numpy only).

`tri_grid`  follows src/MFEM/MFEM.jl:134-170  (`Mesh(m, n, h; version)`), 
`tet_grid`  follows src/MFEM3/MFEM.jl:124-185 (`Mesh3(m, n, l, h)`, 5 tets per cube, parity-alternating).
Returned connectivity is 0-based (the Julia arrays are 1-based).
"""
import numpy as np


def tri_grid(m, n, h, version=1, rng=None, dtype=np.int64):
    """(m+1)(n+1) nodes, 2mn triangles. version 1/2 = the two diagonal directions, 3 = random per cell.  `dtype`: integer type of the
    connectivity (the mesh constructors ask for int32, what the library takes, when the node ids fit)."""
    assert np.dtype(dtype) == np.int64 or (m + 1) * (n + 1) < 2 ** 31
    a = (np.arange(n, dtype=dtype)[:, None] * (m + 1) + np.arange(m, dtype=dtype)[None, :]).reshape(-1)   # lower-left node of cell (row ii, col jj)

    def fill(el, mask, tris):          # el[cells, 2, 3] <- the two triangles `tris` (offsets from a) of the cells in `mask`, without temporaries
        if mask is None:               # one contiguous pass over the output
            np.add(a[:, None, None], np.asarray(tris, dtype=dtype)[None], out=el)
            return
        for t in range(2):
            for k in range(3):
                el[mask, t, k] = a[mask] + tris[t][k]
    v1 = ((0, 1, m + 1), (1, m + 1, m + 2))                                                                # MFEM.jl:143-145
    v2 = ((0, m + 2, m + 1), (0, 1, m + 2))                                                                # MFEM.jl:146-148
    el = np.empty((m * n, 2, 3), dtype=dtype)
    if version == 1:
        fill(el, None, v1)
    elif version == 2:
        fill(el, None, v2)
    elif version == 3:
        rng = np.random.default_rng(0) if rng is None else rng
        pick = rng.random(m * n) > 0.5                                                                     # MFEM.jl:149-156
        np.take(np.asarray((v2, v1), dtype=dtype), pick.astype(np.intp), axis=0, out=el)                  # the cell's pair of triangles, by its coin
        el += a[:, None, None]
    else:
        raise ValueError("version must be 1, 2 or 3")
    elems = el.reshape(2 * m * n, 3)
    coords = np.empty((n + 1, m + 1, 2))
    coords[:, :, 0] = (np.arange(m + 1) * float(h))[None, :]
    coords[:, :, 1] = (np.arange(n + 1) * float(h))[:, None]
    return coords.reshape(-1, 2), elems


_TE1 = np.array([[1, 2, 3, 5], [2, 3, 4, 8], [3, 5, 7, 8], [2, 3, 5, 8], [2, 5, 6, 8]]) - 1   # MFEM.jl:131-137
_TE2 = np.array([[1, 2, 4, 6], [1, 5, 6, 7], [4, 6, 7, 8], [1, 4, 6, 7], [1, 3, 4, 7]]) - 1   # MFEM.jl:138-144


def tet_grid(m, n, l, h, dtype=np.int64):
    """(m+1)(n+1)(l+1) nodes, 5mnl tets. Only consistent for m == n (reference quirk Q9).  `dtype`: integer type of the connectivity."""
    assert np.dtype(dtype) == np.int64 or (m + 1) * (n + 1) * (l + 1) < 2 ** 31
    # coords: for k; for j in 1:m+1; for i in 1:n+1  -> x=(i-1)h fastest
    coords = np.empty((l + 1, m + 1, n + 1, 3))
    coords[..., 0] = (np.arange(n + 1) * float(h))[None, None, :]
    coords[..., 1] = (np.arange(m + 1) * float(h))[None, :, None]
    coords[..., 2] = (np.arange(l + 1) * float(h))[:, None, None]
    # 0-based id of node (i, j, k), 1-based arguments: (k-1)(n+1)(m+1) + (j-1)(m+1) + i-1 (MFEM.jl:128-130); cells in the order of the loops
    # `for i; for j; for k`.  A cell's 8 corners are its first node plus fixed offsets, its 5 tetrahedra pick 4 corners each (TE1 / TE2 by the
    # parity of i+j+k): one gather of the 5 x 4 offset table per cell and one in-place add of the cell's first node.
    sj, sk = m + 1, (n + 1) * (m + 1)
    i0, j0, k0 = np.arange(n, dtype=dtype)[:, None, None], np.arange(m, dtype=dtype)[None, :, None], np.arange(l, dtype=dtype)[None, None, :]
    base = (i0 + j0 * sj + k0 * sk).reshape(-1)
    even = ((i0 + j0 + k0) % 2 == 1).reshape(-1)                 # 1-based (i + j + k) even
    corner = np.array([0, 1, sj, sj + 1, sk, sk + 1, sk + sj, sk + sj + 1], dtype=dtype)
    table = np.stack([corner[_TE2], corner[_TE1]])               # [parity][tet][vertex]
    elems = table[even.astype(np.intp)]                          # ncell x 5 x 4
    elems += base[:, None, None]
    return coords.reshape(-1, 3), elems.reshape(-1, 4)


def jitter_unstructured(m, n, h, seed=2, jitter=0.3, permute=True):
    """BASELINE config 4's synthetic unstructured mesh: structured grid, interior nodes jittered by
    U(-jitter*h, jitter*h), per-cell random diagonal (version 3), random node + element renumbering."""
    rng = np.random.default_rng(seed)
    coords, elems = tri_grid(m, n, h, version=3, rng=rng)
    ix = np.arange((m + 1) * (n + 1)) % (m + 1)
    iy = np.arange((m + 1) * (n + 1)) // (m + 1)
    interior = (ix > 0) & (ix < m) & (iy > 0) & (iy < n)
    coords = coords + interior[:, None] * rng.uniform(-jitter * h, jitter * h, coords.shape)
    if permute:
        pn = rng.permutation(coords.shape[0])          # new id of old node i is inv[i]
        inv = np.empty_like(pn)
        inv[pn] = np.arange(len(pn))
        coords = np.take(coords, pn, axis=0)           # np.take: the same gathers as fancy indexing at less than half the time
        elems = np.take(inv, elems)
        elems = np.take(elems, rng.permutation(elems.shape[0]), axis=0)
    return coords, elems


def morton_element_order(coords, elems):
    """Element permutation that sorts the centroids along a Morton curve: contiguous element blocks become spatially compact
    (used before an element-block partition of a mesh whose element numbering has no locality)."""
    c = np.take(coords, elems, axis=0).mean(1)
    lo, hi = c.min(0), c.max(0)
    q = np.minimum(((c - lo) / np.maximum(hi - lo, 1e-300) * 65535).astype(np.uint64), 65535)
    dim = c.shape[1]
    key = np.zeros(len(c), dtype=np.uint64)
    # bit b of coordinate d goes to bit dim*b + d: spread the 16 bits of a coordinate with the usual shift-and-mask cascade
    cascade = {2: ((8, 0x00FF00FF), (4, 0x0F0F0F0F), (2, 0x33333333), (1, 0x55555555)),
               3: ((32, 0x1F00000000FFFF), (16, 0x1F0000FF0000FF), (8, 0x100F00F00F00F00F), (4, 0x10C30C30C30C30C3), (2, 0x1249249249249249))}
    if dim in cascade:
        for d in range(dim):
            x = q[:, d].copy()
            for sh, mask in cascade[dim]:
                x = (x | (x << np.uint64(sh))) & np.uint64(mask)
            key |= x << np.uint64(d)
    else:
        for b in range(16):
            for d in range(dim):
                key |= ((q[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(dim * b + d)
    return np.argsort(key, kind="stable")
