"""GPU parity against the CPU oracle AT THE SIZES OF BASELINE.json's configs (through the C ABI, `-m gpu`).

The GPU assembles the full configuration (config 2: Mesh(4096,4096) = 33.5 M triangles; config 3: Mesh(4096,1024) P1 elasticity = 8.4 M;
config 4: 2 M jittered, randomly renumbered P2 triangles; config 5: Mesh3(64,64,64) = 1.3 M tetrahedra).  The
oracle keeps one heap object per element like the reference (2 us per element to build), so it assembles ELEMENT SUBSETS of the same mesh
as meshes of their own — same vertices, same element order, same coefficients:

  * every CSR row whose incident elements all lie in the subset must have exactly the subset's columns and, entry by entry, the subset's
    values (bar: |a-b| <= 1e-12 max(|a|, |b|, 1e-3 ||ref||_inf));
  * the adjoint of an element reads only dK entries between its own dofs, so the gradient of EVERY element of the subset must equal
    the oracle's with dK restricted to the subset's entries.

Subsets sit where the kernels change regime: the first and last node rows / cube layers, interior slabs that straddle the row chunks of the
structured kernels and the z-chunks of the two-stream tetrahedral forward, and boxes of the unstructured mesh.
"""
import numpy as np
import pytest

import adfem_jl_b200 as A
from adfem_jl_b200 import meshgen, ops

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
RTOL = 1e-12


def close(a, b, ref_scale, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    err = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), ref_scale * 1e-3)
    assert err.max() <= RTOL, f"{what}: max rel err {err.max():.3e} at {err.argmax()}"


def check_subset(oracle, mesh, S, op, cpg, coef_t, vals_t, dK_t, grad_t, what):
    """`S`: ascending element ids of the subset.  coef_t / vals_t / dK_t / grad_t: the full-size device tensors."""
    dim, d, g = mesh.dim, mesh.elem_ndof, mesh.gauss_per_elem
    nc = dim if op == 2 else 1
    S = np.asarray(S, dtype=np.int64)
    gverts, gconn = mesh.elems[S], mesh.conn[S]
    gv = np.unique(gverts)
    lverts = np.searchsorted(gv, gverts)
    o = (oracle.Mesh2D if dim == 2 else oracle.Mesh3D)(mesh.nodes[gv], lverts, degree=mesh.degree)
    assert np.array_equal(o.elems, lverts), "the subset keeps the orientation-fixed vertex order of the full mesh"
    l2g = np.full(o.ndof, -1, dtype=np.int64)
    l2g[o.conn] = gconn
    assert (l2g >= 0).all() and np.array_equal(l2g[o.conn], gconn), "element-wise dof correspondence subset <-> full mesh"
    # coefficients of the subset's elements (element-major, e*g + k, cpg values per Gauss point)
    gidx = ((S[:, None] * g + np.arange(g)[None, :])[:, :, None] * cpg + np.arange(cpg)[None, None, :]).reshape(-1)
    gidx_t = torch.from_numpy(gidx).cuda()
    coef = coef_t[gidx_t].cpu().numpy()
    N = nc * o.ndof
    fwd = {0: o.laplace_fwd, 1: o.mass_fwd, 2: o.stiffness_fwd}[op]
    bwd = {0: o.laplace_bwd, 1: o.mass_bwd, 2: o.stiffness_bwd}[op]
    ind, vv = fwd(coef)
    rp, ci, ref = oracle.canonical_csr(ind, vv, N)
    # position of every subset entry in the full-size value array
    rowptr, colind = mesh.csr_pattern(1)
    nnz_s, ndof = int(rowptr[-1]), mesh.ndof
    rows_l = np.repeat(np.arange(N, dtype=np.int64), np.diff(rp))
    a, rl = rows_l // o.ndof, rows_l % o.ndof
    b, cl = ci // o.ndof, ci % o.ndof
    gr, gc = l2g[rl], l2g[cl]
    urows = np.unique(l2g)
    lens = rowptr[urows + 1] - rowptr[urows]
    idx = np.repeat(rowptr[urows] - (np.cumsum(lens) - lens), lens) + np.arange(int(lens.sum()))
    key = np.repeat(urows, lens) * np.int64(ndof) + colind[idx]
    q = gr * np.int64(ndof) + gc
    k = np.searchsorted(key, q)
    assert (key[np.minimum(k, len(key) - 1)] == q).all(), "every entry of the subset exists in the full pattern"
    glen = rowptr[gr + 1] - rowptr[gr]
    pos = nc * (a * nnz_s + rowptr[gr]) + b * glen + (idx[k] - rowptr[gr])
    pos_t = torch.from_numpy(pos).cuda()
    # rows that are complete inside the subset: pattern (same length as the full row => same columns) and values
    cnt_full = np.bincount(mesh.conn.reshape(-1), minlength=ndof)
    cnt_sub = np.bincount(o.conn.reshape(-1), minlength=o.ndof)
    complete = cnt_full[l2g] == cnt_sub
    assert complete.sum() > 0.5 * o.ndof / 3
    ce = complete[rl]
    llen = np.diff(rp)[rows_l]
    assert np.array_equal(llen[ce], nc * glen[ce]), "complete rows have the full mesh's row length (pattern bit-exact)"
    scale = np.abs(ref).max()
    close(vals_t[pos_t].cpu().numpy()[ce], ref[ce], scale, what + " values")
    dK = dK_t[pos_t].cpu().numpy()
    expect = bwd(oracle.csr_adjoint_to_slots(rp, ci, dK, ind, N))
    close(grad_t[gidx_t].cpu().numpy(), expect, np.abs(expect).max(), what + " gradient")
    return len(S), int(ce.sum())


def run_config(mesh, op, cpg, seed=0):
    nc = mesh.dim if op == 2 else 1
    rowptr, _ = mesh.csr_pattern(1)
    nnz = nc * nc * int(rowptr[-1])
    gen = torch.Generator(device="cuda").manual_seed(seed)
    coef = torch.rand(mesh.ngauss * cpg, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    dK = torch.rand(nnz, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    fn = {0: ops.compute_fem_laplace_matrix1, 1: ops.compute_fem_mass_matrix1, 2: ops.compute_fem_stiffness_matrix}[op]
    shape = {1: (mesh.ngauss,), 9: (mesh.ngauss, 3, 3), 36: (mesh.ngauss, 6, 6)}[cpg]
    k = coef.view(shape).requires_grad_(True)
    T = fn(k, mesh, mode="csr")
    (g,) = torch.autograd.grad(T.values, k, dK)
    return coef, T.values.detach(), dK, g.reshape(-1)


def row_slab(m, r0, r1):
    """elements of cell rows [r0, r1) of Mesh(m, n, h): two triangles per cell, row-major cells (src/MFEM/MFEM.jl:134-146)"""
    return np.arange(2 * m * r0, 2 * m * r1, dtype=np.int64)


def test_config2_full_size(oracle):
    """Mesh(4096, 4096, 1/4096): the structured kernels (strips x row chunks, 64-bit row bases) against the oracle on five row slabs."""
    n = 4096
    mesh = A.Mesh(n, n, 1.0 / n)
    assert A._lib.lib().adfem_mesh_info(mesh.handle, A._lib.INFO_STRUCTURED) == 1
    for op in (0, 1):
        coef, vals, dK, grad = run_config(mesh, op, 1, seed=op)
        for r0, r1 in ((0, 12), (500, 516), (2040, 2056), (3333, 3345), (n - 12, n)):
            check_subset(oracle, mesh, row_slab(n, r0, r1), op, 1, coef, vals, dK, grad, "config 2 op %d rows %d-%d" % (op, r0, r1))
    # the general tile kernels at the same size (what a non-rectilinear mesh of this size gets)
    mesh.set_option("structured", 0)
    coef, vals, dK, grad = run_config(mesh, 0, 1, seed=5)
    for r0, r1 in ((0, 12), (2040, 2056), (n - 12, n)):
        check_subset(oracle, mesh, row_slab(n, r0, r1), 0, 1, coef, vals, dK, grad, "config 2 (tile kernels) rows %d-%d" % (r0, r1))


def test_config3_elasticity_large(oracle):
    """P1 elasticity with a 3x3 tangent per Gauss point on Mesh(4096, 1024, h) (8.4 M triangles): structured elasticity kernels."""
    m, n = 4096, 1024
    mesh = A.Mesh(m, n, 1.0 / m)
    coef, vals, dK, grad = run_config(mesh, 2, 9)
    for r0, r1 in ((0, 8), (300, 310), (509, 517), (n - 8, n)):
        check_subset(oracle, mesh, row_slab(m, r0, r1), 2, 9, coef, vals, dK, grad, "config 3 rows %d-%d" % (r0, r1))


def test_config4_p2_unstructured_large(oracle):
    """P2 Laplace and mass on 2 M jittered triangles with random diagonals, nodes and elements randomly renumbered (the tile kernels with
    Morton-ordered tiles): boxes at a corner, on an edge and in the interior."""
    n = 1000
    c, e = meshgen.jitter_unstructured(n, n, 1.0 / n, seed=2)
    mesh = A.Mesh(c, e, degree=2)
    cen = mesh.nodes[mesh.elems].mean(1)
    boxes = ((0.0, 0.09, 0.0, 0.09), (0.45, 0.56, 0.93, 1.0), (0.40, 0.50, 0.30, 0.39), (0.91, 1.0, 0.91, 1.0))
    for op in (0, 1):
        coef, vals, dK, grad = run_config(mesh, op, 1, seed=10 + op)
        for x0, x1, y0, y1 in boxes:
            S = np.flatnonzero((cen[:, 0] >= x0) & (cen[:, 0] <= x1) & (cen[:, 1] >= y0) & (cen[:, 1] <= y1))
            ns, nrows = check_subset(oracle, mesh, S, op, 1, coef, vals, dK, grad, "config 4 op %d box %s" % (op, (x0, x1, y0, y1)))
            assert ns > 5000 and nrows > 5000


def test_config5_tet_elasticity_large(oracle):
    """P1 tetrahedral elasticity (6x6 Voigt tangent per Gauss point) on Mesh3(64, 64, 64, h) = 1.3 M tetrahedra: Gauss pre-sum + node kernel,
    adjoint one warp per 32 tetrahedra of a cube column; z-slabs at the bottom, in the interior and at the top.  The optional z-chunk
    pipeline of the forward (two streams) must give the same bits."""
    n = l = 64
    c, e = meshgen.tet_grid(n, n, l, 1.0 / n)
    mesh = A.Mesh3(c, e)
    assert A._lib.lib().adfem_mesh_info(mesh.handle, A._lib.INFO_STRUCTURED) == 2
    coef, vals, dK, grad = run_config(mesh, 2, 36)
    cube = np.arange(n * n * l, dtype=np.int64)          # cube index (ci*n + cj)*l + ck, five tetrahedra per cube
    ck = cube % l
    for k0, k1 in ((0, 3), (6, 10), (30, 34), (l - 3, l)):
        S = (5 * cube[(ck >= k0) & (ck < k1)][:, None] + np.arange(5)[None, :]).reshape(-1)
        check_subset(oracle, mesh, S, 2, 36, coef, vals, dK, grad, "config 5 layers %d-%d" % (k0, k1))
    v2 = ops.compute_fem_stiffness_matrix(coef.view(-1, 6, 6), mesh, mode="csr").values
    assert torch.equal(v2, vals), "run-to-run bit-identical (no atomics, fixed summation order)"
    mesh.set_option("tet_chunks", 8)
    v3 = ops.compute_fem_stiffness_matrix(coef.view(-1, 6, 6), mesh, mode="csr").values
    assert torch.equal(v3, vals), "z-chunk pipeline on two streams: same bits"
