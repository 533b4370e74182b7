"""Structured triangulation fast path (adfem.jl_b200/csrc/tri_grid.cuh): detection + closed-form row pointers on the CPU,
parity of the index-free kernels against the oracle and against the general tile kernels on the GPU."""
import ctypes as C

import numpy as np
import pytest

import adfem_jl_b200 as A
from adfem_jl_b200 import _lib, meshgen


def _info(m, what):
    return int(_lib.lib().adfem_mesh_info(m.handle, what))


def rectilinear(m, n, seed):
    """Mesh(m, n, h) connectivity on non-uniform rectilinear nodes, optionally shifted (a slab of a larger grid)."""
    rng = np.random.default_rng(seed)
    c, e = meshgen.tri_grid(m, n, 1.0)
    xs = np.cumsum(rng.uniform(0.5, 1.5, m + 1)) - 3.0
    ys = np.cumsum(rng.uniform(0.5, 1.5, n + 1)) + 7.0
    c = np.stack([xs[c[:, 0].astype(int)], ys[c[:, 1].astype(int)]], 1)
    return c, e


def mapped(m, n, seed):
    """Mesh(m, n, h) connectivity on a smoothly mapped AND jittered grid: no two nodes share an abscissa or ordinate, every triangle keeps its
    orientation (jitter below a quarter of the local spacing)."""
    rng = np.random.default_rng(seed)
    c, e = rectilinear(m, n, seed)
    x, y = c[:, 0].copy(), c[:, 1].copy()
    c = np.stack([x + 0.08 * np.sin(0.7 * y), y + 0.06 * np.cos(0.9 * x)], 1) + rng.uniform(-0.1, 0.1, c.shape)
    return c, e


GRIDS = [(1, 1), (2, 1), (1, 3), (5, 4), (31, 2), (32, 3), (63, 5), (70, 33)]


@pytest.mark.parametrize("m,n", GRIDS)
def test_detection_and_closed_form_rowptr(m, n):
    for c, e in (meshgen.tri_grid(m, n, 0.1), rectilinear(m, n, 1)):
        M = A.Mesh(c, e, host_only=True)
        assert _info(M, _lib.INFO_STRUCTURED) == 1
        M.csr_pattern(1)                       # builds the symbolic pattern, which validates the closed-form row pointers
        assert _info(M, _lib.INFO_STRUCTURED) == 1


def test_not_structured():
    c, e = meshgen.tri_grid(6, 5, 0.1, version=2)                # other diagonal
    assert _info(A.Mesh(c, e, host_only=True), _lib.INFO_STRUCTURED) == 0
    c, e = meshgen.tri_grid(6, 5, 0.1)
    c2 = c.copy(); c2[9, 0] += 1e-13                             # one node off the rectilinear grid: structured connectivity, MAPPED coordinates (3)
    assert _info(A.Mesh(c2, e, host_only=True), _lib.INFO_STRUCTURED) == 3
    c3 = c.copy(); c3[9] = c3[9 + 7 + 1] + 0.01                   # a node pushed across its cell: a triangle flips, the orientation fix reorders it
    assert _info(A.Mesh(c3, e, host_only=True), _lib.INFO_STRUCTURED) == 0
    assert _info(A.Mesh(c, e[::-1].copy(), host_only=True), _lib.INFO_STRUCTURED) == 0      # renumbered elements
    assert _info(A.Mesh(c, e, degree=2, host_only=True), _lib.INFO_STRUCTURED) == 0          # P2
    c, e = meshgen.jitter_unstructured(6, 5, 0.1)
    assert _info(A.Mesh(c, e, host_only=True), _lib.INFO_STRUCTURED) == 0


@pytest.mark.parametrize("m,n", [(1, 1), (1, 6), (6, 1), (2, 2), (7, 5), (70, 64), (300, 3)])
@pytest.mark.parametrize("mapped", [False, True])
def test_closed_form_symbolic_tables(oracle, m, n, mapped):
    """Mesh(m, n, h): the symbolic tables from index arithmetic (ScalarPattern::build_tri_grid, the default on the structured triangulation) are
    the bytes the general build produces (option structured_pattern = 0) — pattern, slot map, both tile plans (which read the dof -> element
    adjacency) — for 1 and 3 host threads, and the pattern is the oracle's."""
    c, e = meshgen.tri_grid(m, n, 0.25)
    if mapped:
        rng = np.random.default_rng(m * 31 + n)
        c = np.ascontiguousarray(np.stack([c[:, 0] + 0.02 * np.sin(3.0 * c[:, 1]), c[:, 1] + 0.02 * np.cos(2.0 * c[:, 0])], 1) + rng.uniform(-0.02, 0.02, c.shape))
    got = {}
    for closed in (1, 0):
        for threads in (1, 3):
            M = A.Mesh(c, e, host_only=True)
            assert _info(M, _lib.INFO_STRUCTURED) == (3 if mapped else 1)
            M.set_option("structured_pattern", closed)
            M.set_option("host_threads", threads)
            rowptr, colind = M.csr_pattern(1)
            got[closed, threads] = [rowptr, colind, M.slot_to_nnz()] + [M.plan_array(w, nc, a, np.int64 if a == 0 else np.uint8) for nc in (1, 2) for w in (0, 1) for a in (0, 1)]
            assert _info(M, _lib.INFO_STRUCTURED) == (3 if mapped else 1)         # the closed-form row pointers of the kernels still validate
    ref = got[0, 1]
    for key, arrs in got.items():
        for x, y in zip(ref, arrs):
            assert x.shape == y.shape and np.array_equal(x, y), key
    o = oracle.Mesh2D(c, e)
    ind, vv = o.laplace_fwd(np.ones(o.ngauss))
    rp, ci, _ = oracle.canonical_csr(ind, vv, o.ndof)
    assert np.array_equal(ref[0], rp) and np.array_equal(ref[1], ci)
    with pytest.raises(_lib.AdfemError):
        M.set_option("structured_pattern", 1)                                   # too late: the tables exist


@pytest.mark.parametrize("n,l", [(1, 1), (2, 3), (4, 2), (5, 5), (17, 16), (3, 40)])
def test_closed_form_symbolic_tables_tetrahedra(oracle, n, l):
    """Mesh3(n, n, l, h): the symbolic tables from the two parity neighbourhoods (ScalarPattern::build_tet_grid, the default on the structured
    tetrahedral grid) are the bytes of the general build (structured_pattern = 0) — pattern, slot map, tile plans — for 1 and 3 host threads,
    also when elements list their vertices in another order (the local positions are read from the mesh), and the pattern is the oracle's."""
    c, e = meshgen.tet_grid(n, n, l, 0.25)
    rng = np.random.default_rng(n * 7 + l)
    e2 = e.copy()
    for at in rng.choice(len(e2), size=max(1, len(e2) // 3), replace=False):
        e2[at] = e2[at][rng.permutation(4)]
    for conn in (e, e2):
        got = {}
        for closed in (1, 0):
            for threads in (1, 3):
                M = A.Mesh3(c, conn, host_only=True)
                assert _info(M, _lib.INFO_STRUCTURED) == 2
                M.set_option("structured_pattern", closed)
                M.set_option("host_threads", threads)
                rowptr, colind = M.csr_pattern(1)
                got[closed, threads] = [rowptr, colind, M.slot_to_nnz()] + [M.plan_array(w, nc, a, np.int64 if a == 0 else np.uint8) for nc in (1, 3) for w in (0, 1) for a in (0, 1)]
                assert _info(M, _lib.INFO_STRUCTURED) == 2
        ref = got[0, 1]
        for key, arrs in got.items():
            for x, y in zip(ref, arrs):
                assert x.shape == y.shape and np.array_equal(x, y), key
        if n * n * l <= 5 ** 3:
            o = oracle.Mesh3D(c, conn)
            ind, vv = o.laplace_fwd(np.ones(o.ngauss))
            rp, ci, _ = oracle.canonical_csr(ind, vv, o.ndof)
            assert np.array_equal(ref[0], rp) and np.array_equal(ref[1], ci)


@pytest.mark.parametrize("n,l", [(1, 1), (2, 3), (4, 2), (5, 5), (20, 9)])
def test_tet_grid_detection(n, l):
    """Mesh3(n, n, l, h) (5 tetrahedra per cube, parity-alternating) is recognised from its arrays, also on rectilinear non-uniform coordinates;
    the closed-form rows of csrc/tet_grid.cuh are validated against the symbolic pattern when it is built."""
    c, e = meshgen.tet_grid(n, n, l, 0.25)
    M = A.Mesh3(c, e, host_only=True)
    assert _info(M, _lib.INFO_STRUCTURED) == 2
    M.csr_pattern(1)
    assert _info(M, _lib.INFO_STRUCTURED) == 2
    rng = np.random.default_rng(n + l)
    xs, ys, zs = (np.concatenate([[0.0], np.cumsum(rng.random(k) + 0.1)]) for k in (n, n, l))
    kk, jj, ii = np.meshgrid(np.arange(l + 1), np.arange(n + 1), np.arange(n + 1), indexing="ij")
    c2 = np.stack([xs[ii.reshape(-1)], ys[jj.reshape(-1)], zs[kk.reshape(-1)]], 1)
    assert _info(A.Mesh3(c2, e, host_only=True), _lib.INFO_STRUCTURED) == 2
    c3 = c.copy(); c3[-1, 2] += 1e-13                                         # one node off the grid
    assert _info(A.Mesh3(c3, e, host_only=True), _lib.INFO_STRUCTURED) == 0
    if n * n * l > 1:
        assert _info(A.Mesh3(c, e[::-1].copy(), host_only=True), _lib.INFO_STRUCTURED) == 0   # renumbered elements
    assert _info(A.Mesh3(c, e, degree=2, host_only=True), _lib.INFO_STRUCTURED) == 0          # P2
    e2 = e.copy(); e2[0] = e2[0][[1, 0, 2, 3]]                                               # another vertex order of the same tetrahedron is fine
    assert _info(A.Mesh3(c, e2, host_only=True), _lib.INFO_STRUCTURED) == 2
    if n >= 16:                                    # the element check runs in slabs over the host threads: one foreign tetrahedron in any slab is found
        for at in (3, len(e) // 2 + 7, len(e) - 2):
            e4 = e.copy(); e4[at, 3] = (e4[at, 3] + 2 * (n + 1)) % len(c)
            assert _info(A.Mesh3(c, e4, host_only=True), _lib.INFO_STRUCTURED) == 0


# ---------------------------------------------------------------------------------------------- GPU
def _close(a, b, rel=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), scale * 1e-3)
    assert err.max() <= rel, f"max rel err {err.max():.3e} at {err.argmax()}"


@pytest.mark.parametrize("m,n", [(1, 1), (5, 4), (70, 33)])
def test_mapped_grid_detection(m, n):
    """structured connectivity on mapped / jittered node positions: recognised as kind 3 (scalar CSR operators take the index-free kernels with
    positions from the coordinate array), the closed-form row pointers still validate"""
    c, e = mapped(m, n, 5)
    M = A.Mesh(c, e, host_only=True)
    assert _info(M, _lib.INFO_STRUCTURED) == 3
    M.csr_pattern(1)
    assert _info(M, _lib.INFO_STRUCTURED) == 3


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["uniform", "rectilinear", "mapped"])
@pytest.mark.parametrize("m,n", GRIDS)
def test_structured_parity(oracle, m, n, kind):
    import torch
    from adfem_jl_b200 import ops
    c, e = meshgen.tri_grid(m, n, 0.37) if kind == "uniform" else (rectilinear(m, n, 2) if kind == "rectilinear" else mapped(m, n, 3))
    M, o = A.Mesh(c, e), oracle.Mesh2D(c, e)
    assert _info(M, _lib.INFO_STRUCTURED) == (3 if kind == "mapped" else 1)
    rng = np.random.default_rng(7)
    coef = rng.random(o.ngauss) + 0.5
    rowptr, colind = M.csr_pattern(1)
    for fn, ofwd, obwd in ((ops.compute_fem_laplace_matrix1, o.laplace_fwd, o.laplace_bwd), (ops.compute_fem_mass_matrix1, o.mass_fwd, o.mass_bwd)):
        ind, vv = ofwd(coef)
        rp, ci, ref = oracle.canonical_csr(ind, vv, o.ndof)
        assert np.array_equal(rowptr, rp) and np.array_equal(colind, ci)
        dv = rng.standard_normal(len(ref))
        expect = obwd(oracle.csr_adjoint_to_slots(rp, ci, dv, ind, o.ndof))
        out = {}
        for structured, rows, heron in ((1, 0, 0), (1, 3, 1), (1, 1, 0), (0, 0, 0)):
            M.set_option("structured", structured)
            M.set_option("grid_rows", rows)
            M.set_option("area_formula_csr", heron)
            k = torch.from_numpy(coef).cuda().requires_grad_(True)
            T = fn(k, M, mode="csr")
            vals = T.values.detach().cpu().numpy()
            (g,) = torch.autograd.grad(T.values, k, torch.from_numpy(dv).cuda())
            _close(vals, ref)
            _close(g.cpu().numpy(), expect)
            out[(structured, rows, heron)] = (vals, g.cpu().numpy())
        # same per-entry summation order as the general tile kernels: agreement within the parity bar (in practice ~1e-15)
        _close(out[(1, 0, 0)][0], out[(0, 0, 0)][0], rel=1e-12)
        _close(out[(1, 0, 0)][1], out[(0, 0, 0)][1], rel=1e-12)
        assert np.array_equal(out[(1, 0, 0)][0], out[(1, 1, 0)][0])          # independent of the row chunking
        M.set_option("structured", 1); M.set_option("grid_rows", 0); M.set_option("area_formula_csr", 0)


@pytest.mark.gpu
def test_structured_large_properties():
    """Config-2-like size (not square, not a multiple of the strip width): K·1 = 0, symmetry, Σ M = area, adjoint identity, and
    bit-identical repeat; the general tile kernels on the same mesh agree to 1e-13."""
    import scipy.sparse as sp
    import torch
    from adfem_jl_b200 import ops
    m, n = 1500, 1111
    M = A.Mesh(m, n, 1.0 / 1024)
    assert _info(M, _lib.INFO_STRUCTURED) == 1
    G = M.ngauss
    gen = torch.Generator(device="cuda").manual_seed(0)
    k1 = torch.rand(G, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    rowptr, colind = M.csr_pattern(1)
    v1 = ops.compute_fem_laplace_matrix1(k1, M, mode="csr").values
    assert torch.equal(v1, ops.compute_fem_laplace_matrix1(k1, M, mode="csr").values)
    K = sp.csr_matrix((v1.cpu().numpy(), colind, rowptr), shape=(M.ndof, M.ndof))
    assert np.abs(K @ np.ones(M.ndof)).max() < 1e-10
    assert abs(K - K.T).max() < 1e-12
    Mv = ops.compute_fem_mass_matrix1(torch.ones(G, dtype=torch.float64, device="cuda"), M, mode="csr").values
    area = m * n / 1024.0 ** 2
    assert abs(Mv.sum().item() - area) < 1e-11 * area
    dK = torch.randn(v1.numel(), dtype=torch.float64, device="cuda", generator=gen)
    kk = k1.clone().requires_grad_(True)
    (g,) = torch.autograd.grad(ops.compute_fem_laplace_matrix1(kk, M, mode="csr").values, kk, dK)
    lhs, rhs = (dK * v1).sum().item(), (g * k1).sum().item()
    assert abs(lhs - rhs) < 1e-10 * max(abs(lhs), 1.0)
    M.set_option("structured", 0)
    v0 = ops.compute_fem_laplace_matrix1(k1, M, mode="csr").values
    (g0,) = torch.autograd.grad(ops.compute_fem_laplace_matrix1(kk, M, mode="csr").values, kk, dK)
    assert (v0 - v1).abs().max().item() <= 1e-13 * v1.abs().max().item()
    assert (g0 - g).abs().max().item() <= 1e-13 * g.abs().max().item()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["rectilinear", "mapped"])
@pytest.mark.parametrize("m,n", [(1, 1), (5, 4), (63, 5), (70, 33)])
def test_structured_host_buffer_pipeline(m, n, kind):
    """adfem_assemble_csr_host / _adjoint_host on a structured mesh stream node-row chunks through three CUDA streams; any chunking
    must give exactly the device-pointer result."""
    import torch
    from adfem_jl_b200 import ops
    c, e = rectilinear(m, n, 3) if kind == "rectilinear" else mapped(m, n, 3)
    M = A.Mesh(c, e)
    L = _lib.lib()
    rng = np.random.default_rng(11)
    rowptr, _ = M.csr_pattern(1)
    nnz = int(rowptr[-1])
    for op, fn in ((0, ops.compute_fem_laplace_matrix1), (1, ops.compute_fem_mass_matrix1)):
        coef = rng.random(M.ngauss) + 0.5
        dv = rng.standard_normal(nnz)
        k = torch.from_numpy(coef).cuda().requires_grad_(True)
        T = fn(k, M, mode="csr")
        (g,) = torch.autograd.grad(T.values, k, torch.from_numpy(dv).cuda())
        ref_v, ref_g = T.values.detach().cpu().numpy(), g.cpu().numpy()
        for chunks in (1, 2, 3, 16, 1000):
            M.set_option("host_chunks", chunks)
            vals, grad = np.full(nnz, np.nan), np.full(M.ngauss, np.nan)
            _lib.check(L.adfem_assemble_csr_host(M.handle, C.c_int(op), coef.ctypes.data_as(_lib.c_dp), vals.ctypes.data_as(_lib.c_dp)))
            _lib.check(L.adfem_assemble_csr_adjoint_host(M.handle, C.c_int(op), dv.ctypes.data_as(_lib.c_dp), grad.ctypes.data_as(_lib.c_dp)))
            assert np.array_equal(vals, ref_v) and np.array_equal(grad, ref_g)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["uniform", "rectilinear", "mapped"])
@pytest.mark.parametrize("m,n", GRIDS)
def test_structured_source_term(oracle, m, n, kind):
    """compute_fem_source_term1 and its adjoint on the structured path (rectilinear and mapped grids) vs the oracle and vs the general kernels."""
    import torch
    from adfem_jl_b200 import ops
    c, e = meshgen.tri_grid(m, n, 0.37) if kind == "uniform" else (rectilinear(m, n, 4) if kind == "rectilinear" else mapped(m, n, 4))
    M, o = A.Mesh(c, e), oracle.Mesh2D(c, e)
    rng = np.random.default_rng(13)
    f, gr = rng.standard_normal(o.ngauss), rng.standard_normal(o.ndof)
    out = {}
    for structured, rows in ((1, 0), (1, 2), (0, 0)):
        M.set_option("structured", structured)
        M.set_option("grid_rows", rows)
        ft = torch.from_numpy(f).cuda().requires_grad_(True)
        rhs = ops.compute_fem_source_term1(ft, M)
        (g,) = torch.autograd.grad(rhs, ft, torch.from_numpy(gr).cuda())
        _close(rhs.detach().cpu().numpy(), o.source_fwd(f))
        _close(g.cpu().numpy(), o.source_bwd(gr))
        out[(structured, rows)] = (rhs.detach().cpu().numpy(), g.cpu().numpy())
    _close(out[(1, 0)][0], out[(0, 0)][0], rel=1e-12)             # different association of the 18 terms of a node; sums cancel (f has both signs)
    _close(out[(1, 0)][1], out[(0, 0)][1], rel=1e-12)
    _close(out[(1, 0)][0], out[(1, 2)][0], rel=1e-12)             # row chunking only changes which inlined copy of the cell code runs
    _close(out[(1, 0)][1], out[(1, 2)][1], rel=1e-12)
    M.set_option("structured", 1); M.set_option("grid_rows", 0)
