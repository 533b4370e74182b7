"""GPU parity tests (run on the B200 box with `-m gpu`): every CUDA path, called through the C ABI of
libadfem_cuda.so, against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): CSR pattern bit-exact; indices bit-exact; values and gradients within a
relative 1e-12 in fp64, measured as |a-b| <= 1e-12 * max(|a|, |b|, 1e-3*||ref||_inf) so that entries that
are analytically zero on right-angle triangles (pure cancellation) do not fail on rounding noise.
"""
import ctypes as C
import os

import numpy as np
import pytest

import adfem_jl_b200 as A
from adfem_jl_b200 import meshgen, ops

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL = 1e-12


def close(a, b, rel=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), scale * 1e-3)
    assert err.max() <= rel, f"max rel err {err.max():.3e} at {err.argmax()}"


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _meshes():
    out = {}
    out["tri_struct"] = (2, *meshgen.tri_grid(33, 20, 0.05))
    out["tri_unstruct"] = (2, *meshgen.jitter_unstructured(40, 37, 0.03, seed=3))
    out["tet"] = (3, *meshgen.tet_grid(6, 6, 5, 0.2))
    return out


MESHES = _meshes()
CASES = [(name, deg) for name in MESHES for deg in (1, 2)]


def make(name, degree, oracle):
    dim, c, e = MESHES[name]
    if dim == 2:
        return A.Mesh(c, e, degree=degree), oracle.Mesh2D(c, e, degree=degree)
    return A.Mesh3(c, e, degree=degree), oracle.Mesh3D(c, e, degree=degree)


# ------------------------------------------------------------------------------------------ COO-compatible mode
@pytest.mark.parametrize("name,degree", CASES)
def test_coo_scalar_ops(oracle, name, degree):
    m, o = make(name, degree, oracle)
    rng = np.random.default_rng(0)
    coef = rng.random(o.ngauss) + 0.5
    for op_name, fwd, bwd in (("laplace", o.laplace_fwd, o.laplace_bwd), ("mass", o.mass_fwd, o.mass_bwd)):
        fn = ops.compute_fem_laplace_matrix1 if op_name == "laplace" else ops.compute_fem_mass_matrix1
        k = dev(coef).requires_grad_(True)
        Sp = fn(k, m, mode="coo")
        ind, vv = fwd(coef)
        assert np.array_equal(Sp.indices.cpu().numpy(), ind)          # bit-exact slot order and dof ids
        close(Sp.values.detach().cpu().numpy(), vv)
        gv = rng.standard_normal(len(vv))
        (g,) = torch.autograd.grad(Sp.values, k, dev(gv))
        close(g.cpu().numpy(), bwd(gv))


@pytest.mark.parametrize("name,degree", CASES)          # incl. P2 tetrahedra (3-D extension N2: 30 x 30 local matrix, 6 x 6 Voigt tangent)
def test_coo_stiffness(oracle, name, degree):
    m, o = make(name, degree, oracle)
    ns = 3 if m.dim == 2 else 6
    rng = np.random.default_rng(1)
    H = rng.random((o.ngauss, ns, ns))                                 # deliberately unsymmetric
    k = dev(H).requires_grad_(True)
    Sp = ops.compute_fem_stiffness_matrix(k, m, mode="coo")
    ind, vv = o.stiffness_fwd(H.reshape(-1))
    assert np.array_equal(Sp.indices.cpu().numpy(), ind)
    close(Sp.values.detach().cpu().numpy(), vv)
    gv = rng.standard_normal(len(vv))
    (g,) = torch.autograd.grad(Sp.values, k, dev(gv))
    close(g.cpu().numpy().reshape(-1), o.stiffness_bwd(gv))


@pytest.mark.parametrize("name,degree", CASES)
def test_source_term(oracle, name, degree):
    m, o = make(name, degree, oracle)
    rng = np.random.default_rng(2)
    f = rng.standard_normal(o.ngauss)
    ft = dev(f).requires_grad_(True)
    rhs = ops.compute_fem_source_term1(ft, m)
    close(rhs.detach().cpu().numpy(), o.source_fwd(f))
    gr = rng.standard_normal(o.ndof)
    (g,) = torch.autograd.grad(rhs, ft, dev(gr))
    close(g.cpu().numpy(), o.source_bwd(gr))
    assert np.allclose(ops.compute_fem_source_term1(f, m), o.source_fwd(f), rtol=1e-12, atol=1e-14)   # eager numpy path


# ------------------------------------------------------------------------------------------ CSR fast path
def _csr_check(m, o, oracle, op, coef, ofwd, obwd, ncomp, tiles):
    n = ncomp * o.ndof
    ind, vv = ofwd(coef.reshape(-1))
    rp, ci, ref = oracle.canonical_csr(ind, vv, n)
    rowptr, colind = m.csr_pattern(ncomp)
    assert np.array_equal(rowptr, rp) and np.array_equal(colind, ci)   # pattern bit-exact
    fn = {0: ops.compute_fem_laplace_matrix1, 1: ops.compute_fem_mass_matrix1, 2: ops.compute_fem_stiffness_matrix}[op]
    if tiles:
        m.set_option("rows_per_tile", tiles[0])
        m.set_option("elems_per_tile", tiles[1])
    k = dev(coef).requires_grad_(True)
    T = fn(k, m, mode="csr")
    close(T.values.detach().cpu().numpy(), ref)
    rng = np.random.default_rng(5)
    dv = rng.standard_normal(len(ref))
    expect = obwd(oracle.csr_adjoint_to_slots(rp, ci, dv, ind, n))
    # tiled / direct-gather adjoint, both area formulas, CTA sizes, persistent (software-pipelined) / one-CTA-per-tile launches,
    # coefficient prefetch into registers on / off
    # grid_limit: 0 = every tile gets its own CTA on these small meshes; 1/2/3/5 CTAs walk ALL tiles (ring wrap-around of the pipeline)
    for tiled, heron, threads, pipe, kpre, glim in ((1, 0, 256, 1, 1, 0), (0, 0, 256, 1, 0, 0), (1, 1, 128, 0, 1, 0), (1, 0, 512, 0, 0, 0), (1, 0, 96, 1, 1, 1),
                                                    (1, 1, 64, 1, 0, 2), (1, 0, 512, 1, 1, 3), (1, 0, 320, 1, 1, 5), (1, 0, 320, 1, 0, 5)):
        m.set_option("grid_limit", glim)
        m.set_option("adjoint_tiled", tiled)
        m.set_option("area_formula_csr", heron)
        m.set_option("tile_threads", threads)
        m.set_option("pipeline", pipe)
        m.set_option("coef_prefetch", kpre)
        T2 = fn(k, m, mode="csr")
        close(T2.values.detach().cpu().numpy(), ref)
        (g,) = torch.autograd.grad(T2.values, k, dev(dv))
        close(g.cpu().numpy().reshape(-1), expect)
    m.set_option("area_formula_csr", 0)
    m.set_option("tile_threads", 0)
    m.set_option("pipeline", 1)
    m.set_option("coef_prefetch", 1)
    m.set_option("grid_limit", 0)
    # eager numpy path returns the same matrix as a scipy CSR
    S = fn(coef, m)
    assert np.array_equal(S.indptr, rp) and np.array_equal(S.indices, ci)
    close(S.data, ref)


@pytest.mark.parametrize("tiles", [None, (24, 40)])
@pytest.mark.parametrize("name,degree", CASES)
def test_csr_scalar_ops(oracle, name, degree, tiles):
    m, o = make(name, degree, oracle)
    rng = np.random.default_rng(3)
    coef = rng.random(o.ngauss) + 0.5
    _csr_check(m, o, oracle, 0, coef, o.laplace_fwd, o.laplace_bwd, 1, tiles)
    if m.dim == 2:      # 3-D mass keeps the reference's per-element COO layout (quirk Q5); CSR mode sums it the same way
        _csr_check(m, o, oracle, 1, coef, o.mass_fwd, o.mass_bwd, 1, tiles)


@pytest.mark.parametrize("name,degree", CASES)
def test_csr_scalar_tile_overlap(oracle, name, degree):
    """Option "tile_overlap": scalar tile forward with one barrier per tile (double-buffered local matrices, kernels.cuh k_tile_fwd_ov): bit-identical
    to the two-barrier kernel, against the oracle, for several tile sizes, CTA sizes and persistent launches that walk all tiles."""
    m, o = make(name, degree, oracle)
    m.set_option("structured", 0)
    rng = np.random.default_rng(12)
    coef = rng.random(o.ngauss) + 0.5
    for op, fn, ofwd in ((0, ops.compute_fem_laplace_matrix1, o.laplace_fwd), (1, ops.compute_fem_mass_matrix1, o.mass_fwd)):
        ind, vv = ofwd(coef)
        rp, ci, ref = oracle.canonical_csr(ind, vv, o.ndof)
        for rows, threads, glim in ((0, 0, 0), (16, 128, 1), (40, 320, 2), (7, 64, 3)):
            m.set_option("rows_per_tile", rows)
            m.set_option("tile_threads", threads)
            m.set_option("grid_limit", glim)
            m.set_option("tile_overlap", 0)
            base = fn(dev(coef), m, mode="csr").values.cpu().numpy()
            m.set_option("tile_overlap", 1)
            got = fn(dev(coef), m, mode="csr").values.cpu().numpy()
            close(got, ref)
            assert np.array_equal(got, base)
    for k_, v_ in (("tile_overlap", 0), ("rows_per_tile", 0), ("tile_threads", 0), ("grid_limit", 0)):
        m.set_option(k_, v_)


@pytest.mark.parametrize("name,degree", CASES)
def test_csr_mass_3d(oracle, name, degree):
    if MESHES[name][0] != 3:
        pytest.skip("3-D only")
    m, o = make(name, degree, oracle)
    rng = np.random.default_rng(4)
    coef = rng.random(o.ngauss) + 0.5
    ind, vv = o.mass_fwd(coef)                                         # N = ne*d*d slots
    rp, ci, ref = oracle.canonical_csr(ind, vv, o.ndof)
    k = dev(coef).requires_grad_(True)
    T = ops.compute_fem_mass_matrix1(k, m, mode="csr")
    close(T.values.detach().cpu().numpy(), ref)
    dv = rng.standard_normal(len(ref))
    expect = o.mass_bwd(oracle.csr_adjoint_to_slots(rp, ci, dv, ind, o.ndof))
    (g,) = torch.autograd.grad(T.values, k, dev(dv))
    close(g.cpu().numpy(), expect)
    Sp = ops.compute_fem_mass_matrix1(k, m, mode="coo")               # compat layout
    assert np.array_equal(Sp.indices.cpu().numpy(), ind)
    close(Sp.values.detach().cpu().numpy(), vv)
    gv = rng.standard_normal(len(vv))
    (g2,) = torch.autograd.grad(Sp.values, k, dev(gv))
    close(g2.cpu().numpy(), o.mass_bwd(gv))


@pytest.mark.parametrize("name,degree", [("tri_struct", 1), ("tri_unstruct", 1), ("tri_unstruct", 2), ("tet", 1)])
def test_csr_stiffness(oracle, name, degree):
    m, o = make(name, degree, oracle)
    ns = 3 if m.dim == 2 else 6
    rng = np.random.default_rng(6)
    H = rng.random((o.ngauss, ns, ns))
    _csr_check(m, o, oracle, 2, H, o.stiffness_fwd, o.stiffness_bwd, m.dim, None)


def test_csr_stiffness_p2_tets_is_refused_cleanly(oracle):
    """P2 tetrahedral elasticity (a 30 x 30 local matrix per element; the reference has no 3-D elasticity operator at all, extension N2) works in
    the COO-compatible mode (test_coo_stiffness[tet-2]); the CSR tile kernels cannot hold the elements of a single vertex row in shared memory, and
    the library says so instead of producing anything."""
    m, o = make("tet", 2, oracle)
    H = np.random.default_rng(6).random((o.ngauss, 6, 6))
    with pytest.raises(A.AdfemError, match="tile too large"):
        ops.compute_fem_stiffness_matrix(dev(H), m, mode="csr")


# ------------------------------------------------------------------------------------------ legacy symbols
def test_legacy_symbols_2d(oracle):
    """The reference's own ccall sequence (src/MFEM/MFEM.jl:93-106, MCore.jl:77-82,112-120) against the oracle."""
    L = A._lib.lib()
    c, e = meshgen.jitter_unstructured(12, 9, 0.1, seed=9)
    o = oracle.Mesh2D(c, e, degree=2)
    c3 = np.zeros((c.shape[0], 3)); c3[:, :2] = c
    e32 = np.ascontiguousarray(e, dtype=np.int32)
    ned = C.c_longlong(0)
    p = L.init_nnfem_mesh(c3.ctypes.data_as(A._lib.c_dp), C.c_int(c.shape[0]), e32.ctypes.data_as(A._lib.c_ip), C.c_int(e.shape[0]),
                          C.c_int(4), C.c_int(6), C.c_int(2), C.byref(ned))
    edges = np.ctypeslib.as_array(p, shape=(2 * ned.value,)).copy()
    C.CDLL(None).free(C.cast(p, C.c_void_p))                           # caller frees (own=true)
    assert np.array_equal(edges.reshape(2, -1).T - 1, o.edges)
    assert L.mfem_get_ngauss() == o.ngauss and L.mfem_get_ndof() == o.ndof and L.mfem_get_elem_ndof() == 6
    conn = np.zeros(o.nelem * 6, dtype=np.int64); L.mfem_get_connectivity(conn.ctypes.data_as(A._lib.c_lp))
    assert np.array_equal(conn.reshape(-1, 6) - 1, o.conn)
    x, y = np.zeros(o.ngauss), np.zeros(o.ngauss); L.mfem_get_gauss(x.ctypes.data_as(A._lib.c_dp), y.ctypes.data_as(A._lib.c_dp))
    assert np.array_equal(np.stack([x, y], 1), o.gauss)
    rng = np.random.default_rng(7)
    kappa = rng.random(o.ngauss) + 1
    N = o.ngauss * 36
    ind, vv = np.zeros(2 * N, dtype=np.int64), np.zeros(N)
    L.FemLaplaceScalar_forward_Julia(ind.ctypes.data_as(A._lib.c_lp), vv.ctypes.data_as(A._lib.c_dp), kappa.ctypes.data_as(A._lib.c_dp))
    oi, ov = o.laplace_fwd(kappa)
    assert np.array_equal(ind.reshape(N, 2), oi)
    close(vv, ov)
    gk = np.zeros(o.ngauss)
    L.FemLaplaceScalar_backward(gk.ctypes.data_as(A._lib.c_dp), ov.ctypes.data_as(A._lib.c_dp), None, None, None)
    close(gk, o.laplace_bwd(ov))
    rhs = np.zeros(o.ndof)
    L.FemSourceScalar_forward_Julia(rhs.ctypes.data_as(A._lib.c_dp), kappa.ctypes.data_as(A._lib.c_dp))
    close(rhs, o.source_fwd(kappa))
    H = rng.random(9 * o.ngauss)
    N2 = o.ngauss * 144
    ind2, vv2 = np.zeros(2 * N2, dtype=np.int64), np.zeros(N2)
    L.ComputeFemStiffnessMatrixMfem_forward_Julia(ind2.ctypes.data_as(A._lib.c_lp), vv2.ctypes.data_as(A._lib.c_dp), H.ctypes.data_as(A._lib.c_dp))
    oi2, ov2 = o.stiffness_fwd(H)
    assert np.array_equal(ind2.reshape(N2, 2), oi2)
    close(vv2, ov2)


def test_legacy_symbols_3d(oracle):
    L = A._lib.lib()
    c, e = meshgen.tet_grid(3, 3, 3, 0.3)
    o = oracle.Mesh3D(c, e, degree=1)
    e32 = np.ascontiguousarray(e, dtype=np.int32)
    cc = np.ascontiguousarray(c)
    ned = C.c_longlong(0)
    p = L.init_nnfem_mesh3(cc.ctypes.data_as(A._lib.c_dp), C.c_int(c.shape[0]), e32.ctypes.data_as(A._lib.c_ip), C.c_int(e.shape[0]),
                           C.c_int(2), C.c_int(1), C.byref(ned))
    C.CDLL(None).free(C.cast(p, C.c_void_p))
    assert ned.value == o.nedge and L.mfem_get_ngauss3() == o.ngauss
    kappa = np.random.default_rng(8).random(o.ngauss) + 1
    N = o.ngauss * 16
    ind, vv = np.zeros(2 * N, dtype=np.int64), np.zeros(N)
    L.FemLaplaceScalarT_forward_Julia(ind.ctypes.data_as(A._lib.c_lp), vv.ctypes.data_as(A._lib.c_dp), kappa.ctypes.data_as(A._lib.c_dp))
    oi, ov = o.laplace_fwd(kappa)
    assert np.array_equal(ind.reshape(N, 2), oi)
    close(vv, ov)
    rhs = np.zeros(o.ndof)
    L.FemSourceScalarT_forward_Julia(rhs.ctypes.data_as(A._lib.c_dp), kappa.ctypes.data_as(A._lib.c_dp))
    close(rhs, o.source_fwd(kappa))


# ------------------------------------------------------------------------------------------ BASELINE config 1
def test_config1_readme_poisson(oracle):
    """README Poisson forward pieces + κ-gradient of the gradtest loss Σ vv² on twoholes_large (65 664 COO slots)."""
    d = np.load(os.path.join(HERE, "golden", "twoholes_large.npz"))
    m, o = A.Mesh(d["nodes"], d["elems"], 2, 1, 2), oracle.Mesh2D(d["nodes"], d["elems"], 2, 1, 2)
    xy = A.gauss_nodes(m)
    assert np.array_equal(xy, o.gauss)
    kappa = np.sin(xy[:, 0]) * (1 + xy[:, 1] ** 2) + 1.0
    f = 1e5 * (xy[:, 0] + xy[:, 1])
    k = dev(kappa).requires_grad_(True)
    K = ops.compute_fem_laplace_matrix1(k, m)
    assert K.values.numel() == 65664
    oi, ov = o.laplace_fwd(kappa)
    assert np.array_equal(K.indices.cpu().numpy(), oi)
    close(K.values.detach().cpu().numpy(), ov)
    loss = (K.values ** 2).sum()
    (g,) = torch.autograd.grad(loss, k)
    close(g.cpu().numpy(), o.laplace_bwd(2 * ov))
    close(ops.compute_fem_source_term1(dev(f), m).cpu().numpy(), o.source_fwd(f))
    Kc = ops.compute_fem_laplace_matrix1(kappa, m)
    rp, ci, ref = oracle.canonical_csr(oi, ov, o.ndof)
    assert np.array_equal(Kc.indptr, rp) and np.array_equal(Kc.indices, ci)
    close(Kc.data, ref)


# ------------------------------------------------------------------------------------------ size-independent properties
def test_large_mesh_properties():
    """At a size the oracle would not finish quickly: K·1 = 0, Σ M = area, linearity in κ, adjoint identity
    <dK, K(κ)> = <κ, K^T(dK)>, and determinism (bit-identical repeat)."""
    import scipy.sparse as sp
    n = 1536
    m = A.Mesh(n, n, 1.0 / n)
    G = m.ngauss
    gen = torch.Generator(device="cuda").manual_seed(0)
    k1 = torch.rand(G, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    k2 = torch.rand(G, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    rowptr, colind = m.csr_pattern(1)
    v1 = ops.compute_fem_laplace_matrix1(k1, m, mode="csr").values
    v2 = ops.compute_fem_laplace_matrix1(k2, m, mode="csr").values
    v12 = ops.compute_fem_laplace_matrix1(k1 + 2 * k2, m, mode="csr").values
    assert torch.equal(v1, ops.compute_fem_laplace_matrix1(k1, m, mode="csr").values)
    assert (v12 - (v1 + 2 * v2)).abs().max().item() < 1e-12 * v12.abs().max().item()
    K = sp.csr_matrix((v1.cpu().numpy(), colind, rowptr), shape=(m.ndof, m.ndof))
    assert np.abs(K @ np.ones(m.ndof)).max() < 1e-10
    assert abs(K - K.T).max() < 1e-12
    Mv = ops.compute_fem_mass_matrix1(torch.ones(G, dtype=torch.float64, device="cuda"), m, mode="csr").values
    assert abs(Mv.sum().item() - 1.0) < 1e-11
    dK = torch.randn(v1.numel(), dtype=torch.float64, device="cuda", generator=gen)
    kk = k1.clone().requires_grad_(True)
    (g,) = torch.autograd.grad(ops.compute_fem_laplace_matrix1(kk, m, mode="csr").values, kk, dK)
    lhs, rhs = (dK * v1).sum().item(), (g * k1).sum().item()
    assert abs(lhs - rhs) < 1e-10 * max(abs(lhs), 1.0)
    rhs1 = ops.compute_fem_source_term1(torch.ones(G, dtype=torch.float64, device="cuda"), m)
    assert abs(rhs1.sum().item() - 1.0) < 1e-11


def test_empty_and_tiny_meshes(oracle):
    """Edge cases: a single element, and a mesh smaller than one tile / one warp."""
    c = np.array([[0.0, 0.0], [2.0, 0.0], [0.0, 1.0]])
    for e in (np.array([[0, 1, 2]]), np.array([[1, 0, 2]])):          # second one is clockwise -> swapped
        for deg in (1, 2):
            m, o = A.Mesh(c, e, degree=deg), oracle.Mesh2D(c, e, degree=deg)
            kappa = np.arange(1, o.ngauss + 1, dtype=float)
            ind, vv = o.laplace_fwd(kappa)
            rp, ci, ref = oracle.canonical_csr(ind, vv, o.ndof)
            S = ops.compute_fem_laplace_matrix1(kappa, m)
            assert np.array_equal(S.indptr, rp) and np.array_equal(S.indices, ci)
            close(S.data, ref)


@pytest.mark.parametrize("degree", [1, 2])
def test_device_resident_csr_pattern(oracle, degree):
    """adfem_csr_pattern_device hands out the handle's device copies of rowptr / colind (solver hand-off): bit-identical to the
    ORACLE's canonical CSR pattern (sorted, duplicates summed, from the reference's COO output), and together with adfem_assemble_csr's
    values a complete device CSR matrix whose dense form equals the oracle's; a device-side solve with it reproduces the oracle's solution."""
    import ctypes as C

    class _Dev:                                    # wrap a raw device pointer for torch (CUDA array interface)
        def __init__(self, ptr, n, typestr):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}

    c, e = meshgen.jitter_unstructured(13, 11, 0.1, seed=9)
    m, o = A.Mesh(c, e, degree=degree), oracle.Mesh2D(c, e, degree=degree)
    rowptr, colind = m.csr_pattern(1)
    prp, pci = C.c_void_p(), C.c_void_p()
    A._lib.check(A._lib.lib().adfem_csr_pattern_device(m.handle, C.byref(prp), C.byref(pci)))
    d_rp = torch.as_tensor(_Dev(prp.value, len(rowptr), "<i8"), device="cuda")
    d_ci = torch.as_tensor(_Dev(pci.value, len(colind), "<i4"), device="cuda")
    assert np.array_equal(d_rp.cpu().numpy(), rowptr) and np.array_equal(d_ci.cpu().numpy(), colind)
    k = torch.rand(m.ngauss, dtype=torch.float64, device="cuda") + 0.5
    ind, vv = o.laplace_fwd(k.cpu().numpy())
    rp, ci, ref = oracle.canonical_csr(ind, vv, o.ndof)
    assert np.array_equal(d_rp.cpu().numpy(), rp) and np.array_equal(d_ci.cpu().numpy().astype(np.int64), ci)          # device pattern == oracle pattern
    T = ops.compute_fem_laplace_matrix1(k, m, mode="csr")
    K = torch.sparse_csr_tensor(d_rp, d_ci.to(torch.int64), T.values, size=(m.ndof, m.ndof))
    import scipy.sparse as sps
    Kref = sps.csr_matrix((ref, ci, rp), shape=(o.ndof, o.ndof)).toarray()
    close(K.to_dense().cpu().numpy(), Kref)
    # solver hand-off: K + M is SPD; solve on the device from the device CSR, compare with the oracle matrices solved on the host
    Mv = ops.compute_fem_mass_matrix1(torch.ones_like(k), m, mode="csr").values
    indm, vm = o.mass_fwd(np.ones(o.ngauss))
    _, _, refm = oracle.canonical_csr(indm, vm, o.ndof)
    Aref = Kref + sps.csr_matrix((refm, ci, rp), shape=(o.ndof, o.ndof)).toarray()
    b = np.random.default_rng(0).standard_normal(o.ndof)
    Ad = torch.sparse_csr_tensor(d_rp, d_ci.to(torch.int64), T.values + Mv, size=(m.ndof, m.ndof)).to_dense()
    x = torch.linalg.solve(Ad, torch.from_numpy(b).cuda())
    close(x.cpu().numpy(), np.linalg.solve(Aref, b), rel=1e-9)
