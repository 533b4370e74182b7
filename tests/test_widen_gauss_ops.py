"""GPU parity tests of the widening rows and of the opt-in kernels written at the end of round 1: Gauss-point operators and
matrix-free terms (csrc/gauss_ops.cu), structured Q1 siblings, coef_presum, the fused constitutive step, the structured-mesh elasticity kernels
(grid_elast.cuh, tet_grid.cuh), the structured scatter kernels and row_gather — against the oracle, through the C ABI (handle API via the Python
mirror, and the reference's legacy host-pointer symbols).

First run on a B200 in round 2 (profiles/pytest_gpu_r02_first.log: 34 passed); part of the normal `-m gpu` suite since."""
import ctypes as C
import os

import numpy as np
import pytest

import adfem_jl_b200 as A
from adfem_jl_b200 import meshgen

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def close(a, b, rel=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    scale = max(np.abs(b).max(), 1e-300) if b.size else 1.0
    err = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), scale * 1e-3)
    assert err.size == 0 or err.max() <= rel, f"max rel err {err.max():.3e}"


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).cuda()


def npy(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("degree", [1, 2])
def test_gauss_ops_2d(oracle, degree):
    c, e = meshgen.jitter_unstructured(23, 17, 0.05, seed=4)
    m, o = A.Mesh(c, e, degree=degree), oracle.Mesh2D(c, e, degree=degree)
    rng = np.random.default_rng(degree)
    G, n, nv = o.ngauss, o.ndof, o.nnode
    u, u2, sig = rng.standard_normal(n), rng.standard_normal(2 * n), rng.standard_normal((G, 3))
    cases = [   # public function, input, oracle forward, oracle adjoint
        (A.fem_to_gauss_points, u, o.fem_to_gauss_fwd, lambda w: np.concatenate([o.fem_to_gauss_bwd(w), np.zeros(n - nv)])),
        (A.dof_to_gauss_points, u, o.dof_to_gauss_fwd, o.dof_to_gauss_bwd),
        (A.eval_grad_on_gauss_pts1, u, o.grad_fwd, o.grad_bwd),
        (A.eval_strain_on_gauss_pts, u2, o.strain_fwd, o.strain_bwd),
        (A.compute_strain_energy_term, sig, o.strain_energy_fwd, lambda w: o.strain_energy_bwd(w).reshape(G, 3)),
    ]
    for fn, x, fwd, bwd in cases:
        xt = dev(x).requires_grad_(True)
        out = fn(xt, m)
        close(npy(out).reshape(-1), fwd(x.reshape(-1)))
        w = rng.standard_normal(tuple(out.shape))
        (g,) = torch.autograd.grad(out, xt, dev(w))
        close(npy(g), bwd(w.reshape(-1)))
        close(fn(x, m).reshape(-1), fwd(x.reshape(-1)))          # numpy in -> numpy out (the reference's eager methods)


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (3, 1), (3, 2)])
def test_laplace_term(oracle, dim, degree):
    rng = np.random.default_rng(10 * dim + degree)
    if dim == 2:
        c, e = meshgen.jitter_unstructured(19, 14, 0.05, seed=5)
        m, o = A.Mesh(c, e, degree=degree), oracle.Mesh2D(c, e, degree=degree)
    else:
        c, e = meshgen.tet_grid(4, 4, 3, 0.25)
        c = c + rng.uniform(-0.03, 0.03, c.shape)
        m, o = A.Mesh3(c, e, degree=degree), oracle.Mesh3D(c, e, degree=degree)
    nu, u, go = rng.random(o.ngauss) + 0.5, rng.standard_normal(o.ndof), rng.standard_normal(o.ndof)
    ut, nt = dev(u).requires_grad_(True), dev(nu).requires_grad_(True)
    out = A.compute_fem_laplace_term1(ut, nt, m)
    close(npy(out), o.laplace_term_fwd(nu, u))
    gu, gnu = torch.autograd.grad(out, [ut, nt], dev(go))
    rnu, ru = o.laplace_term_bwd(go, nu, u)
    close(npy(gnu), rnu); close(npy(gu), ru)
    close(A.compute_fem_laplace_term1(u, nu, m), o.laplace_term_fwd(nu, u))
    # deps/MFEM/ComputeLaplaceTermMfem/ftest.jl:8-15: the term equals the assembled matrix applied to u
    K = A.compute_fem_laplace_matrix1(nu, m)
    close(npy(out), K @ u, rel=1e-10)


def test_gauss_ops_3d(oracle):
    """3-D transfers have no reference twin: checked against the oracle's shape tables and by transposition."""
    rng = np.random.default_rng(6)
    c, e = meshgen.tet_grid(4, 4, 3, 0.25)
    c = c + rng.uniform(-0.03, 0.03, c.shape)
    for degree in (1, 2):
        m, o = A.Mesh3(c, e, degree=degree), oracle.Mesh3D(c, e, degree=degree)
        h, hx, hy, hz = o.shape_tables()
        u = rng.standard_normal(o.ndof)
        ul = u[o.conn]
        close(A.dof_to_gauss_points(u, m), np.einsum("ed,edk->ek", ul, h).reshape(-1))
        close(A.eval_grad_on_gauss_pts1(u, m), np.stack([np.einsum("ed,edk->ek", ul, t) for t in (hx, hy, hz)], axis=2).reshape(-1, 3))
        for fn, nin in ((A.fem_to_gauss_points, o.nnode), (A.dof_to_gauss_points, o.ndof), (A.eval_grad_on_gauss_pts1, o.ndof),
                        (A.eval_strain_on_gauss_pts, 3 * o.ndof), (A.compute_strain_energy_term, (o.ngauss, 6))):
            x = dev(rng.standard_normal(nin)).requires_grad_(True)
            out = fn(x, m)
            w = dev(rng.standard_normal(tuple(out.shape)))
            (g,) = torch.autograd.grad(out, x, w)
            lhs, rhs = float((out * w).sum()), float((x * g).sum())
            assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), abs(rhs), 1.0)


def test_plane_matrices(oracle):
    rng = np.random.default_rng(8)
    N = 1000
    E, nu = rng.random(N) + 0.5, rng.random(N) * 0.45
    for mode, fn in ((0, A.compute_plane_strain_matrix), (1, A.compute_plane_stress_matrix)):
        Et, nt = dev(E).requires_grad_(True), dev(nu).requires_grad_(True)
        H = fn(Et, nt)
        close(npy(H), oracle.plane_matrix_fwd(E, nu, mode), rel=1e-14)
        g = rng.standard_normal((N, 3, 3))
        gE, gnu = torch.autograd.grad(H, [Et, nt], dev(g))
        rE, rnu = oracle.plane_matrix_bwd(g, E, nu, mode)
        close(npy(gE), rE, rel=1e-11); close(npy(gnu), rnu, rel=1e-11)
        close(fn(E, nu), oracle.plane_matrix_fwd(E, nu, mode), rel=1e-14)
    # fused pre-step + assembly: H(E, nu) feeds the stiffness operator and the gradient reaches E
    c, e = meshgen.jitter_unstructured(9, 8, 0.1, seed=1)
    m, o = A.Mesh(c, e), oracle.Mesh2D(c, e)
    E, nu = rng.random(o.ngauss) + 0.5, rng.random(o.ngauss) * 0.4
    Et = dev(E).requires_grad_(True)
    Kt = A.compute_fem_stiffness_matrix(A.compute_plane_stress_matrix(Et, dev(nu)), m, mode="coo")
    w = rng.standard_normal(Kt.values.numel())
    (gE,) = torch.autograd.grad(Kt.values, Et, dev(w))
    rE, _ = oracle.plane_matrix_bwd(o.stiffness_bwd(w), E, nu, 1)
    close(npy(gE), rE, rel=1e-11)


def test_legacy_symbols_gauss_ops(oracle):
    """The reference's eager ccall targets (src/MFEM/MUtils.jl:241-270, MCore.jl:740-750, 804-840) on the global 2-D mesh, incl. their
    accumulate-into-the-caller's-array behaviour."""
    L = A._lib.lib()
    c, e = meshgen.jitter_unstructured(10, 8, 0.1, seed=11)
    o = oracle.Mesh2D(c, e, degree=2)
    c3 = np.zeros((c.shape[0], 3)); c3[:, :2] = c
    e32 = np.ascontiguousarray(e, dtype=np.int32)
    ned = C.c_longlong(0)
    p = L.init_nnfem_mesh(c3.ctypes.data_as(A._lib.c_dp), C.c_int(c.shape[0]), e32.ctypes.data_as(A._lib.c_ip), C.c_int(e.shape[0]),
                          C.c_int(4), C.c_int(6), C.c_int(2), C.byref(ned))
    C.CDLL(None).free(C.cast(p, C.c_void_p))
    d = lambda a: a.ctypes.data_as(A._lib.c_dp)
    rng = np.random.default_rng(12)
    G, n = o.ngauss, o.ndof
    u, u2, nu, sig = rng.standard_normal(n), rng.standard_normal(2 * n), rng.random(G) + 0.5, rng.standard_normal(3 * G)
    out = np.full(G, 7.0); L.FemToGaussPointsMfem_Julia(d(out), d(u)); close(out, o.fem_to_gauss_fwd(u))            # assigns
    out = np.full(G, 7.0); L.DofToGaussPointsMfem_forward_Julia(d(out), d(u)); close(out, o.dof_to_gauss_fwd(u))   # assigns
    out = np.full(2 * G, 7.0); L.FemGradMfem_forward(d(out), d(u)); close(out, o.grad_fwd(u))
    base = rng.standard_normal(3 * G)
    out = base.copy(); L.EvalStrainOnGaussPts_forward_Julia(d(out), d(u2)); close(out, base + o.strain_fwd(u2))    # accumulates
    base = rng.standard_normal(2 * n)
    out = base.copy(); L.ComputeStrainEnergyTermMfem_forward_Julia(d(out), d(sig)); close(out, base + o.strain_energy_fwd(sig))
    base = rng.standard_normal(n)
    out = base.copy(); L.ComputeLaplaceTermMfem_forward_Julia(d(out), d(nu), d(u)); close(out, base + o.laplace_term_fwd(nu, u))
    go = rng.standard_normal(n)
    gnu, gu = np.zeros(G), np.zeros(n)
    L.ComputeLaplaceTermMfem_backward(d(gnu), d(gu), d(go), None, d(nu), d(u))
    rnu, ru = o.laplace_term_bwd(go, nu, u)
    close(gnu, rnu); close(gu, ru)
    w = rng.standard_normal(G)
    g = np.zeros(n); L.DofToGaussPointsMfem_backward(d(g), d(w), None, None); close(g, o.dof_to_gauss_bwd(w))
    g = np.zeros(n); L.FemToGaussPointsMfem_backward(d(g), d(w), None, None); close(g[:o.nnode], o.fem_to_gauss_bwd(w)); assert not g[o.nnode:].any()
    w2 = rng.standard_normal(2 * G)
    g = np.zeros(n); L.FemGradMfem_backward(d(g), d(w2), None, None); close(g, o.grad_bwd(w2))
    w3 = rng.standard_normal(3 * G)
    g = np.zeros(2 * n); L.EvalStrainOnGaussPts_backward(d(g), d(w3)); close(g, o.strain_bwd(w3))
    w4 = rng.standard_normal(2 * n)
    g = np.zeros(3 * G); L.ComputeStrainEnergyTermMfem_backward(d(g), d(w4)); close(g, o.strain_energy_bwd(w4))
    N = 50
    E, pr = rng.random(N) + 0.5, rng.random(N) * 0.45
    H = np.zeros(9 * N); L.PlaneStrainMatrix_forward(d(H), d(E), d(pr), C.c_int(N)); close(H.reshape(N, 3, 3), oracle.plane_matrix_fwd(E, pr, 0), rel=1e-14)
    H = np.zeros(9 * N); L.PlaneStressMatrix_forward(d(H), d(E), d(pr), C.c_int(N)); close(H.reshape(N, 3, 3), oracle.plane_matrix_fwd(E, pr, 1), rel=1e-14)
    gH = rng.standard_normal(9 * N)
    gn, gE = np.zeros(N), np.zeros(N)
    L.PlaneStressMatrix_backward(d(gn), d(gE), d(gH), d(E), d(pr), C.c_int(N))
    rE, rn = oracle.plane_matrix_bwd(gH, E, pr, 1)
    close(gE, rE, rel=1e-11); close(gn, rn, rel=1e-11)


def test_laplace_term_3d_legacy(oracle):
    L = A._lib.lib()
    c, e = meshgen.tet_grid(3, 3, 3, 0.3)
    o = oracle.Mesh3D(c, e, degree=1)
    e32, cc = np.ascontiguousarray(e, dtype=np.int32), np.ascontiguousarray(c)
    ned = C.c_longlong(0)
    p = L.init_nnfem_mesh3(cc.ctypes.data_as(A._lib.c_dp), C.c_int(c.shape[0]), e32.ctypes.data_as(A._lib.c_ip), C.c_int(e.shape[0]),
                           C.c_int(2), C.c_int(1), C.byref(ned))
    C.CDLL(None).free(C.cast(p, C.c_void_p))
    rng = np.random.default_rng(13)
    nu, u = rng.random(o.ngauss) + 0.5, rng.standard_normal(o.ndof)
    out = np.zeros(o.ndof)
    d = lambda a: a.ctypes.data_as(A._lib.c_dp)
    L.ComputeLaplaceTermMfem3_forward_Julia(d(out), d(nu), d(u))
    close(out, o.laplace_term_fwd(nu, u))


@pytest.mark.parametrize("dim", [2, 3])
def test_coef_presum_option(oracle, dim):
    """Option "coef_presum" (P1 elasticity: Gauss-summed coefficients in a streaming pre-pass, per-element gradient expanded afterwards)
    gives the same CSR values and H-gradient as the default path and the oracle, tiled and direct-gather adjoint, with and without
    coefficient staging."""
    rng = np.random.default_rng(20 + dim)
    if dim == 2:
        c, e = meshgen.jitter_unstructured(21, 16, 0.05, seed=6)
        m, o = A.Mesh(c, e), oracle.Mesh2D(c, e)
        ns = 3
    else:
        c, e = meshgen.tet_grid(4, 4, 4, 0.25)
        c = c + rng.uniform(-0.03, 0.03, c.shape)
        m, o = A.Mesh3(c, e), oracle.Mesh3D(c, e)
        ns = 6
    n = dim * o.ndof
    H = rng.random((o.ngauss, ns, ns))
    ind, vv = o.stiffness_fwd(H.reshape(-1))
    rp, ci, ref = oracle.canonical_csr(ind, vv, n)
    dv = rng.standard_normal(len(ref))
    expect = o.stiffness_bwd(oracle.csr_adjoint_to_slots(rp, ci, dv, ind, n))
    for presum, tiled, staged in ((1, 1, 1), (1, 0, 1), (1, 1, 0), (0, 1, 1)):
        m.set_option("coef_presum", presum)
        m.set_option("adjoint_tiled", tiled)
        m.set_option("coef_prefetch", staged)
        k = dev(H).requires_grad_(True)
        T = A.compute_fem_stiffness_matrix(k, m, mode="csr")
        assert np.array_equal(T.rowptr, rp) and np.array_equal(T.colind, ci)
        close(npy(T.values), ref)
        (g,) = torch.autograd.grad(T.values, k, dev(dv))
        close(npy(g).reshape(-1), expect)


def test_quad_scalar_siblings(oracle):
    """Structured Q1 FemLaplace / FemMass / FemSource through the reference-style signatures (coef, m, n, h)."""
    rng = np.random.default_rng(31)
    for m, n, h in ((13, 9, 0.1), (1, 1, 2.0), (64, 33, 1 / 64)):
        coef = rng.random(4 * m * n) + 0.5
        for fn, fwd, bwd in ((A.compute_fem_laplace_matrix1, oracle.quad_laplace_fwd, oracle.quad_laplace_bwd),
                             (A.compute_fem_mass_matrix1, oracle.quad_mass_fwd, oracle.quad_mass_bwd)):
            ri, rj, rv = fwd(coef, m, n, h)
            k = dev(coef).requires_grad_(True)
            S = fn(k, m, n, h)
            idx = npy(S.indices)
            assert np.array_equal(idx[:, 0], ri) and np.array_equal(idx[:, 1], rj)        # 0-based, bit-exact
            close(npy(S.values), rv)
            g = rng.standard_normal(len(rv))
            (gk,) = torch.autograd.grad(S.values, k, dev(g))
            close(npy(gk), bwd(g, m, n, h))
            import scipy.sparse as sp
            N = (m + 1) * (n + 1)
            ref = sp.coo_matrix((rv, (ri, rj)), shape=(N, N)).tocsr()
            assert abs(fn(coef, m, n, h) - ref).max() <= 1e-12 * abs(ref).max()             # numpy in -> scipy out
        f = rng.standard_normal(4 * m * n)
        ft = dev(f).requires_grad_(True)
        rhs = A.compute_fem_source_term1(ft, m, n, h)
        close(npy(rhs), oracle.quad_source_fwd(f, m, n, h))
        w = rng.standard_normal((m + 1) * (n + 1))
        (gf,) = torch.autograd.grad(rhs, ft, dev(w))
        close(npy(gf), oracle.quad_source_bwd(w, m, n, h))
        close(A.compute_fem_source_term1(f, m, n, h), oracle.quad_source_fwd(f, m, n, h))


@pytest.mark.parametrize("plane,mode", [("strain", 0), ("stress", 1)])
def test_fused_plane_stiffness(oracle, plane, mode):
    """adfem_assemble_csr_plane[_adjoint]: same CSR values / (E, nu)-gradients as plane matrix -> stiffness through the oracle."""
    rng = np.random.default_rng(50 + mode)
    c, e = meshgen.jitter_unstructured(17, 13, 0.05, seed=8)
    m, o = A.Mesh(c, e), oracle.Mesh2D(c, e)
    E, nu = rng.random(o.ngauss) + 0.5, rng.random(o.ngauss) * 0.4
    n = 2 * o.ndof
    H = oracle.plane_matrix_fwd(E, nu, mode)
    ind, vv = o.stiffness_fwd(H.reshape(-1))
    rp, ci, ref = oracle.canonical_csr(ind, vv, n)
    Et, nt = dev(E).requires_grad_(True), dev(nu).requires_grad_(True)
    T = A.compute_fem_stiffness_matrix_from_moduli(Et, nt, m, plane=plane)
    assert np.array_equal(T.rowptr, rp) and np.array_equal(T.colind, ci)
    close(npy(T.values), ref)
    dv = rng.standard_normal(len(ref))
    gE, gnu = torch.autograd.grad(T.values, [Et, nt], dev(dv))
    rE, rnu = oracle.plane_matrix_bwd(o.stiffness_bwd(oracle.csr_adjoint_to_slots(rp, ci, dv, ind, n)), E, nu, mode)
    close(npy(gE), rE, rel=1e-11); close(npy(gnu), rnu, rel=1e-11)
    m2 = A.Mesh(c, e, degree=2)
    with pytest.raises(A.AdfemError):                                                                          # P1 only
        A.compute_fem_stiffness_matrix_from_moduli(dev(np.ones(m2.ngauss)), dev(np.full(m2.ngauss, 0.3)), m2)


@pytest.mark.parametrize("mapped", [False, True])
@pytest.mark.parametrize("m,n", [(37, 29), (5, 3), (130, 21)])
def test_structured_elasticity_kernels(oracle, m, n, mapped):
    """Option "structured_elasticity": index-free P1 elasticity kernels (csrc/grid_elast.cuh) on Mesh(m, n, h) against the oracle and against
    the general tile kernels, several rows-per-warp settings (one or many chunks), both area formulas; `mapped`: the same connectivity on
    smoothly mapped + jittered node positions (MAPPED instantiations, positions from the coordinate array)."""
    rng = np.random.default_rng(m + n)
    c, e = meshgen.tri_grid(m, n, 0.05)
    if mapped:
        c = np.stack([c[:, 0] + 0.004 * np.sin(14.0 * c[:, 1]), c[:, 1] + 0.003 * np.cos(18.0 * c[:, 0])], 1) + rng.uniform(-0.005, 0.005, c.shape)
    ms, o = A.Mesh(c, e), oracle.Mesh2D(c, e)
    assert A._lib.lib().adfem_mesh_info(ms.handle, A._lib.INFO_STRUCTURED) == (3 if mapped else 1)
    N2 = 2 * o.ndof
    H = rng.random((o.ngauss, 3, 3)) + 0.1
    ind, vv = o.stiffness_fwd(H.reshape(-1))
    rp, ci, ref = oracle.canonical_csr(ind, vv, N2)
    dv = rng.standard_normal(len(ref))
    expect = o.stiffness_bwd(oracle.csr_adjoint_to_slots(rp, ci, dv, ind, N2))
    for on, rows, heron in ((1, 0, 0), (1, 1, 1), (1, 7, 0), (1, 1000, 0), (0, 0, 0)):
        ms.set_option("structured_elasticity", on)
        ms.set_option("grid_rows", rows)
        ms.set_option("area_formula_csr", heron)
        k = dev(H).requires_grad_(True)
        T = A.compute_fem_stiffness_matrix(k, ms, mode="csr")
        assert np.array_equal(T.rowptr, rp) and np.array_equal(T.colind, ci)
        close(npy(T.values), ref)
        (g,) = torch.autograd.grad(T.values, k, dev(dv))
        close(npy(g).reshape(-1), expect)
    ms.set_option("grid_rows", 0)
    ms.set_option("area_formula_csr", 0)
    # fused constitutive step on the structured kernels (tangents from the moduli, gradients with respect to the moduli)
    E, nu = rng.random(o.ngauss) + 0.5, rng.random(o.ngauss) * 0.4
    for mode, plane in ((0, "strain"), (1, "stress")):
        _, vvm = o.stiffness_fwd(oracle.plane_matrix_fwd(E, nu, mode).reshape(-1))
        _, _, refm = oracle.canonical_csr(ind, vvm, N2)
        rE, rnu = oracle.plane_matrix_bwd(expect, E, nu, mode)
        for on in (1, 0):
            ms.set_option("structured_elasticity", on)
            Et, nt = dev(E).requires_grad_(True), dev(nu).requires_grad_(True)
            T = A.compute_fem_stiffness_matrix_from_moduli(Et, nt, ms, plane=plane)
            close(npy(T.values), refm)
            gE, gnu = torch.autograd.grad(T.values, [Et, nt], dev(dv))
            close(npy(gE), rE, rel=1e-11); close(npy(gnu), rnu, rel=1e-11)


@pytest.mark.parametrize("mapped", [False, True])
def test_structured_scatter_operators(oracle, mapped):
    """Scatter-type Gauss-point operators and the Laplace term on Mesh(m, n, h): the index-free one-thread-per-node kernels (grid_gauss.cuh,
    option "structured" = 1, the default) against the oracle and against the general adjacency-walking kernels ("structured" = 0; on the host the
    two bodies are bit-identical, tests/test_host_emulation.py).  `mapped`: the same connectivity on smoothly mapped + jittered node positions
    (corner positions from the coordinate array instead of the two axis arrays)."""
    rng = np.random.default_rng(61)
    m_, n_ = 45, 31
    c, e = meshgen.tri_grid(m_, n_, 0.05)
    if mapped:
        c = np.ascontiguousarray(np.stack([c[:, 0] + 0.004 * np.sin(14.0 * c[:, 1]), c[:, 1] + 0.003 * np.cos(18.0 * c[:, 0])], 1)
                                 + rng.uniform(-0.005, 0.005, c.shape))
    ms, o = A.Mesh(c, e), oracle.Mesh2D(c, e)
    assert A._lib.lib().adfem_mesh_info(ms.handle, A._lib.INFO_STRUCTURED) == (3 if mapped else 1)
    G, nd = o.ngauss, o.ndof
    nu, u, go, sig = rng.random(G) + 0.5, rng.standard_normal(nd), rng.standard_normal(nd), rng.standard_normal((G, 3))
    w1, w2, w3 = rng.standard_normal(G), rng.standard_normal((G, 2)), rng.standard_normal((G, 3))
    res = {}
    for on in (1, 0):
        ms.set_option("structured", on)
        ut, nt = dev(u).requires_grad_(True), dev(nu).requires_grad_(True)
        term = A.compute_fem_laplace_term1(ut, nt, ms)
        gu, gnu = torch.autograd.grad(term, [ut, nt], dev(go))
        se = A.compute_strain_energy_term(dev(sig), ms)
        adj = []
        for fn, x, w in ((A.fem_to_gauss_points, u, w1), (A.dof_to_gauss_points, u, w1), (A.eval_grad_on_gauss_pts1, u, w2),
                         (A.eval_strain_on_gauss_pts, np.concatenate([u, go]), w3)):
            xt = dev(x).requires_grad_(True)
            (g,) = torch.autograd.grad(fn(xt, ms), xt, dev(w))
            adj.append(npy(g))
        res[on] = [npy(term), npy(gu), npy(gnu), npy(se)] + adj
    ms.set_option("structured", 1)
    rnu, ru = o.laplace_term_bwd(go, nu, u)
    refs = [o.laplace_term_fwd(nu, u), ru, rnu, o.strain_energy_fwd(sig.reshape(-1)), o.fem_to_gauss_bwd(w1), o.dof_to_gauss_bwd(w1),
            o.grad_bwd(w2.reshape(-1)), o.strain_bwd(w3.reshape(-1))]
    for a, b, r in zip(res[1], res[0], refs):
        close(a, r)
        close(a, b, rel=1e-13)       # same summation order and geometry; only the compiler's FMA contraction may differ between the two kernels


@pytest.mark.parametrize("n,l,chunks,node", [(3, 2, 1, 1), (5, 4, 1, 1), (1, 1, 1, 1), (5, 6, 3, 1), (70, 5, 5, 1), (4, 4, 1, 0), (33, 34, 0, 1),
                                             (3, 2, 1, 2), (5, 4, 1, 2), (1, 1, 1, 2), (5, 6, 3, 2), (70, 5, 1, 2), (33, 34, 0, 2)])
def test_structured_tet_elasticity_forward(oracle, n, l, chunks, node):
    """Option "structured_elasticity" on Mesh3(n, n, l, h): Gauss-sum pre-pass + one warp per node for the forward, one warp per 32 tetrahedra
    for the adjoint (csrc/tet_grid.cuh); against the oracle and the general tile kernels."""
    rng = np.random.default_rng(70 + n + l)
    c, e = meshgen.tet_grid(n, n, l, 0.2)
    m, o = A.Mesh3(c, e), oracle.Mesh3D(c, e)
    assert A._lib.lib().adfem_mesh_info(m.handle, A._lib.INFO_STRUCTURED) == 2
    N3 = 3 * o.ndof
    H = rng.random((o.ngauss, 6, 6)) + 0.1
    ind, vv = o.stiffness_fwd(H.reshape(-1))
    rp, ci, ref = oracle.canonical_csr(ind, vv, N3)
    dv = rng.standard_normal(len(ref))
    expect = o.stiffness_bwd(oracle.csr_adjoint_to_slots(rp, ci, dv, ind, N3))
    m.set_option("tet_chunks", chunks)      # > 1: Gauss pre-sum and node kernel pipelined over z-chunks on two streams (0 = automatic)
    m.set_option("tet_node", node)          # 0: first-generation one-warp-per-node forward; 2: incidence list of the 32-tetrahedron parity split over two warps
    for on in ((1, 0) if n < 30 else (1,)):
        m.set_option("structured_elasticity", on)
        k = dev(H).requires_grad_(True)
        T = A.compute_fem_stiffness_matrix(k, m, mode="csr")
        assert np.array_equal(T.rowptr, rp) and np.array_equal(T.colind, ci)
        close(npy(T.values), ref)
        (g,) = torch.autograd.grad(T.values, k, dev(dv))
        close(npy(g).reshape(-1), expect)
        if on == 1 and chunks != 1:          # a second pass must find the side stream ordered behind the first one's readers
            T2 = A.compute_fem_stiffness_matrix(k, m, mode="csr")
            assert torch.equal(T2.values, T.values)


@pytest.mark.parametrize("dim,degree", [(2, 2), (2, 1), (3, 1), (3, 2)])
def test_row_gather_forward(oracle, dim, degree):
    """Option "row_gather": one-thread-per-row forward for Laplace and mass (csrc/row_gather.cuh) against the oracle; the rows per CTA adapt to the
    row lengths (128 for triangles and P1 tetrahedra, fewer for P2 tetrahedra)."""
    rng = np.random.default_rng(80 + 10 * dim + degree)
    if dim == 2:
        c, e = meshgen.jitter_unstructured(41, 37, 0.02, seed=12)
        m, o = A.Mesh(c, e, degree=degree), oracle.Mesh2D(c, e, degree=degree)
    else:
        c, e = meshgen.tet_grid(5, 5, 4, 0.2)
        c = c + rng.uniform(-0.02, 0.02, c.shape)
        m, o = A.Mesh3(c, e, degree=degree), oracle.Mesh3D(c, e, degree=degree)
    coef = rng.random(o.ngauss) + 0.5
    ind, vv = o.laplace_fwd(coef)
    rp, ci, ref = oracle.canonical_csr(ind, vv, o.ndof)
    mass_ref = None
    for on in (0, 1):
        m.set_option("row_gather", on)
        m.set_option("structured", 0)
        k = dev(coef).requires_grad_(True)
        T = A.compute_fem_laplace_matrix1(k, m, mode="csr")
        assert np.array_equal(T.rowptr, rp) and np.array_equal(T.colind, ci)
        close(npy(T.values), ref)
        M = npy(A.compute_fem_mass_matrix1(dev(coef), m, mode="csr").values)
        if mass_ref is None:
            mass_ref = M                                        # tile kernel (already checked against the oracle in test_gpu_parity.py)
        else:
            close(M, mass_ref)


@pytest.mark.parametrize("dim", [2, 3])
def test_row_gather_elasticity_forward(oracle, dim):
    """Option "row_gather" for P1 elasticity on unstructured meshes: Gauss-sum pre-pass + one thread per scalar row, no tile plan; the adjoint keeps
    the tile kernel.  Against the oracle."""
    rng = np.random.default_rng(90 + dim)
    if dim == 2:
        c, e = meshgen.jitter_unstructured(33, 29, 0.03, seed=13)
        m, o = A.Mesh(c, e), oracle.Mesh2D(c, e)
        ns = 3
    else:
        c, e = meshgen.tet_grid(5, 5, 4, 0.2)
        c = c + rng.uniform(-0.02, 0.02, c.shape)
        m, o = A.Mesh3(c, e), oracle.Mesh3D(c, e)
        ns = 6
    n = dim * o.ndof
    H = rng.random((o.ngauss, ns, ns)) + 0.1
    ind, vv = o.stiffness_fwd(H.reshape(-1))
    rp, ci, ref = oracle.canonical_csr(ind, vv, n)
    dv = rng.standard_normal(len(ref))
    expect = o.stiffness_bwd(oracle.csr_adjoint_to_slots(rp, ci, dv, ind, n))
    m.set_option("row_gather", 1)
    k = dev(H).requires_grad_(True)
    T = A.compute_fem_stiffness_matrix(k, m, mode="csr")
    assert np.array_equal(T.rowptr, rp) and np.array_equal(T.colind, ci)
    close(npy(T.values), ref)
    (g,) = torch.autograd.grad(T.values, k, dev(dv))
    close(npy(g).reshape(-1), expect)


def test_structured_tet_scatter_operators(oracle):
    """Scatter-type Gauss-point operators and the Laplace term on Mesh3(n, n, l, h): one-thread-per-node kernels with index arithmetic
    (csrc/tet_gauss.cuh, option "structured" = 1) against the general adjacency-walking kernels and, for the Laplace term, the oracle."""
    rng = np.random.default_rng(71)
    n_, l_ = 5, 4
    c, e = meshgen.tet_grid(n_, n_, l_, 0.2)
    m, o = A.Mesh3(c, e), oracle.Mesh3D(c, e)
    assert A._lib.lib().adfem_mesh_info(m.handle, A._lib.INFO_STRUCTURED) == 2
    G, nd = o.ngauss, o.ndof
    nu, u, go = rng.random(G) + 0.5, rng.standard_normal(nd), rng.standard_normal(nd)
    sig, w1, w3, w6 = rng.standard_normal((G, 6)), rng.standard_normal(G), rng.standard_normal((G, 3)), rng.standard_normal((G, 6))
    res = {}
    for on in (1, 0):
        m.set_option("structured", on)
        ut, nt = dev(u).requires_grad_(True), dev(nu).requires_grad_(True)
        term = A.compute_fem_laplace_term1(ut, nt, m)
        gu, gnu = torch.autograd.grad(term, [ut, nt], dev(go))
        se = A.compute_strain_energy_term(dev(sig), m)
        adj = []
        for fn, x, w in ((A.fem_to_gauss_points, u, w1), (A.dof_to_gauss_points, u, w1), (A.eval_grad_on_gauss_pts1, u, w3),
                         (A.eval_strain_on_gauss_pts, np.concatenate([u, go, u]), w6)):
            xt = dev(x).requires_grad_(True)
            (g,) = torch.autograd.grad(fn(xt, m), xt, dev(w))
            adj.append(npy(g))
        res[on] = [npy(term), npy(gu), npy(gnu), npy(se)] + adj
    m.set_option("structured", 1)
    rnu, ru = o.laplace_term_bwd(go, nu, u)
    close(res[1][0], o.laplace_term_fwd(nu, u)); close(res[1][1], ru); close(res[1][2], rnu)
    for a, b in zip(res[1], res[0]):
        close(a, b, rel=1e-13)


@pytest.mark.parametrize("n,l", [(4, 3), (1, 1)])
def test_structured_tet_scalar_operators(oracle, n, l):
    """Option "structured_tet_scalar" (opt-in: measured slower than the tile kernels) switches the scalar P1 operators on Mesh3(n, n, l, h) to the index-arithmetic kernels of
    csrc/tet_scalar.cuh (FemLaplaceScalarT, mass; forward and adjoint): against the oracle and the general tile kernels."""
    rng = np.random.default_rng(75 + n + l)
    c, e = meshgen.tet_grid(n, n, l, 0.2)
    m, o = A.Mesh3(c, e), oracle.Mesh3D(c, e)
    coef = rng.random(o.ngauss) + 0.5
    ind, vv = o.laplace_fwd(coef)
    rp, ci, ref = oracle.canonical_csr(ind, vv, o.ndof)
    dv = rng.standard_normal(len(ref))
    expect = o.laplace_bwd(oracle.csr_adjoint_to_slots(rp, ci, dv, ind, o.ndof))
    mass = {}
    for on in (1, 0):
        m.set_option("structured_tet_scalar", on)
        k = dev(coef).requires_grad_(True)
        T = A.compute_fem_laplace_matrix1(k, m, mode="csr")
        assert np.array_equal(T.rowptr, rp) and np.array_equal(T.colind, ci)
        close(npy(T.values), ref)
        (g,) = torch.autograd.grad(T.values, k, dev(dv))
        close(npy(g), expect)
        k2 = dev(coef).requires_grad_(True)
        M = A.compute_fem_mass_matrix1(k2, m, mode="csr")
        (gm,) = torch.autograd.grad(M.values, k2, dev(dv))
        mass[on] = (npy(M.values), npy(gm))
    close(mass[1][0], mass[0][0]); close(mass[1][1], mass[0][1])


@pytest.mark.parametrize("type_", [1, 2, 3])
def test_fused_svt_quad_stiffness(oracle, type_):
    """adfem_quad_stiffness1_svt[_grad]: SpatialVaryingTangentElastic fused into the structured compute_fem_stiffness_matrix1 == the two ops chained
    (on the device and through the oracle)."""
    rng = np.random.default_rng(95 + type_)
    m, n, h = 23, 17, 0.05
    mu = rng.random(4 * m * n * type_) + 0.5
    hmat = oracle.svt_fwd(mu, m, n, type_).reshape(4 * m * n, 2, 2)
    ri, rj, rv = oracle.univariate_stiffness_fwd(hmat, m, n, h)
    mt = dev(mu).requires_grad_(True)
    S = A.compute_fem_stiffness_matrix1_from_mu(mt, m, n, h, type_)
    idx = npy(S.indices)
    assert np.array_equal(idx[:, 0], ri - 1) and np.array_equal(idx[:, 1], rj - 1)
    close(npy(S.values), rv)
    g = rng.standard_normal(len(rv))
    (gm,) = torch.autograd.grad(S.values, mt, dev(g))
    close(npy(gm), oracle.svt_bwd(oracle.univariate_stiffness_bwd(g, m, n, h, True), m, n, type_))
    S2 = A.compute_fem_stiffness_matrix1(A.compute_space_varying_tangent_elasticity_matrix(dev(mu), m, n, h, type_), m, n, h)     # unfused
    close(npy(S.values), npy(S2.values), rel=1e-13)
