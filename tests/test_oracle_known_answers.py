"""Pins the CPU oracle against every known answer the reference's own tests hold for the path
 plus hand-derivable goldens and mathematical invariants.  CPU only."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import adfem_jl_b200  # noqa: F401  (package shim)
from adfem_jl_b200 import meshgen

HERE = os.path.dirname(os.path.abspath(__file__))


def csr(O, ind, vv, n):
    rp, ci, v = O.canonical_csr(ind, vv, n)
    return sp.csr_matrix((v, ci, rp), shape=(n, n))


# ----------------------------------------------------------------------------- reference test values
def test_mesh_2x2_reference_values(oracle):
    """test/MFEM2.jl:7-27 — nodes, ngauss == 24, area == 0.125."""
    c, e = meshgen.tri_grid(2, 2, 0.5)
    assert np.allclose(c, [[0, 0], [.5, 0], [1, 0], [0, .5], [.5, .5], [1, .5], [0, 1], [.5, 1], [1, 1]])
    M = oracle.Mesh2D(c, e)
    assert M.ngauss == 24
    assert np.allclose(M.area, np.ones(8) * 0.25 / 2, rtol=0, atol=1e-15)
    # quirk Q3: second triangle of every cell is clockwise in the input and gets v0<->v1 swapped
    assert M.elems[1].tolist() == [3, 1, 4] and M.elems[0].tolist() == [0, 1, 3]


def test_segment_rule_reference_values(oracle):
    """test/MFEM/MCore.jl:1-14 — lorder=6 gives the 4-point Gauss-Legendre rule, ascending."""
    p, w = oracle.segment_rule(6)
    assert len(p) == 4
    vals = 5.0 * (1 - p) + 6.0 * p
    assert np.allclose(vals, [5.069431844202973, 5.330009478207572, 5.669990521792428, 5.930568155797027], rtol=0, atol=1e-14)
    assert abs(w.sum() - 1) < 1e-15


def test_five_point_stencil_golden(oracle):
    """Hand-derivable golden (SURVEY §8c vi): P1 Laplace on UnitSquareMesh(8,8,'left') == Mesh(8,8,1/8)
    has interior rows [-1,-1,4,-1,-1] and structural zeros on the diagonal neighbours."""
    c, e = meshgen.tri_grid(8, 8, 1 / 8)
    M = oracle.Mesh2D(c, e)
    A = csr(oracle, *M.laplace_fwd(np.ones(M.ngauss)), M.ndof)
    i = 4 * 9 + 4
    row = A.getrow(i)
    cols = row.indices.tolist()
    assert cols == [i - 9, i - 8, i - 1, i, i + 1, i + 8, i + 9]
    dense = A.toarray()
    assert np.allclose(dense[i, [i - 9, i - 1, i, i + 1, i + 9]], [-1, -1, 4, -1, -1], atol=1e-13)
    assert abs(dense[i, i - 8]) < 1e-13 and abs(dense[i, i + 8]) < 1e-13     # 'left' diagonals: (i-8),(i+8) present, value 0
    assert A[i, i - 8] == 0 and (i - 8) in cols and (i + 8) in cols          # structural zeros kept
    assert np.allclose(dense, dense.T, atol=1e-13)


@pytest.mark.parametrize("degree", [1, 2])
def test_source_equals_mass_times_ones(oracle, degree):
    """deps/MFEM/FemSource1/ftest.jl:4-16."""
    c, e = meshgen.tri_grid(8, 8, 1 / 8)
    M = oracle.Mesh2D(c, e, degree=degree)
    rng = np.random.default_rng(0)
    coef = rng.random(M.ngauss)
    C = csr(oracle, *M.mass_fwd(coef), M.ndof)
    assert np.allclose(M.source_fwd(coef), C @ np.ones(M.ndof), rtol=0, atol=1e-15)


# ----------------------------------------------------------------------------- invariants / exactness
@pytest.mark.parametrize("degree", [1, 2])
def test_2d_invariants_and_exactness(oracle, degree):
    c, e = meshgen.jitter_unstructured(6, 5, 0.2, seed=5)
    M = oracle.Mesh2D(c, e, degree=degree)
    assert M.g == (3 if degree == 1 else 6) and M.elem_ndof == (3 if degree == 1 else 6)
    assert M.ndof == M.nnode + (0 if degree == 1 else M.nedge)
    assert M.nedge == M.nnode + M.nelem - 1                      # Euler, simply connected
    area = 6 * 5 * 0.04
    assert abs(M.area.sum() - area) < 1e-13 and abs(M.weights.sum() - area) < 1e-13
    one = np.ones(M.ngauss)
    K = csr(oracle, *M.laplace_fwd(one), M.ndof)
    Mm = csr(oracle, *M.mass_fwd(one), M.ndof)
    assert np.abs(K @ np.ones(M.ndof)).max() < 1e-12
    assert abs(Mm.sum() - area) < 1e-13
    assert abs(M.source_fwd(one).sum() - area) < 1e-13
    # nodal interpolant of a polynomial of the element degree is exact
    xy = np.zeros((M.ndof, 2))
    xy[:M.nnode] = M.coords
    if degree == 2:
        xy[M.nnode:] = 0.5 * (M.coords[M.edges[:, 0]] + M.coords[M.edges[:, 1]])
    x, y = xy[:, 0], xy[:, 1]
    if degree == 1:
        u = 2 * x - 3 * y + 1
        exact_grad2 = 13 * area
    else:
        u = x * x + 2 * x * y - y + 0.5
        # |grad u|^2 = (2x+2y)^2 + (2x-1)^2 integrated over [0,1.2]x[0,1.0]
        a, b = 1.2, 1.0
        exact_grad2 = (4 * (a ** 3 / 3 * b + 2 * (a ** 2 / 2) * (b ** 2 / 2) + a * b ** 3 / 3)
                       + (4 * a ** 3 / 3 - 2 * a ** 2 + a) * b)
    assert abs(u @ (K @ u) - exact_grad2) < 1e-11
    # Gauss points lie inside their element and reproduce linear fields
    gp = M.gauss.reshape(M.nelem, M.g, 2)
    cent = M.coords[M.elems].mean(1)
    assert np.allclose((gp * (M.weights.reshape(M.nelem, M.g, 1))).sum(1) / M.area[:, None], cent, atol=1e-13)


@pytest.mark.parametrize("degree", [1, 2])
def test_3d_invariants_and_exactness(oracle, degree):
    c, e = meshgen.tet_grid(3, 3, 2, 0.5)
    M = oracle.Mesh3D(c, e, degree=degree)
    assert M.g == (4 if degree == 1 else 11) and M.elem_ndof == (4 if degree == 1 else 10)
    vol = 1.5 * 1.5 * 1.0
    assert abs(M.volume.sum() - vol) < 1e-13 and abs(M.weights.sum() - vol) < 1e-13
    assert (M.volume > 0).all()
    one = np.ones(M.ngauss)
    K = csr(oracle, *M.laplace_fwd(one), M.ndof)
    Mm = csr(oracle, *M.mass_fwd(one), M.ndof)
    assert np.abs(K @ np.ones(M.ndof)).max() < 1e-12
    assert abs(Mm.sum() - vol) < 1e-12
    assert np.allclose(M.source_fwd(one), Mm @ np.ones(M.ndof), atol=1e-15)
    xyz = np.zeros((M.ndof, 3))
    xyz[:M.nnode] = M.coords
    if degree == 2:
        xyz[M.nnode:] = 0.5 * (M.coords[M.edges[:, 0]] + M.coords[M.edges[:, 1]])
    x, y, z = xyz.T
    if degree == 1:
        u = x - 2 * y + 3 * z
        exact = 14 * vol
    else:
        u = x * x + y * z
        # |grad u|^2 = 4x^2 + z^2 + y^2 over [0,1.5]^2 x [0,1]
        exact = 4 * (1.5 ** 3 / 3) * 1.5 * 1.0 + 1.5 * 1.5 * (1 / 3) + 1.5 * (1.5 ** 3 / 3) * 1.0
    assert abs(u @ (K @ u) - exact) < 1e-11
    # 3-D elasticity extension (N2): rigid-body modes are in the null space, matrix symmetric for symmetric H
    lam, mu = 1.3, 0.7
    H = np.zeros((6, 6))
    H[:3, :3] = lam
    H[np.arange(3), np.arange(3)] += 2 * mu
    H[np.arange(3, 6), np.arange(3, 6)] = mu
    S = csr(oracle, *M.stiffness_fwd(np.tile(H.reshape(-1), M.ngauss)), 3 * M.ndof)
    n = M.ndof
    modes = [np.r_[np.ones(n), np.zeros(2 * n)], np.r_[np.zeros(n), np.ones(n), np.zeros(n)], np.r_[-y, x, np.zeros(n)],
             np.r_[np.zeros(n), -z, y], np.r_[z, np.zeros(n), -x]]
    for r in modes:
        assert np.abs(S @ r).max() < 1e-11
    assert abs(S - S.T).max() < 1e-12
    # patch test: u = (x, 0, 0) has strain energy (lam+2mu)*vol
    ux = np.r_[x, np.zeros(2 * n)]
    assert abs(ux @ (S @ ux) - (lam + 2 * mu) * vol) < 1e-11


def test_2d_elasticity_invariants(oracle):
    c, e = meshgen.jitter_unstructured(4, 4, 0.25, seed=7)
    for degree in (1, 2):
        M = oracle.Mesh2D(c, e, degree=degree)
        E, nu = 2.0, 0.3
        H = E / ((1 + nu) * (1 - 2 * nu)) * np.array([[1 - nu, nu, 0], [nu, 1 - nu, 0], [0, 0, (1 - 2 * nu) / 2]])
        S = csr(oracle, *M.stiffness_fwd(np.tile(H.reshape(-1), M.ngauss)), 2 * M.ndof)
        n = M.ndof
        xy = np.zeros((n, 2))
        xy[:M.nnode] = M.coords
        if degree == 2:
            xy[M.nnode:] = 0.5 * (M.coords[M.edges[:, 0]] + M.coords[M.edges[:, 1]])
        for r in (np.r_[np.ones(n), np.zeros(n)], np.r_[np.zeros(n), np.ones(n)], np.r_[-xy[:, 1], xy[:, 0]]):
            assert np.abs(S @ r).max() < 1e-11
        ux = np.r_[xy[:, 0], np.zeros(n)]
        assert abs(ux @ (S @ ux) - H[0, 0] * 1.0) < 1e-11


# ----------------------------------------------------------------------------- gradients (gradtest.jl idea)
def _fd_check(fwd, bwd, x0, seed=0):
    """Directional derivative of L = sum(w * out) vs <bwd(w), v>; ops are linear in the coefficient so
    a single central difference is exact to rounding."""
    rng = np.random.default_rng(seed)
    out0 = fwd(x0)
    wgt = rng.standard_normal(out0.shape)
    v = rng.standard_normal(x0.shape)
    g = bwd(wgt)
    eps = 1e-3
    fd = (np.sum(wgt * fwd(x0 + eps * v)) - np.sum(wgt * fwd(x0 - eps * v))) / (2 * eps)
    assert abs(fd - g @ v) <= 1e-9 * max(1.0, abs(fd))


@pytest.mark.parametrize("degree", [1, 2])
def test_2d_adjoints(oracle, degree):
    c, e = meshgen.jitter_unstructured(3, 4, 0.3, seed=11)
    M = oracle.Mesh2D(c, e, degree=degree)
    rng = np.random.default_rng(1)
    _fd_check(lambda k: M.laplace_fwd(k)[1], M.laplace_bwd, rng.random(M.ngauss) + 1)
    _fd_check(lambda k: M.mass_fwd(k)[1], M.mass_bwd, rng.random(M.ngauss) + 1)
    _fd_check(lambda k: M.stiffness_fwd(k)[1], M.stiffness_bwd, rng.random(9 * M.ngauss))
    _fd_check(M.source_fwd, M.source_bwd, rng.random(M.ngauss))
    # PCL Jacobian (FemLaplaceScalar.h:65-92) is d vv / d kappa: column-major ngauss x N
    J = M.laplace_jacobian()
    k0 = rng.random(M.ngauss)
    assert np.allclose(J.T @ k0, M.laplace_fwd(k0)[1], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("degree", [1, 2])
def test_3d_adjoints(oracle, degree):
    c, e = meshgen.tet_grid(2, 2, 2, 0.5)
    M = oracle.Mesh3D(c, e, degree=degree)
    rng = np.random.default_rng(2)
    _fd_check(lambda k: M.laplace_fwd(k)[1], M.laplace_bwd, rng.random(M.ngauss) + 1)
    _fd_check(lambda k: M.mass_fwd(k)[1], M.mass_bwd, rng.random(M.ngauss) + 1)
    _fd_check(M.source_fwd, M.source_bwd, rng.random(M.ngauss))
    if degree == 1:
        _fd_check(lambda k: M.stiffness_fwd(k)[1], M.stiffness_bwd, rng.random(36 * M.ngauss))


# ----------------------------------------------------------------------------- Dirichlet
def test_dirichlet_matches_dense_julia_version(oracle):
    """test/mfem.jl:78-88 — COO op == dense slicing version (src/MFEM/MUtils.jl:184-199)."""
    rng = np.random.default_rng(3)
    N = 7
    A = rng.random((N, N))
    ii, jj = np.nonzero(A)
    order = rng.permutation(len(ii))       # arbitrary slot order, with duplicates below
    ii, jj = ii[order], jj[order]
    vv = A[ii, jj]
    ii = np.r_[ii, ii[:5]]
    jj = np.r_[jj, jj[:5]]
    vv = np.r_[vv * 1.0, vv[:5] * 0.0 + 0.25]
    A[ii[-5:], jj[-5:]] += 0.25
    bd = np.array([4, 1, 2])
    bdval = np.array([1.0, 2.0, 3.0])
    rhs = rng.random(N)
    oi, ov, orhs = oracle.impose_dirichlet_fwd(np.stack([ii, jj], 1), vv, bd, rhs, bdval)
    idx = np.ones(N, bool)
    idx[bd] = False
    r = rhs.copy()
    r[idx] = rhs[idx] - A[np.ix_(idx, bd)] @ bdval
    r[bd] = bdval
    B = np.zeros((N, N))
    B[np.ix_(idx, idx)] = A[np.ix_(idx, idx)]
    B[bd, bd] = 1.0
    got = sp.coo_matrix((ov, (oi[:, 0], oi[:, 1])), shape=(N, N)).toarray()
    assert np.allclose(got, B, atol=1e-14) and np.allclose(orhs, r, atol=1e-14)
    # appended diagonals come in ascending dof order (std::map iteration, quirk Q11)
    assert oi[-3:, 0].tolist() == [1, 2, 4] and np.all(ov[-3:] == 1.0)
    # adjoint by finite differences on a random linear functional of (ov, orhs)
    w1, w2 = rng.standard_normal(len(ov)), rng.standard_normal(N)
    gv, gr, gb = oracle.impose_dirichlet_bwd(w1, w2, np.stack([ii, jj], 1), vv, bd, bdval, N)

    def L(vv_, rhs_, bdval_):
        _, a, b = oracle.impose_dirichlet_fwd(np.stack([ii, jj], 1), vv_, bd, rhs_, bdval_)
        return w1 @ a + w2 @ b
    eps = 1e-6
    for (arr, g, which) in ((vv, gv, 0), (rhs, gr, 1), (bdval, gb, 2)):
        d = rng.standard_normal(arr.shape)
        args_p, args_m = [vv, rhs, bdval], [vv, rhs, bdval]
        args_p[which] = arr + eps * d
        args_m[which] = arr - eps * d
        fd = (L(*args_p) - L(*args_m)) / (2 * eps)
        assert abs(fd - g @ d) < 1e-7 * max(1, abs(fd))


# ----------------------------------------------------------------------------- structured-grid ops
def _julia_stiffness1(K, m, n, h):
    """src/Core.jl:98-131 restated (independent pure-Julia implementation in the reference)."""
    pts = [(-1 / np.sqrt(3) + 1) / 2, (1 / np.sqrt(3) + 1) / 2]
    Om = np.zeros((4, 4))
    for xi in pts:
        for eta in pts:
            B = np.array([[-1 / h * (1 - eta), 1 / h * (1 - eta), -1 / h * eta, 1 / h * eta],
                          [-1 / h * (1 - xi), -1 / h * xi, 1 / h * (1 - xi), 1 / h * xi]])
            Om += B.T @ K @ B * 0.25 * h * h
    A = np.zeros(((m + 1) * (n + 1),) * 2)
    for i in range(m):
        for j in range(n):
            kk = np.array([j * (m + 1) + i, j * (m + 1) + i + 1, (j + 1) * (m + 1) + i, (j + 1) * (m + 1) + i + 1])
            A[np.ix_(kk, kk)] += Om
    return A


def test_structured_ops(oracle):
    m, n, h = 4, 3, 0.5
    rng = np.random.default_rng(4)
    K = np.array([[2.0, 0.3], [0.3, 1.0]])
    N = (m + 1) * (n + 1)
    # constant K (Forward2) and the same K tiled per Gauss point (Forward_UFS) both equal the Julia version
    for hm in (K, np.tile(K, (4 * m * n, 1, 1))):
        ii, jj, vv = oracle.univariate_stiffness_fwd(hm, m, n, h)
        assert ii.min() == 1 and ii.max() == N          # 1-based like the op
        A = sp.coo_matrix((vv, (ii - 1, jj - 1)), shape=(N, N)).toarray()
        assert np.allclose(A, _julia_stiffness1(K, m, n, h), atol=1e-13)
    hm = rng.random((4 * m * n, 2, 2))
    _fd_check(lambda x: oracle.univariate_stiffness_fwd(x.reshape(-1, 2, 2), m, n, h)[2],
              lambda g: oracle.univariate_stiffness_bwd(g, m, n, h, True), hm.reshape(-1))
    _fd_check(lambda x: oracle.univariate_stiffness_fwd(x.reshape(2, 2), m, n, h)[2],
              lambda g: oracle.univariate_stiffness_bwd(g, m, n, h, False), K.reshape(-1))
    # elasticity: constant H (FemStiffness) == per-Gauss H tiled (SpatialFemStiffness) after summing Gauss points
    H = np.array([[3.0, 1.0, 0.2], [1.0, 2.5, 0.1], [0.2, 0.1, 0.8]])
    ii, jj, vv = oracle.fem_stiffness_fwd(H, m, n, h)
    A1 = sp.coo_matrix((vv, (ii - 1, jj - 1)), shape=(2 * N, 2 * N)).toarray()
    ii, jj, vv = oracle.spatial_stiffness_fwd(np.tile(H, (4 * m * n, 1, 1)), m, n, h)
    A2 = sp.coo_matrix((vv, (ii - 1, jj - 1)), shape=(2 * N, 2 * N)).toarray()
    assert np.allclose(A1, A2, atol=1e-12)
    x = np.tile(np.arange(m + 1) * h, n + 1)
    y = np.repeat(np.arange(n + 1) * h, m + 1)
    for r in (np.r_[np.ones(N), np.zeros(N)], np.r_[-y, x]):
        assert np.abs(A1 @ r).max() < 1e-12
    _fd_check(lambda xx: oracle.fem_stiffness_fwd(xx.reshape(3, 3), m, n, h)[2], lambda g: oracle.fem_stiffness_bwd(g, m, n, h),
              rng.random(9))
    _fd_check(lambda xx: oracle.spatial_stiffness_fwd(xx, m, n, h)[2], lambda g: oracle.spatial_stiffness_bwd(g, m, n, h),
              rng.random(36 * m * n))
    for t in (1, 2, 3):
        mu = rng.random(4 * m * n * t)
        hmat = oracle.svt_fwd(mu, m, n, t).reshape(-1, 2, 2)
        assert np.allclose(hmat[:, 0, 0], mu[:4 * m * n])
        assert np.allclose(hmat[:, 1, 1], mu[:4 * m * n] if t == 1 else mu[4 * m * n:8 * m * n])
        assert np.allclose(hmat[:, 0, 1], 0 if t < 3 else mu[8 * m * n:]) and np.allclose(hmat[:, 0, 1], hmat[:, 1, 0])
        _fd_check(lambda xx: oracle.svt_fwd(xx, m, n, t), lambda g: oracle.svt_bwd(g, m, n, t), mu)


def test_config1_twoholes_fixture(oracle):
    """BASELINE config 1 inputs: the README Poisson mesh; Euler characteristic of a domain with two holes."""
    d = np.load(os.path.join(HERE, "golden", "twoholes_large.npz"))
    M = oracle.Mesh2D(d["nodes"], d["elems"])
    assert M.nelem == 2432 and M.ngauss == 3 * 2432
    assert M.nnode - M.nedge + M.nelem == 1 - 2
    assert M.ngauss * 9 == 65664                      # COO slots quoted in SURVEY §8a
    K = csr(oracle, *M.laplace_fwd(np.sin(M.gauss[:, 0]) * (1 + M.gauss[:, 1] ** 2) + 1), M.ndof)
    assert np.abs(K @ np.ones(M.ndof)).max() < 1e-10


# ----------------------------------------------------------------------------- Gauss-point operators
@pytest.mark.parametrize("degree", [1, 2])
def test_gauss_point_operators_reference_test_ideas(oracle, degree):
    """The checks the reference's own scripts make for these ops, applied to the oracle restatements."""
    c, e = meshgen.tri_grid(10, 10, 0.1)
    M = oracle.Mesh2D(c, e, degree=degree)
    rng = np.random.default_rng(degree)
    n, G = M.ndof, M.ngauss
    pos = np.zeros((n, 2)); pos[:M.nnode] = c
    if degree == 2:
        pos[M.nnode:] = 0.5 * (c[M.edges[:, 0]] + c[M.edges[:, 1]])
    gx, gy = M.gauss[:, 0], M.gauss[:, 1]
    # deps/MFEM/ComputeLaplaceTermMfem/ftest.jl:8-15: term(u, nu) == K(nu) u
    nu, u = rng.random(G) + 0.5, rng.random(n)
    K = csr(oracle, *M.laplace_fwd(nu), n)
    assert np.abs(M.laplace_term_fwd(nu, u) - K @ u).max() < 1e-12 * np.abs(K @ u).max() * 100
    # deps/MFEM/ComputeStrainEnergyTermMfem/ftest.jl:6-18: strain_energy(K eps(u)) == Q(K) u
    u2 = rng.random(2 * n)
    Kc = np.array([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 0.5]])
    eps = M.strain_fwd(u2).reshape(G, 3)
    e1 = M.strain_energy_fwd((eps @ Kc.T).reshape(-1))
    Q = csr(oracle, *M.stiffness_fwd(np.tile(Kc.reshape(-1), G)), 2 * n)
    assert np.abs(e1 - Q @ u2).max() < 1e-11 * np.abs(Q @ u2).max()
    # deps/MFEM/FemGrad/TestFemGrad.jl:7-30: gradient of 3x + 14y is (3, 14) at every Gauss point
    g = M.grad_fwd(3 * pos[:, 0] + 14 * pos[:, 1]).reshape(G, 2)
    assert np.abs(g - [3.0, 14.0]).max() < 1e-11
    # deps/MFEM/FemToGaussPoints/TestFemToGaussPoints.jl:7-19: vertex interpolation reproduces linear functions at the Gauss points;
    # dof_to_gauss_points (all dofs) also reproduces 2x^2 + y for degree 2
    f_lin = 2 * pos[:, 0] - 0.5 * pos[:, 1] + 1
    assert np.abs(M.fem_to_gauss_fwd(f_lin) - (2 * gx - 0.5 * gy + 1)).max() < 1e-13
    assert np.abs(M.dof_to_gauss_fwd(f_lin) - (2 * gx - 0.5 * gy + 1)).max() < 1e-13
    if degree == 2:
        assert np.abs(M.dof_to_gauss_fwd(2 * pos[:, 0] ** 2 + pos[:, 1]) - (2 * gx ** 2 + gy)).max() < 1e-13
    # adjoints are the transposes (gradtest.jl of each op checks this by finite differences)
    for fwd, bwd, nin, nout in ((M.fem_to_gauss_fwd, M.fem_to_gauss_bwd, M.nnode, G), (M.dof_to_gauss_fwd, M.dof_to_gauss_bwd, n, G),
                                (M.grad_fwd, M.grad_bwd, n, 2 * G), (M.strain_fwd, M.strain_bwd, 2 * n, 3 * G),
                                (M.strain_energy_fwd, M.strain_energy_bwd, 3 * G, 2 * n)):
        x, w = rng.standard_normal(nin), rng.standard_normal(nout)
        xin = np.concatenate([x, np.zeros(n - nin)]) if nin == M.nnode and nin != n else x
        lhs, rhs = fwd(xin) @ w, x @ bwd(w)
        assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), abs(rhs), 1.0)
    go = rng.standard_normal(n)
    gnu, gu = M.laplace_term_bwd(go, nu, u)
    dnu, du = rng.standard_normal(G) * 1e-6, rng.standard_normal(n) * 1e-6
    fd = go @ (M.laplace_term_fwd(nu + dnu, u + du) - M.laplace_term_fwd(nu - dnu, u - du)) / 2
    assert abs(fd - (gnu @ dnu + gu @ du)) <= 1e-6 * abs(fd) + 1e-14


def test_laplace_term_3d_equals_matrix_action(oracle):
    """3-D twin (deps/MFEM3/ComputeLaplaceTermMfem): term(u, nu) == K(nu) u with K from FemLaplaceScalarT."""
    c, e = meshgen.tet_grid(3, 3, 3, 1 / 3)
    for degree in (1, 2):
        M = oracle.Mesh3D(c, e, degree=degree)
        rng = np.random.default_rng(3)
        nu, u = rng.random(M.ngauss) + 0.5, rng.random(M.ndof)
        K = csr(oracle, *M.laplace_fwd(nu), M.ndof)
        assert np.abs(M.laplace_term_fwd(nu, u) - K @ u).max() < 1e-11 * np.abs(K @ u).max()


def test_quad_scalar_siblings_known_answers(oracle):
    """Structured Q1 FemLaplace / FemMass / FemSource: invariants and an independent numpy assembly
    (the reference's own tests compare these kernels with the pure-Julia loops of src/Core.jl, test/invkernel.jl:149-160)."""
    rng = np.random.default_rng(4)
    m, n, h = 6, 4, 0.25
    N = (m + 1) * (n + 1)
    K, f = rng.random(4 * m * n) + 0.5, rng.standard_normal(4 * m * n)
    ii, jj, vv = oracle.quad_laplace_fwd(K, m, n, h)
    L = sp.coo_matrix((vv, (ii, jj)), shape=(N, N)).tocsr()
    assert np.abs(L @ np.ones(N)).max() < 1e-13                               # constants are in the kernel
    assert np.abs((L - L.T)).max() < 1e-15
    ii, jj, vv = oracle.quad_mass_fwd(K, m, n, h)
    M = sp.coo_matrix((vv, (ii, jj)), shape=(N, N)).tocsr()
    assert abs(M.sum() - K.sum() * h * h / 4) < 1e-13                         # partition of unity
    rhs = oracle.quad_source_fwd(f, m, n, h)
    assert abs(rhs.sum() - f.sum() * h * h / 4) < 1e-13
    assert np.abs(rhs - sp.coo_matrix((oracle.quad_mass_fwd(f, m, n, h)[2], (ii, jj)), shape=(N, N)).tocsr() @ np.ones(N)).max() < 1e-14   # source(c) == mass(c) 1
    # independent numpy assembly: x-linear field u = x has energy sum K h^2/4 (|grad u|^2 = 1)
    x = np.tile(np.arange(m + 1) * h, n + 1)
    assert abs(x @ (L @ x) - K.sum() * h * h / 4) < 1e-12
    # with K = 1 the matrix is the classic Q1 stencil: centre 8/3, every neighbour -1/3
    ii, jj, vv = oracle.quad_laplace_fwd(np.ones(4 * m * n), m, n, h)
    L1 = sp.coo_matrix((vv, (ii, jj)), shape=(N, N)).toarray()
    c = 2 * (m + 1) + 3                                                      # an interior node
    assert abs(L1[c, c] - 8 / 3) < 1e-13
    for dn in (-1, 1, -(m + 1), m + 1, -(m + 2), -m, m, m + 2):
        assert abs(L1[c, c + dn] + 1 / 3) < 1e-13
    # adjoints are transposes
    g = rng.standard_normal(64 * m * n)
    assert abs(oracle.quad_laplace_bwd(g, m, n, h) @ K - g @ oracle.quad_laplace_fwd(K, m, n, h)[2]) < 1e-12
    assert abs(oracle.quad_mass_bwd(g, m, n, h) @ K - g @ oracle.quad_mass_fwd(K, m, n, h)[2]) < 1e-12
    w = rng.standard_normal(N)
    assert abs(oracle.quad_source_bwd(w, m, n, h) @ f - w @ rhs) < 1e-12


def test_plane_stress_matrix_published_value(oracle):
    """docs/src/inverse.md:27 prints the reference H of the poroelasticity example, compute_plane_stress_matrix(E = 1, nu = 0.35), to six decimals:
    [[1.604938, 0.864198, 0], [0.864198, 1.604938, 0], [0, 0, 0.37037]].  Pins mode 1 of the PlaneStrainAndStress op (and which of the two
    reference formulas carries which name)."""
    H = oracle.plane_matrix_fwd([1.0], [0.35], 1)[0]
    ref = np.array([[1.604938, 0.864198, 0.0], [0.864198, 1.604938, 0.0], [0.0, 0.0, 0.37037]])
    assert np.abs(H - ref).max() < 6e-7
    assert np.abs(oracle.plane_matrix_fwd([1.0], [0.35], 0)[0] - ref).max() > 0.1          # mode 0 is the other matrix


def test_dirichlet_bd_known_answer(oracle):
    """DirichletBd on a 1 x 1 cell, two components (8 dofs, 1-based like Julia's find(A)): node 1 on the boundary -> dofs 1 and 5.
    Dense check against the slicing definition: A1 = A with rows / columns {1, 5} replaced by the identity, A2 = A[free, (1, 5)]."""
    m = n = 1
    N = 8
    rng = np.random.default_rng(0)
    D = rng.standard_normal((N, N))
    ii, jj = np.nonzero(np.ones((N, N)))
    vv = D[ii, jj]
    ii1, jj1, vv1, ii2, jj2, vv2 = oracle.dirichlet_bd_fwd(ii + 1, jj + 1, vv, np.array([1]), m, n)
    A1 = np.zeros((N, N)); np.add.at(A1, (ii1 - 1, jj1 - 1), vv1)
    A2 = np.zeros((N, 2)); np.add.at(A2, (ii2 - 1, jj2 - 1), vv2)
    E = D.copy(); E[[0, 4], :] = 0; E[:, [0, 4]] = 0; E[0, 0] = E[4, 4] = 1
    F = D[:, [0, 4]].copy(); F[[0, 4], :] = 0
    assert np.array_equal(A1, E) and np.array_equal(A2, F)
    assert list(ii1[-2:]) == [1, 5] and list(vv1[-2:]) == [1.0, 1.0]           # unit diagonal appended in ascending dof order
    g = oracle.dirichlet_bd_bwd(ii + 1, jj + 1, np.arange(len(vv1), dtype=float), 100 + np.arange(len(vv2), dtype=float), np.array([1]), m, n)
    free = ~np.isin(ii, [0, 4])
    assert np.all(g[~free] == 0) and np.count_nonzero(g[free & np.isin(jj, [0, 4])] >= 100) == 12


def test_boundary_edges_and_nodes_known_answers():
    """`bcedge` / `bcnode` (src/MFEM/MCore.jl:385-450): Mesh(2, 2, h) has 8 boundary edges and every node but the centre on the boundary; on a
    jittered, renumbered triangulation the boundary edges are the (lo, hi) pairs counted once by a dictionary walk over the elements, in
    ascending (lo, hi) order; P2 adds the dofs of those edges."""
    import collections
    import adfem_jl_b200 as A
    from adfem_jl_b200 import meshgen
    m = A.Mesh(2, 2, 0.5, host_only=True)
    assert len(A.bcedge(m)) == 8 and np.array_equal(A.bcnode(m), [0, 1, 2, 3, 5, 6, 7, 8])
    c, e = meshgen.jitter_unstructured(13, 9, 0.1, seed=7)
    for degree in (1, 2):
        m = A.Mesh(c, e, degree=degree, host_only=True)
        seen = collections.Counter()
        for t in m.elems:
            for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
                seen[(min(a, b), max(a, b))] += 1
        want = np.array(sorted(k for k, v in seen.items() if v == 1), dtype=np.int64)
        bd = A.bcedge(m)
        assert bd.dtype == np.int64 and np.array_equal(bd, want) and len(want) == 2 * (13 + 9)
        nodes = np.unique(want.reshape(-1))
        got = A.bcnode(m)
        assert np.array_equal(got[:len(nodes)], nodes)
        if degree == 2:
            edge_of = {(min(a, b), max(a, b)): i for i, (a, b) in enumerate(m.edges)}
            assert np.array_equal(got[len(nodes):], np.sort([edge_of[tuple(k)] for k in want]) + m.nnode)
        else:
            assert len(got) == len(nodes)
