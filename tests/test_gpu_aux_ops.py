"""GPU parity tests of the boundary-row and structured-grid kernels against the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

import adfem_jl_b200 as A
from adfem_jl_b200 import meshgen, ops

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def close(a, b, rel=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    scale = max(np.abs(b).max(), 1e-300) if b.size else 1.0
    err = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), scale * 1e-3)
    assert err.size == 0 or err.max() <= rel, f"max rel err {err.max():.3e}"


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


# ------------------------------------------------------------------------------------------ ImposeDirichlet
@pytest.mark.parametrize("dup", [False, True])
def test_impose_dirichlet_coo(oracle, dup):
    """Op-level parity incl. output ORDER (kept slots in input order, then boundary diagonals ascending, quirk Q11),
    unordered / duplicated `bd` (last value wins) and the three gradients."""
    c, e = meshgen.jitter_unstructured(11, 9, 0.1, seed=2)
    m, o = A.Mesh(c, e, degree=2), oracle.Mesh2D(c, e, degree=2)
    rng = np.random.default_rng(0)
    kappa = rng.random(o.ngauss) + 0.5
    ind, vv = o.laplace_fwd(kappa)
    rhs = rng.standard_normal(o.ndof)
    bd = rng.permutation(A.bcnode(m))                       # bcnode is unordered in the reference (Q12)
    if dup:
        bd = np.concatenate([bd, bd[:7]])
    bdval = rng.standard_normal(len(bd))
    oi, ov, orhs = oracle.impose_dirichlet_fwd(ind, vv, bd, rhs, bdval)
    K = A.compute_fem_laplace_matrix1(dev(kappa), m)         # COO SparseTensor straight from the assembly op
    assert np.array_equal(K.indices.cpu().numpy(), ind)
    v_t, r_t, b_t = dev(vv).requires_grad_(True), dev(rhs).requires_grad_(True), dev(bdval).requires_grad_(True)   # oracle values in: kept slots are exact copies
    B, r2 = A.impose_Dirichlet_boundary_conditions(A.SparseTensor(K.indices, v_t, *K.shape), r_t, bd, b_t)
    assert np.array_equal(B.indices.cpu().numpy(), oi)       # bit-exact indices and order
    assert np.array_equal(B.values.detach().cpu().numpy(), ov)   # values are copies (or exactly 1.0)
    close(r2.detach().cpu().numpy(), orhs)
    w1, w2 = rng.standard_normal(len(ov)), rng.standard_normal(o.ndof)
    gv, gr, gb = torch.autograd.grad([B.values, r2], [v_t, r_t, b_t], [dev(w1), dev(w2)])
    egv, egr, egb = oracle.impose_dirichlet_bwd(w1, w2, ind, vv, bd, bdval, o.ndof)
    close(gv.cpu().numpy(), egv); close(gr.cpu().numpy(), egr); close(gb.cpu().numpy(), egb)


def test_impose_dirichlet_is_bit_reproducible(oracle):
    """rhs correction and grad_bdval are segmented reductions in slot order (no atomics): identical bits run to run, and the rhs is
    bit-identical to the oracle's sequential loop when every product is formed the same way (values chosen exactly representable)."""
    c, e = meshgen.jitter_unstructured(40, 30, 0.1, seed=3)
    m, o = A.Mesh(c, e), oracle.Mesh2D(c, e)
    rng = np.random.default_rng(5)
    ind, _ = o.laplace_fwd(np.ones(o.ngauss))
    vv = rng.integers(-8, 9, len(ind)).astype(np.float64) / 4          # dyadic values: products and partial sums are exact in any association
    rhs = rng.integers(-8, 9, o.ndof).astype(np.float64)
    bd = rng.permutation(A.bcnode(m))
    bdval = rng.integers(-8, 9, len(bd)).astype(np.float64) / 2
    oi, ov, orhs = oracle.impose_dirichlet_fwd(ind, vv, bd, rhs, bdval)
    outs = []
    for _ in range(3):
        v_t, r_t, b_t = dev(vv).requires_grad_(True), dev(rhs).requires_grad_(True), dev(bdval).requires_grad_(True)
        B, r2 = A.impose_Dirichlet_boundary_conditions(A.SparseTensor(dev(ind), v_t, o.ndof, o.ndof), r_t, bd, b_t)
        w = torch.arange(o.ndof, dtype=torch.float64, device="cuda") / 16
        (gb,) = torch.autograd.grad(r2, b_t, w)
        outs.append((r2.detach().cpu().numpy(), gb.cpu().numpy()))
    assert np.array_equal(outs[0][0], orhs)
    for r, g in outs[1:]:
        assert np.array_equal(r, outs[0][0]) and np.array_equal(g, outs[0][1])
    # irrational values: still bit-identical run to run
    vv2 = rng.standard_normal(len(ind))
    res = [A.impose_Dirichlet_boundary_conditions(A.SparseTensor(dev(ind), dev(vv2), o.ndof, o.ndof), dev(rhs), bd, dev(bdval))[1].cpu().numpy() for _ in range(3)]
    assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2])
    close(res[0], oracle.impose_dirichlet_fwd(ind, vv2, bd, rhs, bdval)[2])


@pytest.mark.parametrize("base,dup", [(1, False), (0, False), (1, True)])
def test_dirichlet_bd_op(oracle, base, dup):
    """DirichletBd (deps/DirichletBd/DirichletBd.h:8-60): both outputs bit-exact in order, indices and values (values are copies or 1.0), and
    the gradient; the COO input is the structured two-component stiffness operator (UnivariateFemStiffness slots, duplicates included)."""
    m, n, h = 7, 5, 0.2
    rng = np.random.default_rng(3 + base)
    nn = (m + 1) * (n + 1)
    # a two-component COO operator with duplicates: every cell couples its 4 nodes in both components
    cells = np.arange(m * n)
    ci, cj = cells % m, cells // m
    nodes = np.stack([cj * (m + 1) + ci, cj * (m + 1) + ci + 1, (cj + 1) * (m + 1) + ci, (cj + 1) * (m + 1) + ci + 1], 1)
    dofs = np.concatenate([nodes, nodes + nn], 1)                                   # 8 dofs per cell
    ii = np.repeat(dofs, 8, axis=1).reshape(-1) + base
    jj = np.tile(dofs, (1, 8)).reshape(-1) + base
    vv = rng.standard_normal(len(ii))
    bnode = np.unique(np.concatenate([np.arange(m + 1), np.arange(n + 1) * (m + 1), np.arange(n + 1) * (m + 1) + m]))   # bottom, left, right
    bd = rng.permutation(bnode) + base
    if dup:
        bd = np.concatenate([bd, bd[:4]])
    ref = oracle.dirichlet_bd_fwd(ii, jj, vv, bd, m, n)
    Acoo = A.SparseTensor(dev(np.stack([ii, jj], 1)), dev(vv).requires_grad_(True), 2 * nn + base, 2 * nn + base)
    A1, A2 = A.fem_impose_Dirichlet_boundary_condition_experimental(Acoo, bd, m, n, h)
    got = [A1.indices[:, 0], A1.indices[:, 1], A1.values, A2.indices[:, 0], A2.indices[:, 1], A2.values]
    for g, r in zip(got, ref):
        assert np.array_equal(g.detach().cpu().numpy(), r)
    assert A2.shape[1] == 2 * len(bd) and int(A2.indices[:, 1].max()) <= 2 * len(bd) and int(A2.indices[:, 1].min()) >= 1
    w1, w2 = rng.standard_normal(len(ref[2])), rng.standard_normal(len(ref[5]))
    (g,) = torch.autograd.grad([A1.values, A2.values], Acoo.values, [dev(w1), dev(w2)])
    assert np.array_equal(g.cpu().numpy(), oracle.dirichlet_bd_bwd(ii, jj, w1, w2, bd, m, n))
    with pytest.raises(A.AdfemError):
        A.fem_impose_Dirichlet_boundary_condition_experimental(Acoo, np.array([10 ** 6]), m, n, h)


def test_edge_cases_empty_boundary_and_single_elements(oracle):
    """Degenerate inputs the reference accepts: an empty boundary list (both Dirichlet ops return the matrix unchanged, in order), a
    boundary list covering every dof, and meshes of a single triangle / a single tetrahedron through the CSR and COO paths."""
    rng = np.random.default_rng(8)
    c, e = meshgen.tri_grid(3, 2, 0.5)
    o = oracle.Mesh2D(c, e)
    ind, vv = o.laplace_fwd(np.ones(o.ngauss))
    rhs = rng.standard_normal(o.ndof)
    none = np.zeros(0, dtype=np.int64)
    B, r = A.impose_Dirichlet_boundary_conditions(A.SparseTensor(dev(ind), dev(vv), o.ndof, o.ndof), dev(rhs), none, dev(np.zeros(0)))
    assert np.array_equal(B.indices.cpu().numpy(), ind) and np.array_equal(B.values.cpu().numpy(), vv) and np.array_equal(r.cpu().numpy(), rhs)
    allbd = np.arange(o.ndof)
    bv = rng.standard_normal(o.ndof)
    B, r = A.impose_Dirichlet_boundary_conditions(A.SparseTensor(dev(ind), dev(vv), o.ndof, o.ndof), dev(rhs), allbd, dev(bv))
    oi, ov, orhs = oracle.impose_dirichlet_fwd(ind, vv, allbd, rhs, bv)
    assert np.array_equal(B.indices.cpu().numpy(), oi) and np.array_equal(B.values.cpu().numpy(), ov) and np.array_equal(r.cpu().numpy(), orhs)
    assert len(ov) == o.ndof and np.all(ov == 1.0)
    ii, jj = ind[:, 0] + 1, ind[:, 1] + 1
    Acoo = A.SparseTensor(dev(np.stack([ii, jj], 1)), dev(vv), 2 * 12 + 1, 2 * 12 + 1)
    A1, A2 = A.fem_impose_Dirichlet_boundary_condition_experimental(Acoo, np.zeros(0, dtype=np.int32), 3, 2, 0.5)
    assert np.array_equal(A1.indices.cpu().numpy(), np.stack([ii, jj], 1)) and np.array_equal(A1.values.cpu().numpy(), vv) and A2.values.numel() == 0
    for dim, coords, elems in ((2, np.array([[0.0, 0.0], [1.0, 0.2], [0.3, 0.9]]), np.array([[0, 1, 2]])),
                               (2, np.array([[0.0, 0.0], [1.0, 0.2], [0.3, 0.9]]), np.array([[1, 0, 2]])),          # negatively oriented: the fix swaps v0, v1
                               (3, np.array([[0.0, 0.0, 0.0], [1.0, 0.1, 0.0], [0.2, 0.9, 0.1], [0.1, 0.2, 0.8]]), np.array([[0, 1, 2, 3]]))):
        for degree in (1, 2):
            m = (A.Mesh if dim == 2 else A.Mesh3)(coords, elems, degree=degree)
            om = (oracle.Mesh2D if dim == 2 else oracle.Mesh3D)(coords, elems, degree=degree)
            kap = rng.random(om.ngauss) + 0.5
            ind1, vv1 = om.laplace_fwd(kap)
            rp, ci, ref = oracle.canonical_csr(ind1, vv1, om.ndof)
            kt = dev(kap).requires_grad_(True)
            T = A.compute_fem_laplace_matrix1(kt, m, mode="csr")
            assert np.array_equal(T.rowptr, rp) and np.array_equal(T.colind, ci)
            close(T.values.detach().cpu().numpy(), ref)
            dv = rng.standard_normal(len(ref))
            (g,) = torch.autograd.grad(T.values, kt, dev(dv))
            close(g.cpu().numpy(), om.laplace_bwd(oracle.csr_adjoint_to_slots(rp, ci, dv, ind1, om.ndof)))
            S = A.compute_fem_laplace_matrix1(dev(kap), m, mode="coo")
            assert np.array_equal(S.indices.cpu().numpy(), ind1)
            close(S.values.cpu().numpy(), vv1)


def test_impose_dirichlet_eager_matches_dense_julia_version():
    """test/mfem.jl:78-88: the op equals the dense slicing implementation (src/MFEM/MUtils.jl:184-199)."""
    rng = np.random.default_rng(1)
    N = 9
    Am = rng.random((N, N))
    bd, bdval, rhs = np.array([4, 1, 2]), np.array([1.0, 2.0, 3.0]), rng.random(N)
    B, r = A.impose_Dirichlet_boundary_conditions(sp.csr_matrix(Am), rhs, bd, bdval)
    idx = np.ones(N, bool); idx[bd] = False
    r0 = rhs.copy(); r0[idx] = rhs[idx] - Am[np.ix_(idx, bd)] @ bdval; r0[bd] = bdval
    B0 = np.zeros((N, N)); B0[np.ix_(idx, idx)] = Am[np.ix_(idx, idx)]; B0[bd, bd] = 1.0
    assert np.allclose(B.toarray(), B0, atol=1e-15) and np.allclose(r, r0, atol=1e-14)
    Bh = A.impose_Dirichlet_boundary_conditions(sp.csr_matrix(Am), bd)      # homogeneous helper (MUtils.jl:220-225)
    assert np.allclose(Bh.toarray(), B0, atol=1e-15)


def test_impose_dirichlet_edge_cases(oracle):
    ind = np.array([[0, 0], [0, 1], [1, 0], [1, 1], [2, 2], [0, 0]])
    vv = np.arange(1.0, 7.0)
    rhs = np.array([1.0, 2.0, 3.0])
    for bd in (np.zeros(0, dtype=np.int64), np.array([0, 1, 2]), np.array([1])):      # no boundary, everything boundary, one dof
        bdval = np.arange(len(bd), dtype=float) + 0.5
        oi, ov, orhs = oracle.impose_dirichlet_fwd(ind, vv, bd, rhs, bdval)
        B, r = A.impose_Dirichlet_boundary_conditions(A.SparseTensor(dev(ind), dev(vv), 3, 3), dev(rhs), bd, dev(bdval))
        assert np.array_equal(B.indices.cpu().numpy().reshape(-1, 2), oi) and np.array_equal(B.values.cpu().numpy(), ov)
        close(r.cpu().numpy(), orhs)
    with pytest.raises(A.AdfemError):
        A.impose_Dirichlet_boundary_conditions(A.SparseTensor(dev(ind), dev(vv), 3, 3), dev(rhs), np.array([7]), dev(np.ones(1)))


# ------------------------------------------------------------------------------------------ structured quads
@pytest.mark.parametrize("m,n", [(5, 3), (1, 1), (16, 37)])
def test_univariate_stiffness(oracle, m, n):
    h = 0.3
    rng = np.random.default_rng(2)
    for hm in (rng.random((4 * m * n, 2, 2)), rng.random((2, 2))):
        ii, jj, vv = oracle.univariate_stiffness_fwd(hm, m, n, h)
        t = dev(hm).requires_grad_(True)
        S = A.compute_fem_stiffness_matrix1(t, m, n, h)
        assert np.array_equal(S.indices.cpu().numpy(), np.stack([ii - 1, jj - 1], 1))
        close(S.values.detach().cpu().numpy(), vv)
        gv = rng.standard_normal(len(vv))
        (g,) = torch.autograd.grad(S.values, t, dev(gv))
        close(g.cpu().numpy().reshape(-1), oracle.univariate_stiffness_bwd(gv, m, n, h, hm.ndim == 3))
        Se = A.compute_fem_stiffness_matrix1(hm, m, n, h)                              # eager numpy path
        ref = sp.coo_matrix((vv, (ii - 1, jj - 1)), shape=Se.shape).tocsr()
        assert abs(Se - ref).max() < 1e-12 * abs(vv).max()


@pytest.mark.parametrize("m,n", [(4, 6), (1, 2), (23, 9)])
def test_grid_elasticity_and_svt(oracle, m, n):
    h = 0.25
    rng = np.random.default_rng(3)
    for hm, fwd, bwd in ((rng.random((3, 3)), oracle.fem_stiffness_fwd, oracle.fem_stiffness_bwd),
                         (rng.random((4 * m * n, 3, 3)), oracle.spatial_stiffness_fwd, oracle.spatial_stiffness_bwd)):
        ii, jj, vv = fwd(hm, m, n, h)
        t = dev(hm).requires_grad_(True)
        S = A.compute_fem_stiffness_matrix(t, m, n, h)
        assert S.shape == (2 * (m + 1) * (n + 1),) * 2
        assert np.array_equal(S.indices.cpu().numpy(), np.stack([ii - 1, jj - 1], 1))
        close(S.values.detach().cpu().numpy(), vv)
        gv = rng.standard_normal(len(vv))
        (g,) = torch.autograd.grad(S.values, t, dev(gv))
        close(g.cpu().numpy().reshape(-1), bwd(gv, m, n, h))
        if hm.ndim == 2:      # constant-H adjoint is a fixed-order reduction: bit-reproducible
            (g2,) = torch.autograd.grad(A.compute_fem_stiffness_matrix(t, m, n, h).values, t, dev(gv))
            assert torch.equal(g, g2)
    for type_ in (1, 2, 3):
        mu = rng.random(4 * m * n * type_)
        t = dev(mu).requires_grad_(True)
        H = A.compute_space_varying_tangent_elasticity_matrix(t, m, n, h, type_)
        assert tuple(H.shape) == (4 * m * n, 2, 2)
        assert np.array_equal(H.detach().cpu().numpy().reshape(-1), oracle.svt_fwd(mu, m, n, type_))
        gh = rng.standard_normal(16 * m * n)
        (g,) = torch.autograd.grad(H, t, dev(gh).view(-1, 2, 2))
        assert np.array_equal(g.cpu().numpy(), oracle.svt_bwd(gh, m, n, type_))
    # config 3's structured variant: SVT type 1 feeding compute_fem_stiffness_matrix1, gradient through both ops
    mu = dev(rng.random(4 * m * n) + 0.5).requires_grad_(True)
    S = A.compute_fem_stiffness_matrix1(A.compute_space_varying_tangent_elasticity_matrix(mu, m, n, h, 1), m, n, h)
    (g,) = torch.autograd.grad((S.values ** 2).sum(), mu)
    hm = oracle.svt_fwd(mu.detach().cpu().numpy(), m, n, 1).reshape(-1, 2, 2)
    _, _, vv = oracle.univariate_stiffness_fwd(hm, m, n, h)
    expect = oracle.svt_bwd(oracle.univariate_stiffness_bwd(2 * vv, m, n, h, True), m, n, 1)
    close(g.cpu().numpy(), expect)


# ------------------------------------------------------------------------------------------ PCL Jacobians (src/pcl.jl, test/pcl.jl)
@pytest.mark.parametrize("degree", [1, 2])
def test_pcl_laplace_jacobian(oracle, degree):
    """pcl_compute_fem_laplace_matrix1 == the oracle's restatement of pcl_FemLaplaceScalar_Jacobian, == autograd of the COO op (the check
    test/pcl.jl:34-50 makes with finite differences), through the device API and through the legacy host symbol."""
    import ctypes as C
    c, e = meshgen.jitter_unstructured(4, 3, 0.25, seed=5)
    m, o = A.Mesh(c, e, degree=degree), oracle.Mesh2D(c, e, degree=degree)
    J = ops.pcl_compute_fem_laplace_matrix1(m)
    ref = o.laplace_jacobian()
    assert tuple(J.shape) == ref.shape == (o.ngauss, o.ngauss * o.elem_ndof ** 2)
    close(J.cpu().numpy(), ref)
    k = torch.rand(o.ngauss, dtype=torch.float64, device="cuda").requires_grad_(True)
    vv = ops.compute_fem_laplace_matrix1(k, m, mode="coo").values
    Ja = torch.autograd.functional.jacobian(lambda kk: ops.compute_fem_laplace_matrix1(kk, m, mode="coo").values, k)     # [N, G]
    close(J.cpu().numpy(), Ja.t().cpu().numpy())
    assert vv.numel() == J.shape[1]
    # legacy symbol on the global 2-D mesh: host pointer, caller-zeroed, column-major
    L = A._lib.lib()
    v3 = np.zeros((c.shape[0], 3)); v3[:, :2] = c
    ne = C.c_longlong()
    L.init_nnfem_mesh(np.ascontiguousarray(v3).ctypes.data_as(A._lib.c_dp), c.shape[0], np.ascontiguousarray(e, dtype=np.int32).ctypes.data_as(A._lib.c_ip),
                      e.shape[0], 2 if degree == 1 else 4, 6, degree, C.byref(ne))
    H = np.zeros(ref.size)
    L.pcl_FemLaplaceScalar_Jacobian(H.ctypes.data_as(A._lib.c_dp))
    close(H.reshape(ref.shape[1], ref.shape[0]).T, ref)


def test_pcl_impose_dirichlet(oracle):
    """pcl_impose_Dirichlet_boundary_conditions: J[i, j] = d (v_B)_j / d (v_A)_i is the 0/1 selection of the kept slots, in order —
    checked against autograd of the ImposeDirichlet op and through the legacy host symbol (1-based, column-major indices)."""
    import ctypes as C
    c, e = meshgen.tri_grid(4, 3, 0.25)
    m = A.Mesh(c, e)
    k = torch.rand(m.ngauss, dtype=torch.float64, device="cuda")
    K = ops.compute_fem_laplace_matrix1(k, m, mode="coo")
    bd = A.bcnode(m)
    B = ops.impose_Dirichlet_boundary_conditions(K, bd)
    ind = K.indices.cpu().numpy()
    J = ops.pcl_impose_Dirichlet_boundary_conditions(ind, bd, B.values.numel())
    Ja = torch.autograd.functional.jacobian(lambda v: ops.impose_Dirichlet_boundary_conditions(ops.SparseTensor(K.indices, v, m.ndof, m.ndof), bd).values,
                                            K.values.detach().clone())                                                  # [outdof, sN]
    assert np.array_equal(J.cpu().numpy(), Ja.t().cpu().numpy())
    isbd = np.zeros(m.ndof, dtype=bool); isbd[np.asarray(bd)] = True
    kept = np.flatnonzero(~isbd[ind[:, 0]] & ~isbd[ind[:, 1]])
    Jr = np.zeros(J.shape); Jr[kept, np.arange(len(kept))] = 1.0
    assert np.array_equal(J.cpu().numpy(), Jr)
    L = A._lib.lib()
    sN = ind.shape[0]
    ind1 = np.ascontiguousarray((ind + 1).T.reshape(-1), dtype=np.int64)          # column-major sN x 2, 1-based
    bd1 = np.ascontiguousarray(np.asarray(bd) + 1, dtype=np.int64)
    Jh = np.zeros(sN * J.shape[1])
    L.pcl_ImposeDirichlet(Jh.ctypes.data_as(A._lib.c_dp), ind1.ctypes.data_as(A._lib.c_lp), bd1.ctypes.data_as(A._lib.c_lp), C.c_int(len(bd1)), C.c_int(sN))
    assert np.array_equal(Jh.reshape(J.shape[1], sN).T, Jr)
