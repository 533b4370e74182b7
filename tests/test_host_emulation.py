"""CPU tests of the Gauss-point operator KERNEL BODIES (adfem.jl_b200/csrc/gauss_ops.cuh) against the oracle.

The bodies are `__host__ __device__`; tests/host_emul/emul.cu compiles them for the host (nvcc, no GPU needed) and runs them in loops
that mirror the kernels of gauss_ops.cu.  This checks the arithmetic the GPU executes where no GPU exists; the GPU parity tests proper
(through the C ABI) are in tests/test_widen_gauss_ops.py.  The harness is test infrastructure: nothing in the package can reach it.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import adfem_jl_b200 as A
from adfem_jl_b200 import meshgen

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host_emul", "emul.cu")
SO = os.path.join(HERE, "host_emul", "_build", "libadfem_emul.so")
CSRC = os.path.join(ROOT, "adfem.jl_b200", "csrc")

FEM_TO_GAUSS, DOF_TO_GAUSS, GRAD, STRAIN, STRAIN_ENERGY = range(5)


@pytest.fixture(scope="module")
def emul():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("gauss_ops.cuh", "quad_ops.cuh", "grid_elast.cuh", "grid_gauss.cuh", "grid_index.cuh", "tet_grid.cuh", "tet_grid_tables.h", "row_gather.cuh", "tet_gauss.cuh", "tet_scalar.cuh", "device_fem.cuh", "quadrature.h")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call([nvcc, "-x", "cu", "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
                               "-shared", "-Xcompiler", "-fPIC", "-I", CSRC, SRC, "-o", SO])
    return C.CDLL(SO)


def close(a, b, rel=1e-12):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    scale = max(np.abs(b).max(), 1e-300) if b.size else 1.0
    err = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), scale * 1e-3)
    assert err.size == 0 or err.max() <= rel, f"max rel err {err.max():.3e}"


class HostTables:
    """What the device sees of a mesh (struct-of-arrays connectivity, packed coordinates, dof -> (element, local) adjacency in ascending
    element order), built here from the ORACLE's tables — the harness shares no mesh code with the product."""

    def __init__(self, o):
        self.o = o
        self.dim = o.dim
        self.coords = np.ascontiguousarray(o.coords[:, :o.dim], dtype=np.float64)
        self.verts = np.ascontiguousarray(o.elems.T, dtype=np.int32)         # [dim+1][ne], post orientation fix
        self.conn = np.ascontiguousarray(o.conn.T, dtype=np.int32)           # [d][ne]
        ne, d = o.conn.shape
        dof = o.conn.reshape(-1)
        order = np.argsort(dof, kind="stable")                                 # stable: ascending element, then local index
        self.adj_elem = np.ascontiguousarray(order // d, dtype=np.int32)
        self.adj_loc = np.ascontiguousarray(order % d, dtype=np.uint8)
        self.adj_ptr = np.zeros(o.ndof + 1, dtype=np.int64)
        np.cumsum(np.bincount(dof, minlength=o.ndof), out=self.adj_ptr[1:])

    def _mesh_args(self):
        o = self.o
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        return (C.c_int(o.dim), C.c_int(o.degree), C.c_int(o.order), C.c_int(o.nnode), C.c_int(o.nelem), C.c_int(o.ndof),
                p(self.coords, C.c_double), p(self.verts, C.c_int), p(self.conn, C.c_int))

    def _adj_args(self):
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        return (p(self.adj_ptr, C.c_longlong), p(self.adj_elem, C.c_int), p(self.adj_loc, C.c_ubyte))

    def gauss_op(self, L, kind, adjoint, x, nout):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.full(nout, np.nan)                 # the kernels overwrite: no zero fill
        rc = L.emul_gauss_op(*self._mesh_args(), *self._adj_args(), C.c_int(kind), C.c_int(int(adjoint)),
                             x.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        return out

    def laplace_term(self, L, nu, u):
        nu, u = np.ascontiguousarray(nu, dtype=np.float64), np.ascontiguousarray(u, dtype=np.float64)
        out = np.full(self.o.ndof, np.nan)
        d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        assert L.emul_laplace_term(*self._mesh_args(), *self._adj_args(), d(nu), d(u), d(out)) == 0
        return out

    def laplace_term_grad_nu(self, L, u, go):
        u, go = np.ascontiguousarray(u, dtype=np.float64), np.ascontiguousarray(go, dtype=np.float64)
        out = np.full(self.o.ngauss, np.nan)
        d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        assert L.emul_laplace_term_grad_nu(*self._mesh_args(), d(u), d(go), d(out)) == 0
        return out


def mesh2(oracle, degree):
    c, e = meshgen.jitter_unstructured(9, 7, 0.1, seed=3)
    return oracle.Mesh2D(c, e, degree=degree)


@pytest.mark.parametrize("degree", [1, 2])
def test_gauss_ops_2d_match_oracle(emul, oracle, degree):
    o = mesh2(oracle, degree)
    T = HostTables(o)
    rng = np.random.default_rng(degree)
    G, n, nv = o.ngauss, o.ndof, o.nnode
    u, u2 = rng.standard_normal(n), rng.standard_normal(2 * n)
    cases = [   # kind, forward input, forward output length, oracle forward, oracle adjoint, adjoint output length
        (FEM_TO_GAUSS, u[:nv], G, o.fem_to_gauss_fwd, o.fem_to_gauss_bwd, nv),
        (DOF_TO_GAUSS, u, G, o.dof_to_gauss_fwd, o.dof_to_gauss_bwd, n),
        (GRAD, u, 2 * G, o.grad_fwd, o.grad_bwd, n),
        (STRAIN, u2, 3 * G, o.strain_fwd, o.strain_bwd, 2 * n),
        (STRAIN_ENERGY, rng.standard_normal(3 * G), 2 * n, o.strain_energy_fwd, o.strain_energy_bwd, 3 * G),
    ]
    for kind, x, nout, fwd, bwd, nin in cases:
        x_full = np.concatenate([x, np.zeros(n - nv)]) if kind == FEM_TO_GAUSS else x     # the oracle indexes u by node: any length >= nnode
        close(T.gauss_op(emul, kind, False, x, nout), fwd(x_full))
        w = rng.standard_normal(nout)
        close(T.gauss_op(emul, kind, True, w, nin), bwd(w))


@pytest.mark.parametrize("degree", [1, 2])
def test_laplace_term_2d_matches_oracle(emul, oracle, degree):
    o = mesh2(oracle, degree)
    T = HostTables(o)
    rng = np.random.default_rng(10 + degree)
    nu, u, go = rng.random(o.ngauss) + 0.5, rng.standard_normal(o.ndof), rng.standard_normal(o.ndof)
    close(T.laplace_term(emul, nu, u), o.laplace_term_fwd(nu, u))
    gnu, gu = o.laplace_term_bwd(go, nu, u)
    close(T.laplace_term_grad_nu(emul, u, go), gnu)
    close(T.laplace_term(emul, nu, go), gu)         # adfem_laplace_term_adjoint: grad_u = term(nu, grad_out)
    # second opinion: the term is the assembled Laplace matrix applied to u
    import scipy.sparse as sp
    ind, vv = o.laplace_fwd(nu)
    K = sp.coo_matrix((vv, (ind[:, 0], ind[:, 1])), shape=(o.ndof, o.ndof)).tocsr()
    close(T.laplace_term(emul, nu, u), K @ u, rel=1e-10)


@pytest.mark.parametrize("degree", [1, 2])
def test_ops_3d(emul, oracle, degree):
    c, e = meshgen.tet_grid(3, 3, 2, 0.25)
    rng = np.random.default_rng(5)
    c = c + rng.uniform(-0.03, 0.03, c.shape)
    o = oracle.Mesh3D(c, e, degree=degree)
    T = HostTables(o)
    nu, u, go = rng.random(o.ngauss) + 0.5, rng.standard_normal(o.ndof), rng.standard_normal(o.ndof)
    close(T.laplace_term(emul, nu, u), o.laplace_term_fwd(nu, u))
    gnu, gu = o.laplace_term_bwd(go, nu, u)
    close(T.laplace_term_grad_nu(emul, u, go), gnu)
    close(T.laplace_term(emul, nu, go), gu)
    # 3-D transfers (no reference twin): against the oracle's own shape tables
    h, hx, hy, hz = o.shape_tables()                       # [ne, d, g]
    ul = u[o.conn]                                         # [ne, d]
    G = o.ngauss
    close(T.gauss_op(emul, DOF_TO_GAUSS, False, u, G), np.einsum("ed,edk->ek", ul, h).reshape(-1))
    grad = np.stack([np.einsum("ed,edk->ek", ul, t) for t in (hx, hy, hz)], axis=2).reshape(-1)
    close(T.gauss_op(emul, GRAD, False, u, 3 * G), grad)
    # adjoints: <A x, w> == <x, A^T w>
    for kind, nin, nout in ((FEM_TO_GAUSS, o.nnode, G), (DOF_TO_GAUSS, o.ndof, G), (GRAD, o.ndof, 3 * G), (STRAIN, 3 * o.ndof, 6 * G),
                            (STRAIN_ENERGY, 6 * G, 3 * o.ndof)):
        x, w = rng.standard_normal(nin), rng.standard_normal(nout)
        lhs = T.gauss_op(emul, kind, False, x, nout) @ w
        rhs = x @ T.gauss_op(emul, kind, True, w, nin)
        assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), abs(rhs), 1.0)
    # 3-D strain (Voigt xx, yy, zz, yz, xz, xy; engineering shears) of a linear displacement field is exact
    Agrad = rng.standard_normal((3, 3))
    pos = np.zeros((o.ndof, 3)); pos[:o.nnode] = c
    if degree == 2:
        pos[o.nnode:] = 0.5 * (c[o.edges[:, 0]] + c[o.edges[:, 1]])
    disp = pos @ Agrad.T                                   # u_i = A_ij x_j
    eps = T.gauss_op(emul, STRAIN, False, disp.T.reshape(-1), 6 * G).reshape(G, 6)
    ref = np.array([Agrad[0, 0], Agrad[1, 1], Agrad[2, 2], Agrad[1, 2] + Agrad[2, 1], Agrad[0, 2] + Agrad[2, 0], Agrad[0, 1] + Agrad[1, 0]])
    assert np.abs(eps - ref).max() < 1e-11


def test_plane_matrix(emul, oracle):
    rng = np.random.default_rng(7)
    N = 257
    E, nu = rng.random(N) + 0.5, rng.random(N) * 0.45
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for mode in (0, 1):
        H = np.full(9 * N, np.nan)
        emul.emul_plane_matrix(C.c_int(mode), C.c_longlong(N), d(E), d(nu), d(H))
        ref = oracle.plane_matrix_fwd(E, nu, mode)
        assert np.array_equal(H.reshape(N, 3, 3), ref)                    # same expressions: bit-exact
        g = rng.standard_normal(9 * N)
        gE, gnu = np.full(N, np.nan), np.full(N, np.nan)
        emul.emul_plane_matrix_grad(C.c_int(mode), C.c_longlong(N), d(E), d(nu), d(g), d(gE), d(gnu))
        rE, rnu = oracle.plane_matrix_bwd(g, E, nu, mode)
        close(gE, rE, rel=1e-11); close(gnu, rnu, rel=1e-11)
    # the documented closed forms (src/Core.jl:742-768)
    E0, n0 = 2.0, 0.3
    ref0 = E0 * (1 - n0) / (1 + n0) / (1 - 2 * n0) * np.array([[1, n0 / (1 - n0), n0 / (1 - n0)], [n0 / (1 - n0), 1, n0 / (1 - n0)], [n0 / (1 - n0), n0 / (1 - n0), 1]])
    ref1 = E0 / (1 + n0) / (1 - 2 * n0) * np.array([[1 - n0, n0, 0], [n0, 1 - n0, 0], [0, 0, (1 - 2 * n0) / 2]])
    assert np.allclose(oracle.plane_matrix_fwd([E0], [n0], 0)[0], ref0, rtol=1e-15)
    assert np.allclose(oracle.plane_matrix_fwd([E0], [n0], 1)[0], ref1, rtol=1e-15)


@pytest.mark.parametrize("dim", [2, 3])
def test_coef_presum_bodies(emul, oracle, dim):
    """Option "coef_presum": for P1 elements the stiffness blocks of (sum_k w_k H_k at ONE point of weight 1) equal those of the g-point
    rule, and the per-Gauss-point gradient is w_k times the one-point gradient — checked with the oracle's own stiffness ops."""
    rng = np.random.default_rng(dim)
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    if dim == 2:
        c, e = meshgen.jitter_unstructured(6, 5, 0.1, seed=1)
        o, o1 = oracle.Mesh2D(c, e), oracle.Mesh2D(c, e, order=1)
        ns2, wsum = 9, 0.5
    else:
        c, e = meshgen.tet_grid(2, 2, 2, 0.5)
        o, o1 = oracle.Mesh3D(c, e), oracle.Mesh3D(c, e, order=1)
        ns2, wsum = 36, 1.0 / 6.0
    assert o1.g == 1 and o.g > 1
    H = rng.random(o.ngauss * ns2)
    hbar = np.full(o.nelem * ns2, np.nan)
    assert emul.emul_presum_coef(C.c_int(dim), C.c_int(o.order), C.c_longlong(o.nelem), C.c_int(ns2), d(H), d(hbar)) == 0
    # the tile kernel applies weight 1 to hbar; the oracle's one-point rule has weight wsum, so feed it hbar / wsum
    _, vv = o.stiffness_fwd(H)
    _, vv1 = o1.stiffness_fwd(hbar / wsum)
    D2 = (dim * o.elem_ndof) ** 2
    close(vv.reshape(o.nelem, o.g, D2).sum(1), vv1.reshape(o.nelem, D2), rel=1e-11)
    # adjoint: per-element gradient (one-point rule, weight 1 => oracle gradient / wsum) expanded with w_k
    gvv_e = rng.standard_normal((o.nelem, D2))
    gbar = o1.stiffness_bwd(gvv_e.reshape(-1)) / wsum
    grad = np.full(o.ngauss * ns2, np.nan)
    assert emul.emul_expand_grad(C.c_int(dim), C.c_int(o.order), C.c_longlong(o.nelem), C.c_int(ns2), d(np.ascontiguousarray(gbar)), d(grad)) == 0
    ref = o.stiffness_bwd(np.repeat(gvv_e[:, None, :], o.g, axis=1).reshape(-1))       # every Gauss-point block of an element gets the same upstream
    close(grad, ref, rel=1e-11)


def test_quad_scalar_siblings(emul, oracle):
    """FemLaplace / FemMass / FemSource bodies (csrc/quad_ops.cuh) against the oracle: indices bit-exact (0-based), values within 1e-12."""
    rng = np.random.default_rng(11)
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    l = lambda a: a.ctypes.data_as(C.POINTER(C.c_longlong))
    for m, n, h in ((5, 3, 0.1), (1, 1, 2.0), (2, 7, 0.37)):
        coef = rng.random(4 * m * n) + 0.5
        N = 64 * m * n
        for op, fwd, bwd in ((0, oracle.quad_laplace_fwd, oracle.quad_laplace_bwd), (1, oracle.quad_mass_fwd, oracle.quad_mass_bwd)):
            ii, jj, vv = np.full(N, -1, dtype=np.int64), np.full(N, -1, dtype=np.int64), np.full(N, np.nan)
            emul.emul_quad_scalar(C.c_int(op), d(coef), C.c_int(m), C.c_int(n), C.c_double(h), l(ii), l(jj), d(vv))
            ri, rj, rv = fwd(coef, m, n, h)
            assert np.array_equal(ii, ri) and np.array_equal(jj, rj)
            close(vv, rv)
            g = rng.standard_normal(N)
            out = np.full(4 * m * n, np.nan)
            emul.emul_quad_scalar_grad(C.c_int(op), d(g), C.c_int(m), C.c_int(n), C.c_double(h), d(out))
            close(out, bwd(g, m, n, h))
        rhs = np.full((m + 1) * (n + 1), np.nan)
        emul.emul_quad_source(d(coef), C.c_int(m), C.c_int(n), C.c_double(h), d(rhs))
        close(rhs, oracle.quad_source_fwd(coef, m, n, h))
        w = rng.standard_normal((m + 1) * (n + 1))
        gf = np.full(4 * m * n, np.nan)
        emul.emul_quad_source_grad(d(w), C.c_int(m), C.c_int(n), C.c_double(h), d(gf))
        close(gf, oracle.quad_source_bwd(w, m, n, h))


@pytest.mark.parametrize("mode", [0, 1])
def test_fused_plane_presum_bodies(emul, oracle, mode):
    """adfem_assemble_csr_plane: pre-sum straight from (E, nu) == pre-sum of the materialised plane matrices; the gradient expansion equals
    the oracle's plane-matrix adjoint applied to the expanded per-Gauss-point gradient."""
    rng = np.random.default_rng(40 + mode)
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ne, order, g = 37, 2, 3
    E, nu = rng.random(ne * g) + 0.5, rng.random(ne * g) * 0.45
    H = oracle.plane_matrix_fwd(E, nu, mode).reshape(-1)
    ref, got = np.full(ne * 9, np.nan), np.full(ne * 9, np.nan)
    assert emul.emul_presum_coef(C.c_int(2), C.c_int(order), C.c_longlong(ne), C.c_int(9), d(H), d(ref)) == 0
    assert emul.emul_presum_plane(C.c_int(order), C.c_longlong(ne), C.c_int(mode), d(E), d(nu), d(got)) == 0
    assert np.array_equal(got, ref)
    gbar = rng.standard_normal(ne * 9)
    gH = np.full(ne * g * 9, np.nan)
    assert emul.emul_expand_grad(C.c_int(2), C.c_int(order), C.c_longlong(ne), C.c_int(9), d(gbar), d(gH)) == 0
    rE, rnu = oracle.plane_matrix_bwd(gH, E, nu, mode)
    gE, gnu = np.full(ne * g, np.nan), np.full(ne * g, np.nan)
    assert emul.emul_expand_plane_grad(C.c_int(order), C.c_longlong(ne), C.c_int(mode), d(E), d(nu), d(gbar), d(gE), d(gnu)) == 0
    close(gE, rE, rel=1e-11); close(gnu, rnu, rel=1e-11)


@pytest.mark.parametrize("m,n,rpw,heron", [(5, 4, 8, 0), (1, 1, 3, 1), (33, 3, 2, 0), (70, 9, 4, 1), (32, 2, 100, 0), (31, 6, 1, 0)])
def test_structured_elasticity_warp_phases(emul, oracle, m, n, rpw, heron):
    """grid_elast.cuh: every warp of the structured P1-elasticity kernels run as loops over its lanes (shared memory poisoned first), on
    rectilinear non-uniform grids with one or several strips / chunks, against the canonical CSR of the oracle's stiffness op."""
    rng = np.random.default_rng(m * 100 + n)
    xs = np.concatenate([[0.0], np.cumsum(rng.random(m) * 0.1 + 0.05)])
    ys = np.concatenate([[0.3], 0.3 + np.cumsum(rng.random(n) * 0.1 + 0.05)])
    c, e = meshgen.tri_grid(m, n, 1.0)
    c = np.stack([np.tile(xs, n + 1), np.repeat(ys, m + 1)], 1)
    o = oracle.Mesh2D(c, e)
    N2 = 2 * o.ndof
    H = rng.random(o.ngauss * 9) + 0.1
    ind, vv = o.stiffness_fwd(H)
    rp, ci, ref = oracle.canonical_csr(ind, vv, N2)
    nnz_s = len(ref) // 4
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    vals = np.full(len(ref), np.nan)
    assert emul.emul_grid_elast_fwd(C.c_int(m), C.c_int(n), d(xs), d(ys), C.c_int(2), C.c_int(heron), C.c_int(rpw), C.c_longlong(nnz_s), d(H), d(vals),
                                    C.c_int(-1), None) == 0
    close(vals, ref, rel=1e-12)
    dv = rng.standard_normal(len(ref))
    expect = o.stiffness_bwd(oracle.csr_adjoint_to_slots(rp, ci, dv, ind, N2))
    grad = np.full(o.ngauss * 9, np.nan)
    assert emul.emul_grid_elast_adj(C.c_int(m), C.c_int(n), d(xs), d(ys), C.c_int(2), C.c_int(heron), C.c_int(rpw), C.c_longlong(nnz_s), d(dv), d(grad),
                                    C.c_int(-1), None, None, None) == 0
    close(grad, expect, rel=1e-12)
    # fused constitutive step: tangents straight from the moduli, gradients with respect to the moduli
    for mode in (0, 1):
        E, nu = rng.random(o.ngauss) + 0.5, rng.random(o.ngauss) * 0.4
        Hm = oracle.plane_matrix_fwd(E, nu, mode).reshape(-1)
        _, vvm = o.stiffness_fwd(Hm)
        _, _, refm = oracle.canonical_csr(ind, vvm, N2)
        vals = np.full(len(ref), np.nan)
        assert emul.emul_grid_elast_fwd(C.c_int(m), C.c_int(n), d(xs), d(ys), C.c_int(2), C.c_int(heron), C.c_int(rpw), C.c_longlong(nnz_s), d(E), d(vals),
                                        C.c_int(mode), d(nu)) == 0
        close(vals, refm, rel=1e-12)
        rE, rnu = oracle.plane_matrix_bwd(expect, E, nu, mode)
        gE, gnu = np.full(o.ngauss, np.nan), np.full(o.ngauss, np.nan)
        assert emul.emul_grid_elast_adj(C.c_int(m), C.c_int(n), d(xs), d(ys), C.c_int(2), C.c_int(heron), C.c_int(rpw), C.c_longlong(nnz_s), d(dv), d(gE),
                                        C.c_int(mode), d(E), d(nu), d(gnu)) == 0
        close(gE, rE, rel=1e-11); close(gnu, rnu, rel=1e-11)


@pytest.mark.parametrize("m,n,rpw,heron", [(5, 4, 3, 0), (1, 1, 8, 1), (33, 3, 2, 0), (70, 9, 4, 1)])
def test_structured_elasticity_warp_phases_mapped_grid(emul, oracle, m, n, rpw, heron):
    """The MAPPED instantiation of the same phases: structured connectivity on smoothly mapped + jittered node positions, the seven stencil
    positions read from the coordinate array; against the canonical CSR of the oracle's stiffness op and its adjoint."""
    rng = np.random.default_rng(m * 100 + n + 7)
    c, e = meshgen.tri_grid(m, n, 1.0)
    x, y = c[:, 0], c[:, 1]
    c = np.stack([x + 0.08 * np.sin(0.7 * y), y + 0.06 * np.cos(0.9 * x)], 1) + rng.uniform(-0.1, 0.1, c.shape)
    c = np.ascontiguousarray(c)
    o = oracle.Mesh2D(c, e)
    assert np.array_equal(o.elems, oracle.Mesh2D(*meshgen.tri_grid(m, n, 1.0)).elems)       # no triangle flipped: same connectivity after the orientation fix
    N2 = 2 * o.ndof
    H = rng.random(o.ngauss * 9) + 0.1
    ind, vv = o.stiffness_fwd(H)
    rp, ci, ref = oracle.canonical_csr(ind, vv, N2)
    nnz_s = len(ref) // 4
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    vals = np.full(len(ref), np.nan)
    assert emul.emul_grid_elast_fwd_mapped(C.c_int(m), C.c_int(n), d(c), C.c_int(2), C.c_int(heron), C.c_int(rpw), C.c_longlong(nnz_s), d(H), d(vals)) == 0
    close(vals, ref, rel=1e-12)
    dv = rng.standard_normal(len(ref))
    expect = o.stiffness_bwd(oracle.csr_adjoint_to_slots(rp, ci, dv, ind, N2))
    grad = np.full(o.ngauss * 9, np.nan)
    assert emul.emul_grid_elast_adj_mapped(C.c_int(m), C.c_int(n), d(c), C.c_int(2), C.c_int(heron), C.c_int(rpw), C.c_longlong(nnz_s), d(dv), d(grad)) == 0
    close(grad, expect, rel=1e-12)


@pytest.mark.parametrize("m,n", [(5, 4), (1, 1), (9, 2)])
def test_structured_scatter_operators(emul, oracle, m, n):
    """grid_gauss.cuh: the scatter-type operators on Mesh(m, n, h) as one thread per node with index arithmetic — against the oracle, and
    bit-identical to the general adjacency-walking bodies (same summation order, same geometry)."""
    rng = np.random.default_rng(m * 10 + n)
    xs = np.concatenate([[0.0], np.cumsum(rng.random(m) * 0.1 + 0.05)])
    ys = np.concatenate([[0.3], 0.3 + np.cumsum(rng.random(n) * 0.1 + 0.05)])
    _, e = meshgen.tri_grid(m, n, 1.0)
    c = np.stack([np.tile(xs, n + 1), np.repeat(ys, m + 1)], 1)
    o = oracle.Mesh2D(c, e)
    T = HostTables(o)
    G, nd = o.ngauss, o.ndof
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    cases = [   # basis, weighted, input length, output length, oracle, general-path kind / adjoint flag
        (0, 0, G, nd, o.fem_to_gauss_bwd, FEM_TO_GAUSS, True), (1, 0, G, nd, o.dof_to_gauss_bwd, DOF_TO_GAUSS, True),
        (2, 0, 2 * G, nd, o.grad_bwd, GRAD, True), (3, 0, 3 * G, 2 * nd, o.strain_bwd, STRAIN, True),
        (3, 1, 3 * G, 2 * nd, o.strain_energy_fwd, STRAIN_ENERGY, False)]
    for basis, weighted, nin, nout, ref, kind, adjoint in cases:
        x = rng.standard_normal(nin)
        out = np.full(nout, np.nan)
        assert emul.emul_grid_gp_scatter(C.c_int(m), C.c_int(n), d(xs), d(ys), C.c_int(2), C.c_int(1), C.c_int(basis), C.c_int(weighted), d(x), d(out)) == 0
        close(out, ref(x))
        assert np.array_equal(out, T.gauss_op(emul, kind, adjoint, x, nout))
    nu, u = rng.random(G) + 0.5, rng.standard_normal(nd)
    out = np.full(nd, np.nan)
    assert emul.emul_grid_laplace_term(C.c_int(m), C.c_int(n), d(xs), d(ys), C.c_int(2), C.c_int(1), d(nu), d(u), d(out)) == 0
    close(out, o.laplace_term_fwd(nu, u))
    assert np.array_equal(out, T.laplace_term(emul, nu, u))


@pytest.mark.parametrize("m,n", [(5, 4), (1, 1), (9, 2)])
def test_structured_scatter_operators_mapped_grid(emul, oracle, m, n):
    """the same node bodies with the corner positions read from the coordinate array (structured connectivity on mapped + jittered node
    positions): against the oracle and bit-identical to the general adjacency-walking bodies"""
    rng = np.random.default_rng(m * 10 + n + 3)
    c, e = meshgen.tri_grid(m, n, 1.0)
    c = np.ascontiguousarray(np.stack([c[:, 0] + 0.08 * np.sin(0.7 * c[:, 1]), c[:, 1] + 0.06 * np.cos(0.9 * c[:, 0])], 1) + rng.uniform(-0.1, 0.1, c.shape))
    o = oracle.Mesh2D(c, e)
    T = HostTables(o)
    G, nd = o.ngauss, o.ndof
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    cases = [(0, 0, G, nd, o.fem_to_gauss_bwd, FEM_TO_GAUSS, True), (1, 0, G, nd, o.dof_to_gauss_bwd, DOF_TO_GAUSS, True),
             (2, 0, 2 * G, nd, o.grad_bwd, GRAD, True), (3, 0, 3 * G, 2 * nd, o.strain_bwd, STRAIN, True),
             (3, 1, 3 * G, 2 * nd, o.strain_energy_fwd, STRAIN_ENERGY, False)]
    for basis, weighted, nin, nout, ref, kind, adjoint in cases:
        x = rng.standard_normal(nin)
        out = np.full(nout, np.nan)
        assert emul.emul_grid_gp_scatter_mapped(C.c_int(m), C.c_int(n), d(c), C.c_int(2), C.c_int(1), C.c_int(basis), C.c_int(weighted), d(x), d(out)) == 0
        close(out, ref(x))
        assert np.array_equal(out, T.gauss_op(emul, kind, adjoint, x, nout))
    nu, u = rng.random(G) + 0.5, rng.standard_normal(nd)
    out = np.full(nd, np.nan)
    assert emul.emul_grid_laplace_term_mapped(C.c_int(m), C.c_int(n), d(c), C.c_int(2), C.c_int(1), d(nu), d(u), d(out)) == 0
    close(out, o.laplace_term_fwd(nu, u))
    assert np.array_equal(out, T.laplace_term(emul, nu, u))


@pytest.mark.parametrize("n,l", [(2, 2), (3, 4), (1, 1), (4, 3)])
def test_structured_tet_elasticity_warp_phases(emul, oracle, n, l):
    """tet_grid.cuh: one warp per node of Mesh3(n, n, l, h) run as loops over its lanes (shared memory poisoned), rectilinear non-uniform
    coordinates, against the canonical CSR of the oracle's 3-D stiffness op; the closed-form rows must equal the symbolic pattern."""
    rng = np.random.default_rng(n * 10 + l)
    xs = np.concatenate([[0.0], np.cumsum(rng.random(n) * 0.1 + 0.05)])
    ys = np.concatenate([[0.2], 0.2 + np.cumsum(rng.random(n) * 0.1 + 0.05)])
    zs = np.concatenate([[-0.1], -0.1 + np.cumsum(rng.random(l) * 0.1 + 0.05)])
    _, e = meshgen.tet_grid(n, n, l, 1.0)
    k, j, i = np.meshgrid(np.arange(l + 1), np.arange(n + 1), np.arange(n + 1), indexing="ij")
    c = np.stack([xs[i.reshape(-1)], ys[j.reshape(-1)], zs[k.reshape(-1)]], 1)
    o = oracle.Mesh3D(c, e)
    N3 = 3 * o.ndof
    H = rng.random(o.ngauss * 36) + 0.1
    ind, vv = o.stiffness_fwd(H)
    _, _, ref = oracle.canonical_csr(ind, vv, N3)
    li, lv = o.laplace_fwd(np.ones(o.ngauss))
    rp, _, _ = oracle.canonical_csr(li, lv, o.ndof)
    nnz_s = int(rp[-1])
    assert 9 * nnz_s == len(ref)
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    hbar = np.full(o.nelem * 36, np.nan)
    assert emul.emul_presum_coef(C.c_int(3), C.c_int(o.order), C.c_longlong(o.nelem), C.c_int(36), d(H), d(hbar)) == 0
    vals = np.full(len(ref), np.nan)
    rp64 = np.ascontiguousarray(rp, dtype=np.int64)
    rc = emul.emul_tet_grid_elast_fwd(C.c_int(n), C.c_int(l), d(xs), d(ys), d(zs), C.c_longlong(nnz_s), rp64.ctypes.data_as(C.POINTER(C.c_longlong)),
                                      d(hbar), d(vals))
    assert rc == 0, rc
    close(vals, ref)
    rp3, ci3, _ = oracle.canonical_csr(ind, vv, N3)
    dv = rng.standard_normal(len(ref))
    expect = o.stiffness_bwd(oracle.csr_adjoint_to_slots(rp3, ci3, dv, ind, N3))
    grad = np.full(o.ngauss * 36, np.nan)
    assert emul.emul_tet_grid_elast_adj(C.c_int(n), C.c_int(l), d(xs), d(ys), d(zs), C.c_int(o.order), C.c_longlong(nnz_s),
                                        rp64.ctypes.data_as(C.POINTER(C.c_longlong)), d(dv), d(grad)) == 0
    close(grad, expect)


@pytest.mark.parametrize("dim,degree", [(2, 1), (2, 2), (3, 1), (3, 2)])
def test_row_gather_forward(emul, oracle, dim, degree):
    """row_gather.cuh: one thread per dof row (adjacency walk, local-matrix row in registers, binary search of the column positions, CTA-wide
    coalesced copy) for Laplace and mass against the canonical CSR of the oracle's COO."""
    rng = np.random.default_rng(dim * 10 + degree)
    if dim == 2:
        c, e = meshgen.jitter_unstructured(17, 13, 0.05, seed=9)
        o = oracle.Mesh2D(c, e, degree=degree)
    else:
        c, e = meshgen.tet_grid(3, 3, 3, 0.3)
        o = oracle.Mesh3D(c + rng.uniform(-0.02, 0.02, c.shape), e, degree=degree)
    T = HostTables(o)
    coef = rng.random(o.ngauss) + 0.5
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for op, fwd in ((0, o.laplace_fwd), (1, o.mass_fwd)):
        ind, vv = fwd(coef)
        if dim == 3 and op == 1:
            continue                                   # the reference's 3-D mass op has its own slot layout (quirk Q5); covered by the GPU parity tests
        rp, ci, ref = oracle.canonical_csr(ind, vv, o.ndof)
        rp64, ci32 = np.ascontiguousarray(rp, dtype=np.int64), np.ascontiguousarray(ci, dtype=np.int32)
        vals = np.full(len(ref), np.nan)
        for rows in (128, 64, 32):                     # the host picks the largest row count whose CTAs fit the staging (P2 tetrahedra: 32)
            vals = np.full(len(ref), np.nan)
            rc = emul.emul_row_gather_fwd(*T._mesh_args(), *T._adj_args(), rp64.ctypes.data_as(C.POINTER(C.c_longlong)),
                                          ci32.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(op), d(coef), d(vals), C.c_int(rows))
            if rc == 0:
                break
        assert rc == 0, rc
        assert rows == (32 if (dim, degree) == (3, 2) else 128) or dim == 3
        close(vals, ref)


def test_warp_phases_do_not_depend_on_lane_order(emul, oracle):
    """The emulator runs the lanes of a phase one after the other; with the lanes visited in reverse order the structured-elasticity, tetrahedral
    and row-gather kernels must give the same bits (no lane reads what another lane writes in the same phase)."""
    rng = np.random.default_rng(99)
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    m, n = 37, 5
    xs, ys = np.arange(m + 1) * 0.1, np.arange(n + 1) * 0.07
    nnz_s = (oracle.canonical_csr(*oracle.Mesh2D(*meshgen.tri_grid(m, n, 0.1)).laplace_fwd(np.ones(6 * m * n)), (m + 1) * (n + 1))[0])[-1]
    H, dv = rng.random(2 * m * n * 27), rng.standard_normal(4 * nnz_s)
    nt, lt = 3, 2
    c3, e3 = meshgen.tet_grid(nt, nt, lt, 0.5)
    o3 = oracle.Mesh3D(c3, e3)
    rp3 = np.ascontiguousarray(oracle.canonical_csr(*o3.laplace_fwd(np.ones(o3.ngauss)), o3.ndof)[0], dtype=np.int64)
    hb, dv3 = rng.random(o3.nelem * 36), rng.standard_normal(9 * int(rp3[-1]))
    ax = [np.arange(nt + 1) * 0.5, np.arange(nt + 1) * 0.5, np.arange(lt + 1) * 0.5]
    out = {}
    for rev in (0, 1):
        emul.emul_set_reverse(C.c_int(rev))
        v, g = np.full(4 * nnz_s, np.nan), np.full(2 * m * n * 27, np.nan)
        assert emul.emul_grid_elast_fwd(C.c_int(m), C.c_int(n), d(xs), d(ys), C.c_int(2), C.c_int(0), C.c_int(3), C.c_longlong(nnz_s), d(H), d(v), C.c_int(-1), None) == 0
        assert emul.emul_grid_elast_adj(C.c_int(m), C.c_int(n), d(xs), d(ys), C.c_int(2), C.c_int(0), C.c_int(3), C.c_longlong(nnz_s), d(dv), d(g), C.c_int(-1), None, None, None) == 0
        v3, g3 = np.full(len(dv3), np.nan), np.full(o3.ngauss * 36, np.nan)
        assert emul.emul_tet_grid_elast_fwd(C.c_int(nt), C.c_int(lt), d(ax[0]), d(ax[1]), d(ax[2]), C.c_longlong(int(rp3[-1])), rp3.ctypes.data_as(C.POINTER(C.c_longlong)), d(hb), d(v3)) == 0
        assert emul.emul_tet_grid_elast_adj(C.c_int(nt), C.c_int(lt), d(ax[0]), d(ax[1]), d(ax[2]), C.c_int(2), C.c_longlong(int(rp3[-1])), rp3.ctypes.data_as(C.POINTER(C.c_longlong)), d(dv3), d(g3)) == 0
        out[rev] = [v, g, v3, g3]
    emul.emul_set_reverse(C.c_int(0))
    for a, b in zip(out[0], out[1]):
        assert not np.isnan(a).any() and np.array_equal(a, b)


@pytest.mark.parametrize("dim", [2, 3])
def test_row_gather_elasticity_forward(emul, oracle, dim):
    """row_gather.cuh, P1 elasticity on unstructured meshes: one thread per scalar row fills its NC x NC blocks from the Gauss-summed tangents;
    against the canonical CSR of the oracle's stiffness op, lanes in both orders."""
    rng = np.random.default_rng(200 + dim)
    if dim == 2:
        c, e = meshgen.jitter_unstructured(15, 11, 0.05, seed=10)
        o = oracle.Mesh2D(c, e)
    else:
        c, e = meshgen.tet_grid(4, 4, 3, 0.3)
        o = oracle.Mesh3D(c + rng.uniform(-0.02, 0.02, c.shape), e)
    T = HostTables(o)
    ns2 = 9 if dim == 2 else 36
    H = rng.random(o.ngauss * ns2) + 0.1
    ind, vv = o.stiffness_fwd(H)
    _, _, ref = oracle.canonical_csr(ind, vv, dim * o.ndof)
    rp, ci, _ = oracle.canonical_csr(*o.laplace_fwd(np.ones(o.ngauss)), o.ndof)
    rp64, ci32 = np.ascontiguousarray(rp, dtype=np.int64), np.ascontiguousarray(ci, dtype=np.int32)
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    hbar = np.full(o.nelem * ns2, np.nan)
    assert emul.emul_presum_coef(C.c_int(dim), C.c_int(o.order), C.c_longlong(o.nelem), C.c_int(ns2), d(H), d(hbar)) == 0
    res = []
    for rev in (0, 1):
        emul.emul_set_reverse(C.c_int(rev))
        vals = np.full(len(ref), np.nan)
        rc = emul.emul_row_gather_elast_fwd(C.c_int(dim), C.c_int(o.order), C.c_int(o.nnode), C.c_int(o.nelem), d(T.coords),
                                            T.verts.ctypes.data_as(C.POINTER(C.c_int)), T.conn.ctypes.data_as(C.POINTER(C.c_int)), *T._adj_args(),
                                            rp64.ctypes.data_as(C.POINTER(C.c_longlong)), ci32.ctypes.data_as(C.POINTER(C.c_int)), d(hbar), d(vals))
        assert rc == 0, rc
        close(vals, ref)
        res.append(vals)
    emul.emul_set_reverse(C.c_int(0))
    assert np.array_equal(res[0], res[1])


@pytest.mark.parametrize("n,l", [(2, 2), (3, 1)])
def test_structured_tet_scatter_operators(emul, oracle, n, l):
    """tet_gauss.cuh: scatter-type Gauss-point operators and the Laplace term on Mesh3(n, n, l, h), one thread per node with index arithmetic and
    the orientation fix replayed (Gauss points are indexed in the element's own vertex order) — bit-identical to the general adjacency-walking
    bodies, and the Laplace term against the oracle (deps/MFEM3/ComputeLaplaceTermMfem)."""
    rng = np.random.default_rng(300 + n + l)
    xs = np.concatenate([[0.0], np.cumsum(rng.random(n) * 0.1 + 0.05)])
    ys = np.concatenate([[0.2], 0.2 + np.cumsum(rng.random(n) * 0.1 + 0.05)])
    zs = np.concatenate([[-0.1], -0.1 + np.cumsum(rng.random(l) * 0.1 + 0.05)])
    _, e = meshgen.tet_grid(n, n, l, 1.0)
    k, j, i = np.meshgrid(np.arange(l + 1), np.arange(n + 1), np.arange(n + 1), indexing="ij")
    c = np.stack([xs[i.reshape(-1)], ys[j.reshape(-1)], zs[k.reshape(-1)]], 1)
    o = oracle.Mesh3D(c, e)
    T = HostTables(o)
    G, nd = o.ngauss, o.ndof
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    for basis, weighted, nin, nout, kind, adjoint in ((0, 0, G, nd, FEM_TO_GAUSS, True), (1, 0, G, nd, DOF_TO_GAUSS, True), (2, 0, 3 * G, nd, GRAD, True),
                                                      (3, 0, 6 * G, 3 * nd, STRAIN, True), (3, 1, 6 * G, 3 * nd, STRAIN_ENERGY, False)):
        x = rng.standard_normal(nin)
        out = np.full(nout, np.nan)
        assert emul.emul_tet_gp_scatter(C.c_int(n), C.c_int(l), d(xs), d(ys), d(zs), C.c_int(o.order), C.c_int(basis), C.c_int(weighted), d(x), d(out)) == 0
        assert np.array_equal(out, T.gauss_op(emul, kind, adjoint, x, nout))
    nu, u = rng.random(G) + 0.5, rng.standard_normal(nd)
    out = np.full(nd, np.nan)
    assert emul.emul_tet_laplace_term(C.c_int(n), C.c_int(l), d(xs), d(ys), d(zs), C.c_int(o.order), d(nu), d(u), d(out)) == 0
    close(out, o.laplace_term_fwd(nu, u))
    assert np.array_equal(out, T.laplace_term(emul, nu, u))


@pytest.mark.parametrize("n,l", [(2, 2), (3, 4), (1, 1)])
def test_structured_tet_scalar_operators(emul, oracle, n, l):
    """tet_scalar.cuh: CSR Laplace / mass and their adjoints on Mesh3(n, n, l, h) (the reference's own 3-D ops) by index arithmetic, against the
    oracle; mass is checked through the general row-gather body (the oracle's 3-D mass has its own COO layout) and by transposition."""
    rng = np.random.default_rng(400 + n + l)
    xs = np.concatenate([[0.0], np.cumsum(rng.random(n) * 0.1 + 0.05)])
    ys = np.concatenate([[0.2], 0.2 + np.cumsum(rng.random(n) * 0.1 + 0.05)])
    zs = np.concatenate([[-0.1], -0.1 + np.cumsum(rng.random(l) * 0.1 + 0.05)])
    _, e = meshgen.tet_grid(n, n, l, 1.0)
    k, j, i = np.meshgrid(np.arange(l + 1), np.arange(n + 1), np.arange(n + 1), indexing="ij")
    c = np.stack([xs[i.reshape(-1)], ys[j.reshape(-1)], zs[k.reshape(-1)]], 1)
    o = oracle.Mesh3D(c, e)
    T = HostTables(o)
    coef = rng.random(o.ngauss) + 0.5
    ind, vv = o.laplace_fwd(coef)
    rp, ci, ref = oracle.canonical_csr(ind, vv, o.ndof)
    rp64, ci32 = np.ascontiguousarray(rp, dtype=np.int64), np.ascontiguousarray(ci, dtype=np.int32)
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    lp = rp64.ctypes.data_as(C.POINTER(C.c_longlong))
    args = (C.c_int(n), C.c_int(l), d(xs), d(ys), d(zs), C.c_int(o.order))
    vals = np.full(len(ref), np.nan)
    assert emul.emul_tet_grid_scalar(*args, C.c_int(0), C.c_int(0), lp, d(coef), d(vals)) == 0
    close(vals, ref)
    dv = rng.standard_normal(len(ref))
    g = np.full(o.ngauss, np.nan)
    assert emul.emul_tet_grid_scalar(*args, C.c_int(0), C.c_int(1), lp, d(dv), d(g)) == 0
    close(g, o.laplace_bwd(oracle.csr_adjoint_to_slots(rp, ci, dv, ind, o.ndof)))
    # mass: forward equals the general row-gather body bit for bit; adjoint is its transpose
    mv, mref = np.full(len(ref), np.nan), np.full(len(ref), np.nan)
    assert emul.emul_tet_grid_scalar(*args, C.c_int(1), C.c_int(0), lp, d(coef), d(mv)) == 0
    assert emul.emul_row_gather_fwd(*T._mesh_args(), *T._adj_args(), lp, ci32.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(1), d(coef), d(mref), C.c_int(128)) == 0
    assert np.array_equal(mv, mref)
    gm = np.full(o.ngauss, np.nan)
    assert emul.emul_tet_grid_scalar(*args, C.c_int(1), C.c_int(1), lp, d(dv), d(gm)) == 0
    lhs, rhs = mv @ dv, gm @ coef
    assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), abs(rhs), 1.0)
    assert abs(mv.sum() - (coef * o.weights).sum()) <= 1e-12 * abs(mv.sum())          # partition of unity: sum of M = int rho


@pytest.mark.parametrize("type_", [1, 2, 3])
def test_fused_svt_quad_stiffness(emul, oracle, type_):
    """quad_ops.cuh: SpatialVaryingTangentElastic fused into UnivariateFemStiffness == the two oracle ops chained (slot order, 1-based indices,
    values), and the mu-gradient == the two oracle adjoints chained."""
    rng = np.random.default_rng(500 + type_)
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    lp = lambda a: a.ctypes.data_as(C.POINTER(C.c_longlong))
    for m, n, h in ((5, 3, 0.1), (1, 1, 2.0), (2, 7, 0.37)):
        mu = rng.random(4 * m * n * type_) + 0.5
        hmat = oracle.svt_fwd(mu, m, n, type_).reshape(4 * m * n, 2, 2)
        ri, rj, rv = oracle.univariate_stiffness_fwd(hmat, m, n, h)
        N = 64 * m * n
        ii, jj, vv = np.full(N, -1, dtype=np.int64), np.full(N, -1, dtype=np.int64), np.full(N, np.nan)
        emul.emul_quad_stiffness1_svt(d(mu), C.c_int(type_), C.c_int(m), C.c_int(n), C.c_double(h), lp(ii), lp(jj), d(vv))
        assert np.array_equal(ii, ri) and np.array_equal(jj, rj)
        close(vv, rv)
        g = rng.standard_normal(N)
        ref = oracle.svt_bwd(oracle.univariate_stiffness_bwd(g, m, n, h, True), m, n, type_)
        gmu = np.full(len(mu), np.nan)
        emul.emul_quad_stiffness1_svt_grad(d(g), C.c_int(type_), C.c_int(m), C.c_int(n), C.c_double(h), d(gmu))
        close(gmu, ref)
