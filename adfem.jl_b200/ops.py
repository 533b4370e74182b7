"""Host-side mirror of the reference's assembly operators (same names and argument meaning).

Reference: src/MFEM/MCore.jl:65-82 (`compute_fem_source_term1`), :100-120 (`compute_fem_laplace_matrix1`),
:170-181 (`compute_fem_mass_matrix1`), :275-315 (`compute_fem_stiffness_matrix`); 3-D twins in
src/MFEM3/MCore.jl:22-126.

Two call styles, like the reference's eager-vs-graph split:
  * numpy arrays in  -> `scipy.sparse.csr_matrix` / numpy vector out (the reference's `Array` methods that
    return a `SparseMatrixCSC`); computed on the GPU through the CSR fast path with host<->device copies.
  * CUDA `torch.Tensor` in -> differentiable result on the device (the reference's `PyObject` methods that
    return a `SparseTensor` inside the TF graph).  `mode="coo"` gives the reference op's exact output
    (`SparseTensor` with duplicate entries, one block per Gauss point); `mode="csr"` gives the summed CSR
    values.  Gradients flow through `torch.autograd` instead of ADCME's `load_op_and_grad`.
Everything runs in libadfem_cuda.so; there is no CPU implementation behind these functions.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib
from ._lib import OP_LAPLACE, OP_MASS, OP_STIFFNESS, check, lib


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ncomp(mesh, op):
    return mesh.dim if op == OP_STIFFNESS else 1


def _coef_len(mesh, op):
    if op == OP_STIFFNESS:
        return mesh.ngauss * (9 if mesh.dim == 2 else 36)
    return mesh.ngauss


def _check_coef(t, mesh, op):
    if t.dtype != torch.float64 or not t.is_cuda:
        raise TypeError("coefficients must be a float64 CUDA tensor")
    if t.numel() != _coef_len(mesh, op):
        raise AssertionError(f"coefficient length {t.numel()} != {_coef_len(mesh, op)}")     # the reference @asserts length == ngauss
    return t.contiguous().view(-1)


class SparseTensor:
    """COO sparse matrix with duplicates, the analogue of ADCME's `RawSparseTensor(indices, vv, n, n)`
    (src/MFEM/MCore.jl:106): `indices` int64 [N,2] 0-based (mesh-static), `values` float64 [N] (differentiable)."""

    def __init__(self, indices, values, m, n):
        self.indices, self.values, self.shape = indices, values, (m, n)

    def to_scipy(self):
        i = self.indices.cpu().numpy()
        return sp.coo_matrix((self.values.detach().cpu().numpy(), (i[:, 0], i[:, 1])), shape=self.shape).tocsr()


class CSRTensor:
    """Summed CSR matrix: mesh-static `rowptr` (int64) / `colind` (int32) on the host, differentiable `values`."""

    def __init__(self, rowptr, colind, values, n):
        self.rowptr, self.colind, self.values, self.shape = rowptr, colind, values, (n, n)

    def to_scipy(self):
        return sp.csr_matrix((self.values.detach().cpu().numpy(), self.colind, self.rowptr), shape=self.shape)


class _AssembleCSR(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coef, mesh, op):
        coef = _check_coef(coef, mesh, op)
        nnz = lib().adfem_csr_nnz(mesh.handle, C.c_int(_ncomp(mesh, op)))
        if nnz < 0:
            raise _lib.AdfemError(_lib.last_error())
        vals = torch.empty(nnz, dtype=torch.float64, device=coef.device)
        check(lib().adfem_assemble_csr(mesh.handle, C.c_int(op), _ptr(coef), _ptr(vals), _stream()))
        ctx.mesh, ctx.op, ctx.shape = mesh, op, coef.shape
        return vals

    @staticmethod
    def backward(ctx, dvals):
        dvals = dvals.contiguous()
        g = torch.empty(ctx.shape, dtype=torch.float64, device=dvals.device)
        check(lib().adfem_assemble_csr_adjoint(ctx.mesh.handle, C.c_int(ctx.op), _ptr(dvals), _ptr(g), _stream()))
        return g, None, None


class _AssembleCOO(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coef, mesh, op):
        coef = _check_coef(coef, mesh, op)
        N = lib().adfem_coo_nslots(mesh.handle, C.c_int(op))
        vv = torch.empty(N, dtype=torch.float64, device=coef.device)
        check(lib().adfem_assemble_coo(mesh.handle, C.c_int(op), _ptr(coef), _ptr(vv), _stream()))
        ctx.mesh, ctx.op, ctx.shape = mesh, op, coef.shape
        return vv

    @staticmethod
    def backward(ctx, grad_vv):
        grad_vv = grad_vv.contiguous()
        g = torch.empty(ctx.shape, dtype=torch.float64, device=grad_vv.device)
        check(lib().adfem_assemble_coo_adjoint(ctx.mesh.handle, C.c_int(ctx.op), _ptr(grad_vv), _ptr(g), _stream()))
        return g, None, None


class _Source(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, mesh):
        f = _check_coef(f, mesh, OP_LAPLACE)
        rhs = torch.empty(mesh.ndof, dtype=torch.float64, device=f.device)
        check(lib().adfem_source(mesh.handle, _ptr(f), _ptr(rhs), _stream()))
        ctx.mesh = mesh
        return rhs

    @staticmethod
    def backward(ctx, grad_rhs):
        grad_rhs = grad_rhs.contiguous()
        g = torch.empty(ctx.mesh.ngauss, dtype=torch.float64, device=grad_rhs.device)
        check(lib().adfem_source_adjoint(ctx.mesh.handle, _ptr(grad_rhs), _ptr(g), _stream()))
        return g, None


def coo_indices(mesh, op, device="cuda"):
    """Mesh-static COO indices of an operator's slots, int64 [N,2], 0-based (cached per mesh)."""
    cache = mesh.__dict__.setdefault("_coo_idx", {})
    if op not in cache:
        N = lib().adfem_coo_nslots(mesh.handle, C.c_int(op))
        idx = torch.empty((N, 2), dtype=torch.int64, device=device)
        check(lib().adfem_coo_indices(mesh.handle, C.c_int(op), _ptr(idx), _stream()))
        cache[op] = idx
    return cache[op]


def _assemble(coef, mesh, op, mode):
    n = _ncomp(mesh, op) * mesh.ndof
    if isinstance(coef, np.ndarray):                          # eager path of the reference: SparseMatrixCSC out
        coef = np.ascontiguousarray(coef, dtype=np.float64).reshape(-1)
        assert coef.size == _coef_len(mesh, op)
        rowptr, colind = mesh.csr_pattern(_ncomp(mesh, op))
        vals = np.empty(rowptr[-1])
        check(lib().adfem_assemble_csr_host(mesh.handle, C.c_int(op), coef.ctypes.data_as(_lib.c_dp), vals.ctypes.data_as(_lib.c_dp)))
        return sp.csr_matrix((vals, colind, rowptr), shape=(n, n))
    if mode == "coo":
        vv = _AssembleCOO.apply(coef, mesh, op)
        return SparseTensor(coo_indices(mesh, op, coef.device), vv, n, n)
    if mode == "csr":
        vals = _AssembleCSR.apply(coef, mesh, op)
        rowptr, colind = mesh.csr_pattern(_ncomp(mesh, op))
        return CSRTensor(rowptr, colind, vals, n)
    raise ValueError("mode must be 'coo' or 'csr'")


def compute_fem_laplace_matrix1(kappa, mesh, *grid, mode="coo"):
    """`∫ κ ∇u·∇v` — src/MFEM/MCore.jl:100-120 (2-D), src/MFEM3/MCore.jl (3-D). `kappa` has one value per Gauss point.
    `compute_fem_laplace_matrix1(K, m, n, h)` is the structured Q1 sibling (src/InvCore.jl:443-449, op FemLaplace)."""
    if grid:
        return _quad_scalar(kappa, 0, mesh, *grid)
    return _assemble(kappa, mesh, OP_LAPLACE, mode)


def compute_fem_mass_matrix1(rho, mesh=None, *grid, mode="coo"):
    """`∫ ρ u v` — src/MFEM/MCore.jl:170-181. `compute_fem_mass_matrix1(mesh)` uses ρ ≡ 1 like the reference.
    `compute_fem_mass_matrix1(rho, m, n, h)` is the structured Q1 sibling (src/InvCore.jl:364-369, op FemMass)."""
    if grid:
        return _quad_scalar(rho, 1, mesh, *grid)
    if mesh is None:
        mesh, rho = rho, None
    if rho is None:
        rho = np.ones(mesh.ngauss)
    return _assemble(rho, mesh, OP_MASS, mode)


def compute_fem_stiffness_matrix(kappa, mesh, mode="coo"):
    """Elasticity stiffness `∫ ε(v):H:ε(u)` — src/MFEM/MCore.jl:275-315.  `kappa` is one ns×ns matrix or ngauss×ns×ns
    (ns = 3 in 2-D; the 3-D extension takes 6×6 Voigt matrices).  Returns a (dim·ndof)² matrix."""
    ns = 3 if mesh.dim == 2 else 6
    if isinstance(kappa, np.ndarray):
        if kappa.ndim == 2:
            assert kappa.shape == (ns, ns)
            kappa = np.broadcast_to(kappa, (mesh.ngauss, ns, ns))          # MCore.jl:287-294
        assert kappa.shape == (mesh.ngauss, ns, ns)
        kappa = np.ascontiguousarray(kappa).reshape(-1)
    else:
        if kappa.dim() == 2:
            kappa = kappa.reshape(1, ns, ns).expand(mesh.ngauss, ns, ns)   # MCore.jl:276-278
        assert tuple(kappa.shape) == (mesh.ngauss, ns, ns)
        kappa = kappa.reshape(-1)
    return _assemble(kappa, mesh, OP_STIFFNESS, mode)


def compute_fem_source_term1(f, mesh, *grid):
    """`∫ f v` — src/MFEM/MCore.jl:65-82.  `compute_fem_source_term1(f, m, n, h)` is the structured Q1 sibling
    (src/InvCore.jl:307-312, op FemSource)."""
    if grid:
        return _quad_source(f, mesh, *grid)
    if isinstance(f, np.ndarray):
        f = np.ascontiguousarray(f, dtype=np.float64)
        assert f.size == mesh.ngauss
        return compute_fem_source_term1(torch.from_numpy(f).cuda(), mesh).cpu().numpy()
    return _Source.apply(f, mesh)


def compute_fem_source_term(f1, f2, mesh):
    """src/MFEM/MCore.jl:87-89."""
    a, b = compute_fem_source_term1(f1, mesh), compute_fem_source_term1(f2, mesh)
    return np.concatenate([a, b]) if isinstance(a, np.ndarray) else torch.cat([a, b])


# =====================================================================================================
# Gauss-point operators and matrix-free terms (csrc/gauss_ops.cu)
# =====================================================================================================
GP_FEM_TO_GAUSS, GP_DOF_TO_GAUSS, GP_GRAD, GP_STRAIN, GP_STRAIN_ENERGY = range(5)


def _gp_len(mesh, kind, output):
    n = lib().adfem_gauss_op_len(mesh.handle, C.c_int(kind), C.c_int(output))
    if n < 0:
        raise _lib.AdfemError(_lib.last_error())
    return n


class _GaussOp(torch.autograd.Function):
    """A linear dof <-> Gauss-point map and its transpose (adfem_gauss_op / adfem_gauss_op_adjoint)."""

    @staticmethod
    def forward(ctx, x, mesh, kind):
        if x.dtype != torch.float64 or not x.is_cuda:
            raise TypeError("input must be a float64 CUDA tensor")
        x = x.contiguous().view(-1)
        if x.numel() != _gp_len(mesh, kind, 0):
            raise AssertionError(f"input length {x.numel()} != {_gp_len(mesh, kind, 0)}")      # the reference @asserts the lengths
        out = torch.empty(_gp_len(mesh, kind, 1), dtype=torch.float64, device=x.device)
        check(lib().adfem_gauss_op(mesh.handle, C.c_int(kind), _ptr(x), _ptr(out), _stream()))
        ctx.mesh, ctx.kind = mesh, kind
        return out

    @staticmethod
    def backward(ctx, grad_out):
        grad_out = grad_out.contiguous()
        g = torch.empty(_gp_len(ctx.mesh, ctx.kind, 0), dtype=torch.float64, device=grad_out.device)
        check(lib().adfem_gauss_op_adjoint(ctx.mesh.handle, C.c_int(ctx.kind), _ptr(grad_out), _ptr(g), _stream()))
        return g, None, None


def _gauss_op(x, mesh, kind):
    if isinstance(x, np.ndarray):          # the reference's eager `Array` methods
        return _GaussOp.apply(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).reshape(-1)).cuda(), mesh, kind).cpu().numpy()
    return _GaussOp.apply(x, mesh, kind)


def fem_to_gauss_points(u, mesh):
    """Vertex values -> Gauss points with the linear shapes — src/MFEM/MUtils.jl:231-248 (op FemToGaussPointsMfem,
    deps/MFEM/FemToGaussPoints/FemToGaussPointsMfem.h:6-35).  Only the first nnode entries of `u` are read, like the reference."""
    nv = _gp_len(mesh, GP_FEM_TO_GAUSS, 0)
    u = u.reshape(-1)
    assert u.shape[0] >= nv
    return _gauss_op(u[:nv], mesh, GP_FEM_TO_GAUSS)


def dof_to_gauss_points(u, mesh):
    """All dof values (edge dofs included for P2) -> Gauss points — src/MFEM/MUtils.jl:256-270 (op DofToGaussPointsMfem)."""
    return _gauss_op(u, mesh, GP_DOF_TO_GAUSS)


def eval_grad_on_gauss_pts1(u, mesh):
    """Gradient of a scalar dof field at the Gauss points, ngauss x dim — src/MFEM/MCore.jl:239-246 (op FemGradMfem)."""
    return _gauss_op(u, mesh, GP_GRAD).reshape(mesh.ngauss, mesh.dim)


def eval_strain_on_gauss_pts(u, mesh):
    """Strain (exx, eyy, gxy) of a displacement field `u` (2 ndof, component-blocked) at the Gauss points, ngauss x 3 —
    src/MFEM/MCore.jl:834-849 (op EvalStrainOnGaussPts).  3-D meshes (extension): ngauss x 6, Voigt xx, yy, zz, yz, xz, xy."""
    return _gauss_op(u, mesh, GP_STRAIN).reshape(mesh.ngauss, 3 if mesh.dim == 2 else 6)


def compute_strain_energy_term(Sigma, mesh):
    """`∫ σ : ε(δu)` for `Sigma` = ngauss x 3 rows (σ11, σ22, σ12); length 2 ndof — src/MFEM/MCore.jl:804-823
    (op ComputeStrainEnergyTermMfem)."""
    ns = 3 if mesh.dim == 2 else 6
    assert tuple(Sigma.shape) == (mesh.ngauss, ns)
    return _gauss_op(Sigma.reshape(-1), mesh, GP_STRAIN_ENERGY)


class _LaplaceTerm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, nu, mesh):
        if u.dtype != torch.float64 or nu.dtype != torch.float64 or not u.is_cuda or not nu.is_cuda:
            raise TypeError("u and nu must be float64 CUDA tensors")
        u, nu = u.contiguous().view(-1), nu.contiguous().view(-1)
        assert u.numel() == mesh.ndof and nu.numel() == mesh.ngauss                             # src/MFEM/MCore.jl:741-742
        out = torch.empty(mesh.ndof, dtype=torch.float64, device=u.device)
        check(lib().adfem_laplace_term(mesh.handle, _ptr(nu), _ptr(u), _ptr(out), _stream()))
        ctx.mesh = mesh
        ctx.save_for_backward(u, nu)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        u, nu = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        gu, gnu = torch.empty_like(u), torch.empty_like(nu)
        check(lib().adfem_laplace_term_adjoint(ctx.mesh.handle, _ptr(nu), _ptr(u), _ptr(grad_out), _ptr(gnu), _ptr(gu), _stream()))
        return gu, gnu, None


def compute_fem_laplace_term1(u, nu, mesh):
    """`∫ ν ∇u·∇δu` (the action of the Laplace matrix without forming it) — src/MFEM/MCore.jl:740-763, 3-D
    src/MFEM3/MCore.jl:6-20 (ops ComputeLaplaceTermMfem / ComputeLaplaceTermMfemT)."""
    if isinstance(u, np.ndarray) and isinstance(nu, np.ndarray):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(-1)).cuda()
        return _LaplaceTerm.apply(t(u), t(nu), mesh).cpu().numpy()
    dev = u.device if torch.is_tensor(u) else nu.device
    t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(-1)).to(dev)
    return _LaplaceTerm.apply(t(u), t(nu), mesh)


class _PlaneMatrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, E, nu, mode):
        if E.dtype != torch.float64 or nu.dtype != torch.float64 or not E.is_cuda or not nu.is_cuda:
            raise TypeError("E and nu must be float64 CUDA tensors")
        E, nu = E.contiguous().view(-1), nu.contiguous().view(-1)
        assert E.numel() == nu.numel()                                                          # src/Core.jl:780
        H = torch.empty((E.numel(), 3, 3), dtype=torch.float64, device=E.device)
        check(lib().adfem_plane_matrix(C.c_int(mode), C.c_longlong(E.numel()), _ptr(E), _ptr(nu), _ptr(H), _stream()))
        ctx.mode = mode
        ctx.save_for_backward(E, nu)
        return H

    @staticmethod
    def backward(ctx, gH):
        E, nu = ctx.saved_tensors
        gE, gnu = torch.empty_like(E), torch.empty_like(nu)
        check(lib().adfem_plane_matrix_grad(C.c_int(ctx.mode), C.c_longlong(E.numel()), _ptr(E), _ptr(nu), _ptr(gH.contiguous()), _ptr(gE),
                                            _ptr(gnu), _stream()))
        return gE, gnu, None


def _plane(E, nu, mode):
    if isinstance(E, np.ndarray) and isinstance(nu, np.ndarray):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(-1)).cuda()
        return _PlaneMatrix.apply(t(E), t(nu), mode).cpu().numpy()
    dev = E.device if torch.is_tensor(E) else nu.device
    t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(-1)).to(dev)
    return _PlaneMatrix.apply(t(E), t(nu), mode)


class _AssemblePlane(torch.autograd.Function):
    """CSR elasticity values straight from the moduli (adfem_assemble_csr_plane): H(E, nu) is never materialised."""

    @staticmethod
    def forward(ctx, E, nu, mesh, mode):
        if E.dtype != torch.float64 or nu.dtype != torch.float64 or not E.is_cuda or not nu.is_cuda:
            raise TypeError("E and nu must be float64 CUDA tensors")
        E, nu = E.contiguous().view(-1), nu.contiguous().view(-1)
        assert E.numel() == mesh.ngauss and nu.numel() == mesh.ngauss
        nnz = lib().adfem_csr_nnz(mesh.handle, C.c_int(mesh.dim))
        if nnz < 0:
            raise _lib.AdfemError(_lib.last_error())
        vals = torch.empty(nnz, dtype=torch.float64, device=E.device)
        check(lib().adfem_assemble_csr_plane(mesh.handle, C.c_int(mode), _ptr(E), _ptr(nu), _ptr(vals), _stream()))
        ctx.mesh, ctx.mode = mesh, mode
        ctx.save_for_backward(E, nu)
        return vals

    @staticmethod
    def backward(ctx, dvals):
        E, nu = ctx.saved_tensors
        gE, gnu = torch.empty_like(E), torch.empty_like(nu)
        check(lib().adfem_assemble_csr_plane_adjoint(ctx.mesh.handle, C.c_int(ctx.mode), _ptr(E), _ptr(nu), _ptr(dvals.contiguous()), _ptr(gE),
                                                     _ptr(gnu), _stream()))
        return gE, gnu, None, None


def compute_fem_stiffness_matrix_from_moduli(E, nu, mesh, plane="stress"):
    """`compute_fem_stiffness_matrix(compute_plane_stress_matrix(E, nu), mesh)` (or `..._strain_...`) in one fused device pass for P1
    triangle meshes: E, nu are CUDA tensors with one value per Gauss point; returns a `CSRTensor` whose values are
    differentiable with respect to both."""
    mode = {"strain": 0, "stress": 1}[plane]
    vals = _AssemblePlane.apply(E, nu, mesh, mode)
    rowptr, colind = mesh.csr_pattern(mesh.dim)
    return CSRTensor(rowptr, colind, vals, mesh.dim * mesh.ndof)


def compute_plane_strain_matrix(E, nu):
    """Pointwise N x 3 x 3 tangent, the reference's `PlaneStrainMatrix` (op PlaneStrainAndStress, mode 0) — src/Core.jl:777-786."""
    return _plane(E, nu, 0)


def compute_plane_stress_matrix(E, nu):
    """Pointwise N x 3 x 3 tangent, the reference's `PlaneStressMatrix` (op PlaneStrainAndStress, mode 1) — src/Core.jl:794-803."""
    return _plane(E, nu, 1)


# =====================================================================================================
# Structured-grid Q1 operators (src/InvCore.jl:67-110, 206-211) and the algebraic Dirichlet step
# =====================================================================================================
class _QuadOp(torch.autograd.Function):
    """kind 0: UnivariateFemStiffness, 1: FemStiffness / SpatialFemStiffness."""

    @staticmethod
    def forward(ctx, hmat, kind, m, n, h):
        hm = hmat.contiguous()
        if kind == 0:
            flag, nslot = int(hm.dim() == 3), 64 * m * n
            fn = lib().adfem_quad_stiffness1
        else:
            flag, nslot = int(hm.dim() == 3), (256 if hm.dim() == 3 else 64) * m * n
            fn = lib().adfem_quad_elasticity
        vv = torch.empty(nslot, dtype=torch.float64, device=hm.device)
        check(fn(_ptr(hm), C.c_int(flag), C.c_int(m), C.c_int(n), C.c_double(h), None, None, _ptr(vv), _stream()))
        ctx.args = (kind, flag, m, n, h, hm.shape)
        return vv

    @staticmethod
    def backward(ctx, grad_vv):
        kind, flag, m, n, h, shape = ctx.args
        g = torch.empty(shape, dtype=torch.float64, device=grad_vv.device)
        fn = lib().adfem_quad_stiffness1_grad if kind == 0 else lib().adfem_quad_elasticity_grad
        check(fn(_ptr(grad_vv.contiguous()), C.c_int(flag), C.c_int(m), C.c_int(n), C.c_double(h), _ptr(g), _stream()))
        return g, None, None, None, None


class _QuadScalar(torch.autograd.Function):
    """FemLaplace (op 0) / FemMass (op 1) on an m x n grid: coef [4mn] -> vv [64mn]."""

    @staticmethod
    def forward(ctx, coef, op, m, n, h):
        if coef.dtype != torch.float64 or not coef.is_cuda:
            raise TypeError("coefficients must be a float64 CUDA tensor")
        coef = coef.contiguous().view(-1)
        assert coef.numel() == 4 * m * n
        vv = torch.empty(64 * m * n, dtype=torch.float64, device=coef.device)
        check(lib().adfem_quad_scalar(C.c_int(op), _ptr(coef), C.c_longlong(m), C.c_longlong(n), C.c_double(h), None, None, _ptr(vv), _stream()))
        ctx.args = (op, m, n, h)
        return vv

    @staticmethod
    def backward(ctx, grad_vv):
        op, m, n, h = ctx.args
        g = torch.empty(4 * m * n, dtype=torch.float64, device=grad_vv.device)
        check(lib().adfem_quad_scalar_grad(C.c_int(op), _ptr(grad_vv.contiguous()), C.c_longlong(m), C.c_longlong(n), C.c_double(h), _ptr(g), _stream()))
        return g, None, None, None, None


class _QuadSource(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, m, n, h):
        if f.dtype != torch.float64 or not f.is_cuda:
            raise TypeError("f must be a float64 CUDA tensor")
        f = f.contiguous().view(-1)
        assert f.numel() == 4 * m * n
        rhs = torch.empty((m + 1) * (n + 1), dtype=torch.float64, device=f.device)
        check(lib().adfem_quad_source(_ptr(f), C.c_longlong(m), C.c_longlong(n), C.c_double(h), _ptr(rhs), _stream()))
        ctx.args = (m, n, h)
        return rhs

    @staticmethod
    def backward(ctx, grad_rhs):
        m, n, h = ctx.args
        g = torch.empty(4 * m * n, dtype=torch.float64, device=grad_rhs.device)
        check(lib().adfem_quad_source_grad(_ptr(grad_rhs.contiguous()), C.c_longlong(m), C.c_longlong(n), C.c_double(h), _ptr(g), _stream()))
        return g, None, None, None


def _quad_scalar_indices(op, m, n, h, device):
    ii = torch.empty(64 * m * n, dtype=torch.int64, device=device)
    jj = torch.empty_like(ii)
    vv = torch.empty(64 * m * n, dtype=torch.float64, device=device)
    dummy = torch.zeros(4 * m * n, dtype=torch.float64, device=device)
    check(lib().adfem_quad_scalar(C.c_int(op), _ptr(dummy), C.c_longlong(m), C.c_longlong(n), C.c_double(h), _ptr(ii), _ptr(jj), _ptr(vv), _stream()))
    return torch.stack([ii, jj], 1)                      # these ops emit 0-based indices already


def _quad_scalar(coef, op, m, n, h):
    m, n, h = int(m), int(n), float(h)
    N = (m + 1) * (n + 1)
    as_numpy = isinstance(coef, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(coef, dtype=np.float64).reshape(-1)).cuda() if as_numpy else coef
    vv = _QuadScalar.apply(t, op, m, n, h)
    S = SparseTensor(_quad_scalar_indices(op, m, n, h, vv.device), vv, N, N)
    return S.to_scipy() if as_numpy else S


def _quad_source(f, m, n, h):
    m, n, h = int(m), int(n), float(h)
    if isinstance(f, np.ndarray):
        return _QuadSource.apply(torch.from_numpy(np.ascontiguousarray(f, dtype=np.float64).reshape(-1)).cuda(), m, n, h).cpu().numpy()
    return _QuadSource.apply(f, m, n, h)


def _quad_indices(kind, flag, m, n, h, device):
    nslot = (64 if kind == 0 or not flag else 256) * m * n
    ii = torch.empty(nslot, dtype=torch.int64, device=device)
    jj = torch.empty(nslot, dtype=torch.int64, device=device)
    vv = torch.empty(nslot, dtype=torch.float64, device=device)
    dummy = torch.zeros((4 * m * n, 2, 2) if kind == 0 else ((4 * m * n, 3, 3) if flag else (3, 3)), dtype=torch.float64, device=device)
    fn = lib().adfem_quad_stiffness1 if kind == 0 else lib().adfem_quad_elasticity
    check(fn(_ptr(dummy), C.c_int(1 if kind == 0 else flag), C.c_int(m), C.c_int(n), C.c_double(h), _ptr(ii), _ptr(jj), _ptr(vv), _stream()))
    return torch.stack([ii - 1, jj - 1], 1)              # the ops emit 1-based ii/jj; SparseTensor here is 0-based


def _quad(hmat, kind, m, n, h, ncomp):
    N = ncomp * (m + 1) * (n + 1)
    as_numpy = isinstance(hmat, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(hmat, dtype=np.float64)).cuda() if as_numpy else hmat
    if t.dtype != torch.float64 or not t.is_cuda:
        raise TypeError("hmat must be float64 (numpy array or CUDA tensor)")
    vv = _QuadOp.apply(t, kind, int(m), int(n), float(h))
    idx = _quad_indices(kind, int(t.dim() == 3), int(m), int(n), float(h), t.device)
    S = SparseTensor(idx, vv, N, N)
    return S.to_scipy() if as_numpy else S


def compute_fem_stiffness_matrix1(hmat, m, n, h):
    """Scalar anisotropic stiffness `∫(K∇u)·∇v` on an m×n Q1 grid — src/InvCore.jl:67-76 (op UnivariateFemStiffness).
    `hmat` is 2×2 (constant) or 4mn×2×2 (per Gauss point)."""
    if hmat.ndim not in (2, 3):
        raise ValueError("Only 4mn x 2 x 2 or 2 x 2 `hmat` is supported.")       # InvCore.jl:68-70
    assert hmat.shape[-1] == 2 and hmat.shape[-2] == 2
    assert hmat.ndim == 2 or hmat.shape[0] == 4 * m * n
    return _quad(hmat, 0, m, n, h, 1)


def compute_fem_stiffness_matrix_grid(hmat, m, n, h):
    """Q1 elasticity stiffness on an m×n grid — src/InvCore.jl:84-110 (`compute_fem_stiffness_matrix(hmat, m, n, h)`):
    3×3 constant H → op FemStiffness, 4mn×3×3 → op SpatialFemStiffness."""
    if hmat.ndim not in (2, 3):
        raise ValueError("size hmat not valid")                                 # InvCore.jl:92
    assert hmat.shape[-1] == 3 and hmat.shape[-2] == 3
    assert hmat.ndim == 2 or hmat.shape[0] == 4 * m * n
    return _quad(hmat, 1, m, n, h, 2)


_mfem_stiffness = compute_fem_stiffness_matrix


def compute_fem_stiffness_matrix(*args, **kw):     # noqa: F811  (Julia dispatches on argument types; so do we)
    """`compute_fem_stiffness_matrix(kappa, mesh)` (src/MFEM/MCore.jl:275-315) or
    `compute_fem_stiffness_matrix(hmat, m, n, h)` (src/InvCore.jl:84-110)."""
    if len(args) == 4:
        return compute_fem_stiffness_matrix_grid(*args)
    return _mfem_stiffness(*args, **kw)


class _SVT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, m, n, type_):
        mu = mu.contiguous()
        out = torch.empty((4 * m * n, 2, 2), dtype=torch.float64, device=mu.device)
        check(lib().adfem_svt(_ptr(mu), C.c_longlong(m), C.c_longlong(n), C.c_int(type_), _ptr(out), _stream()))
        ctx.args = (m, n, type_, mu.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        m, n, type_, shape = ctx.args
        out = torch.empty(shape, dtype=torch.float64, device=g.device)
        check(lib().adfem_svt_grad(_ptr(g.contiguous()), C.c_longlong(m), C.c_longlong(n), C.c_int(type_), _ptr(out), _stream()))
        return out, None, None, None


def compute_space_varying_tangent_elasticity_matrix(mu, m, n, h, type=1):   # noqa: A002  (the reference's argument name)
    """4mn×2×2 tangent matrices from `mu` — src/InvCore.jl:206-211 (op SpatialVaryingTangentElastic)."""
    as_numpy = isinstance(mu, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(mu, dtype=np.float64)).cuda() if as_numpy else mu
    assert t.numel() == 4 * m * n * type
    out = _SVT.apply(t.view(-1), int(m), int(n), int(type))
    return out.cpu().numpy() if as_numpy else out


class _QuadStiff1SVT(torch.autograd.Function):
    """compute_fem_stiffness_matrix1(compute_space_varying_tangent_elasticity_matrix(mu, ...), ...) fused (adfem_quad_stiffness1_svt)."""

    @staticmethod
    def forward(ctx, mu, type_, m, n, h):
        if mu.dtype != torch.float64 or not mu.is_cuda:
            raise TypeError("mu must be a float64 CUDA tensor")
        mu = mu.contiguous().view(-1)
        assert mu.numel() == 4 * m * n * type_
        vv = torch.empty(64 * m * n, dtype=torch.float64, device=mu.device)
        check(lib().adfem_quad_stiffness1_svt(_ptr(mu), C.c_int(type_), C.c_int(m), C.c_int(n), C.c_double(h), None, None, _ptr(vv), _stream()))
        ctx.args = (type_, m, n, h, mu.numel())
        return vv

    @staticmethod
    def backward(ctx, grad_vv):
        type_, m, n, h, nmu = ctx.args
        g = torch.empty(nmu, dtype=torch.float64, device=grad_vv.device)
        check(lib().adfem_quad_stiffness1_svt_grad(_ptr(grad_vv.contiguous()), C.c_int(type_), C.c_int(m), C.c_int(n), C.c_double(h), _ptr(g), _stream()))
        return g, None, None, None, None


def compute_fem_stiffness_matrix1_from_mu(mu, m, n, h, type=1):   # noqa: A002
    """`compute_fem_stiffness_matrix1(compute_space_varying_tangent_elasticity_matrix(mu, m, n, h, type), m, n, h)` (src/InvCore.jl:67-76,
    206-211) in one device pass: the 4mn x 2 x 2 tangent tensor is never materialised.  Same `SparseTensor` as the two ops."""
    m, n, h, type = int(m), int(n), float(h), int(type)
    as_numpy = isinstance(mu, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(mu, dtype=np.float64).reshape(-1)).cuda() if as_numpy else mu
    vv = _QuadStiff1SVT.apply(t, type, m, n, h)
    N = (m + 1) * (n + 1)
    S = SparseTensor(_quad_indices(0, 1, m, n, h, vv.device), vv, N, N)
    return S.to_scipy() if as_numpy else S


class _Dirichlet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vv, rhs, bdval, indices, bd1):
        vv, rhs, bdval = vv.contiguous(), rhs.contiguous(), bdval.contiguous()
        sN, N, bdN = vv.numel(), rhs.numel(), bd1.numel()
        L = lib()
        L.adfem_impose_dirichlet_count.restype = C.c_longlong
        S = L.adfem_impose_dirichlet_count(_ptr(indices), C.c_longlong(sN), _ptr(bd1), C.c_longlong(bdN), C.c_longlong(N), _stream())
        if S < 0:
            raise _lib.AdfemError(_lib.last_error())
        oind = torch.empty((S, 2), dtype=torch.int64, device=vv.device)
        ov = torch.empty(S, dtype=torch.float64, device=vv.device)
        orhs = torch.empty(N, dtype=torch.float64, device=vv.device)
        check(L.adfem_impose_dirichlet(_ptr(indices), _ptr(vv), C.c_longlong(sN), _ptr(bd1), _ptr(bdval), C.c_longlong(bdN), _ptr(rhs),
                                       C.c_longlong(N), _ptr(oind), _ptr(ov), _ptr(orhs), _stream()))
        ctx.save_for_backward(vv, bdval, indices, bd1)
        ctx.N = N
        ctx.mark_non_differentiable(oind)
        return ov, orhs, oind

    @staticmethod
    def backward(ctx, g_ov, g_orhs, _g_ind):
        vv, bdval, indices, bd1 = ctx.saved_tensors
        sN, N, bdN = vv.numel(), ctx.N, bd1.numel()
        g_ov = torch.zeros(0, dtype=torch.float64, device=vv.device) if g_ov is None else g_ov.contiguous()
        g_orhs = torch.zeros(N, dtype=torch.float64, device=vv.device) if g_orhs is None else g_orhs.contiguous()
        gv, gr, gb = (torch.empty(k, dtype=torch.float64, device=vv.device) for k in (sN, N, bdN))
        check(lib().adfem_impose_dirichlet_grad(_ptr(g_ov), _ptr(g_orhs), _ptr(indices), _ptr(vv), C.c_longlong(sN), _ptr(bd1), _ptr(bdval),
                                                C.c_longlong(bdN), C.c_longlong(N), _ptr(gv), _ptr(gr), _ptr(gb), _stream()))
        return gv, gr, gb, None, None


def impose_Dirichlet_boundary_conditions(A, rhs=None, bdnode=None, bdval=None):
    """Algebraic Dirichlet conditions — src/MFEM/MUtils.jl:184-225 (op ImposeDirichlet).  `A` is a `SparseTensor`
    (device, differentiable) or a scipy sparse matrix / dense array (eager); `bdnode` holds 0-based dofs here.
    `impose_Dirichlet_boundary_conditions(A, bdnode)` is the homogeneous helper of MUtils.jl:220-225."""
    if bdnode is None:                                                     # helper form: (A, bdnode)
        bdnode, rhs = rhs, None
    eager = not isinstance(A, SparseTensor)
    if eager:
        M = sp.coo_matrix(A)
        dev_ = torch.device("cuda")
        A = SparseTensor(torch.from_numpy(np.stack([M.row, M.col], 1).astype(np.int64)).to(dev_), torch.from_numpy(M.data.astype(np.float64)).to(dev_),
                         *M.shape)
    dev_ = A.values.device
    N = A.shape[0]
    assert A.shape[0] == A.shape[1]
    bd = torch.as_tensor(np.asarray(bdnode), dtype=torch.int64, device=dev_)
    helper = rhs is None
    rhs_t = torch.zeros(N, dtype=torch.float64, device=dev_) if helper else torch.as_tensor(rhs, dtype=torch.float64, device=dev_)
    bdval_t = torch.zeros(bd.numel(), dtype=torch.float64, device=dev_) if bdval is None else torch.as_tensor(bdval, dtype=torch.float64, device=dev_)
    assert rhs_t.numel() == N and bdval_t.numel() == bd.numel() and bd.numel() <= N            # MUtils.jl:205-207
    ov, orhs, oind = _Dirichlet.apply(A.values, rhs_t, bdval_t, A.indices.contiguous(), (bd + 1).contiguous())
    B = SparseTensor(oind, ov, N, N)
    if eager:
        B, orhs = B.to_scipy(), orhs.cpu().numpy()
    return B if helper else (B, orhs)


class _DirichletBd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vv, ii, jj, bd, m, n):
        vv = vv.contiguous()
        N, L = vv.numel(), lib()
        n1, n2 = C.c_longlong(0), C.c_longlong(0)
        check(L.adfem_dirichlet_bd_count(_ptr(ii), _ptr(jj), C.c_longlong(N), _ptr(bd), C.c_int(bd.numel()), C.c_int(m), C.c_int(n), C.byref(n1), C.byref(n2),
                                         _stream()))
        i64 = lambda k: torch.empty(k, dtype=torch.int64, device=vv.device)
        f64 = lambda k: torch.empty(k, dtype=torch.float64, device=vv.device)
        ii1, jj1, vv1, ii2, jj2, vv2 = i64(n1.value), i64(n1.value), f64(n1.value), i64(n2.value), i64(n2.value), f64(n2.value)
        check(L.adfem_dirichlet_bd(_ptr(ii), _ptr(jj), _ptr(vv), C.c_longlong(N), _ptr(bd), C.c_int(bd.numel()), C.c_int(m), C.c_int(n), _ptr(ii1), _ptr(jj1),
                                   _ptr(vv1), _ptr(ii2), _ptr(jj2), _ptr(vv2), _stream()))
        ctx.save_for_backward(ii, jj, bd)
        ctx.mn = (m, n, n1.value, n2.value)
        ctx.mark_non_differentiable(ii1, jj1, ii2, jj2)
        return vv1, vv2, ii1, jj1, ii2, jj2

    @staticmethod
    def backward(ctx, g1, g2, *_):
        ii, jj, bd = ctx.saved_tensors
        m, n, n1, n2 = ctx.mn
        g1 = torch.zeros(n1, dtype=torch.float64, device=ii.device) if g1 is None else g1.contiguous()
        g2 = torch.zeros(n2, dtype=torch.float64, device=ii.device) if g2 is None else g2.contiguous()
        g = torch.empty(ii.numel(), dtype=torch.float64, device=ii.device)
        check(lib().adfem_dirichlet_bd_grad(_ptr(ii), _ptr(jj), C.c_longlong(ii.numel()), _ptr(bd), C.c_int(bd.numel()), C.c_int(m), C.c_int(n), _ptr(g1),
                                            _ptr(g2), _ptr(g), _stream()))
        return g, None, None, None, None, None


def fem_impose_Dirichlet_boundary_condition_experimental(A, bdnode, m, n, h=None, coupled=False):
    """`fem_impose_Dirichlet_boundary_condition_experimental(A, bdnode, m, n, h)` — src/InvCore.jl:14-23 (op DirichletBd,
    deps/DirichletBd/DirichletBd.h:8-60); `coupled=True` is `fem_impose_coupled_Dirichlet_boundary_condition` (:6-12, m*n extra dofs).
    `A`: SparseTensor with the 2(m+1)(n+1) component-blocked dofs; `bdnode`: boundary NODES in the index base of A's indices (the op
    compares integers only).  Returns (A1, A2): A1 = A with the rows and columns of the boundary dofs (bdnode and bdnode + (m+1)(n+1))
    replaced by the identity, A2 = the free-row / boundary-column block with columns numbered 1..2|bdnode| in list order."""
    assert isinstance(A, SparseTensor)
    dev_ = A.values.device
    ind = A.indices
    ii, jj = ind[:, 0].contiguous(), ind[:, 1].contiguous()
    bd = torch.as_tensor(np.asarray(bdnode), dtype=torch.int32, device=dev_).contiguous()
    vv1, vv2, ii1, jj1, ii2, jj2 = _DirichletBd.apply(A.values, ii, jj, bd, int(m), int(n))
    nd = 2 * (m + 1) * (n + 1) + (m * n if coupled else 0)
    return SparseTensor(torch.stack([ii1, jj1], 1), vv1, nd, nd), SparseTensor(torch.stack([ii2, jj2], 1), vv2, nd, 2 * bd.numel())


# ------------------------------------------------------------------------------------------ PCL Jacobians (src/pcl.jl)
def pcl_compute_fem_laplace_matrix1(mesh):
    """`pcl_compute_fem_laplace_matrix1(mmesh)` — src/pcl.jl:35-39 (kernel pcl_FemLaplaceScalar_Jacobian,
    deps/MFEM/FemLaplace1/FemLaplaceScalar.h:65-92): dense J[t, slot] = d values[slot] / d kappa[t] of the COO output of
    `compute_fem_laplace_matrix1`, shape (ngauss, elem_ndof^2 * ngauss), as a device tensor."""
    G = mesh.ngauss
    N = G * mesh.elem_ndof ** 2
    Ht = torch.zeros(N, G, dtype=torch.float64, device="cuda")            # column-major G x N == row-major N x G
    check(lib().adfem_pcl_laplace_jacobian(mesh.handle, _ptr(Ht), _stream()))
    return Ht.t()


def pcl_impose_Dirichlet_boundary_conditions(indices, bdnode, outdof):
    """`pcl_impose_Dirichlet_boundary_conditions(indices, bdnode, outdof)` — src/pcl.jl:15-22 (kernel pcl_ImposeDirichlet,
    deps/MFEM/ImposeDirichlet/ImposeDirichlet.h:98-112): J[i, j] = d (v_B)_j / d (v_A)_i, shape (n_A, outdof).  `indices` (n_A x 2) and
    `bdnode` are 0-based here (1-based in Julia)."""
    ind = torch.as_tensor(np.asarray(indices), dtype=torch.int64, device="cuda").contiguous()
    bd = torch.as_tensor(np.asarray(bdnode), dtype=torch.int64, device="cuda") + 1
    sN = ind.shape[0]
    N = int(max(ind.max().item() if sN else -1, (bd.max().item() - 1) if bd.numel() else -1)) + 1
    need = int(lib().adfem_impose_dirichlet_count(_ptr(ind), C.c_longlong(sN), _ptr(bd), C.c_longlong(bd.numel()), C.c_longlong(N), _stream()))
    if need < 0:
        raise _lib.AdfemError(_lib.last_error())
    if int(outdof) < need:      # the kernel writes column kpos[k] < nkeep of every kept slot: a smaller J would be written out of bounds
        raise ValueError("outdof = %d is smaller than the %d output slots of impose_Dirichlet_boundary_conditions" % (int(outdof), need))
    Jt = torch.zeros(int(outdof), sN, dtype=torch.float64, device="cuda")  # column-major sN x outdof
    check(lib().adfem_pcl_impose_dirichlet(_ptr(ind), C.c_longlong(sN), _ptr(bd), C.c_longlong(bd.numel()), C.c_longlong(N), _ptr(Jt), _stream()))
    return Jt.t()
