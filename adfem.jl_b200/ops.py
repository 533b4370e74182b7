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


def compute_fem_laplace_matrix1(kappa, mesh, mode="coo"):
    """`∫ κ ∇u·∇v` — src/MFEM/MCore.jl:100-120 (2-D), src/MFEM3/MCore.jl (3-D). `kappa` has one value per Gauss point."""
    return _assemble(kappa, mesh, OP_LAPLACE, mode)


def compute_fem_mass_matrix1(rho, mesh=None, mode="coo"):
    """`∫ ρ u v` — src/MFEM/MCore.jl:170-181. `compute_fem_mass_matrix1(mesh)` uses ρ ≡ 1 like the reference."""
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


def compute_fem_source_term1(f, mesh):
    """`∫ f v` — src/MFEM/MCore.jl:65-82."""
    if isinstance(f, np.ndarray):
        f = np.ascontiguousarray(f, dtype=np.float64)
        assert f.size == mesh.ngauss
        return compute_fem_source_term1(torch.from_numpy(f).cuda(), mesh).cpu().numpy()
    return _Source.apply(f, mesh)


def compute_fem_source_term(f1, f2, mesh):
    """src/MFEM/MCore.jl:87-89."""
    a, b = compute_fem_source_term1(f1, mesh), compute_fem_source_term1(f2, mesh)
    return np.concatenate([a, b]) if isinstance(a, np.ndarray) else torch.cat([a, b])
