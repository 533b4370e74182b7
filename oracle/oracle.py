"""ctypes front-end of the CPU oracle (oracle/adfem_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

All index arrays at this Python level are 0-based; the C symbols keep the reference's conventions
(int32 0-based mesh input, 1-based getters, 0-based int64 COO indices, 1-based `bd`), see
deps/MFEM/API.cpp:3-66 and deps/MFEM/ImposeDirichlet/ImposeDirichlet.h:32.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libadfem_oracle.so")
_lib = None

c_dp = C.POINTER(C.c_double)
c_lp = C.POINTER(C.c_longlong)
c_ip = C.POINTER(C.c_int)


def build(force=False):
    """Compile the oracle with the reference's release flags (g++ -O3 -DNDEBUG)."""
    src = os.path.join(_HERE, "adfem_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.init_nnfem_mesh.restype = c_lp
        _lib.init_nnfem_mesh3.restype = c_lp
        _lib.oracle_ImposeDirichlet_forward.restype = C.c_longlong
    return _lib


def _d(a):
    return a.ctypes.data_as(c_dp)


def _l(a):
    return a.ctypes.data_as(c_lp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_current = [None, None]  # owners of the global 2-D / 3-D singletons (quirk Q10)


class _MeshBase:
    dim = 0

    def _activate(self):
        slot = 0 if self.dim == 2 else 1
        if _current[slot] is not self:
            self._init()
            _current[slot] = self


class Mesh2D(_MeshBase):
    """2-D triangle mesh tables: deps/MFEM/Common.cpp:20-142 through the getters of API.cpp."""
    dim = 2

    def __init__(self, coords, elems, order=-1, degree=1, lorder=-1):
        self.coords = _f64(coords)
        self.elems_in = np.ascontiguousarray(elems, dtype=np.int32)
        if order == -1:
            order = 2 if degree == 1 else 4          # src/MFEM/MFEM.jl:71-77
        if lorder == -1:
            lorder = 6
        self.order, self.degree, self.lorder = order, degree, lorder
        self.nnode, self.nelem = self.coords.shape[0], self.elems_in.shape[0]
        _current[0] = None
        self._activate()

    def _init(self):
        L = lib()
        c = np.zeros((self.nnode, 3))
        c[:, :2] = self.coords
        ne = C.c_longlong(0)
        p = L.init_nnfem_mesh(_d(c), C.c_int(self.nnode), self.elems_in.ctypes.data_as(c_ip), C.c_int(self.nelem),
                              C.c_int(self.order), C.c_int(self.lorder), C.c_int(self.degree), C.byref(ne))
        self.nedge = ne.value
        e = np.ctypeslib.as_array(p, shape=(2 * max(self.nedge, 1),)).copy()
        L.oracle_free(p)
        self.edges = e[:2 * self.nedge].reshape(2, self.nedge).T - 1
        self.ngauss = L.mfem_get_ngauss()
        self.elem_ndof = L.mfem_get_elem_ndof()
        self.ndof = L.mfem_get_ndof()
        conn = np.zeros(self.nelem * self.elem_ndof, dtype=np.int64)
        L.mfem_get_connectivity(_l(conn))
        self.conn = conn.reshape(self.nelem, self.elem_ndof) - 1
        ev = np.zeros(3 * self.nelem, dtype=np.int64)
        L.mfem_get_element_to_vertices(_l(ev))
        self.elems = ev.reshape(3, self.nelem).T - 1
        x, y = np.zeros(self.ngauss), np.zeros(self.ngauss)
        L.mfem_get_gauss(_d(x), _d(y))
        self.gauss = np.stack([x, y], axis=1)
        self.weights = np.zeros(self.ngauss)
        L.mfem_get_gauss_weights(_d(self.weights))
        self.area = np.zeros(self.nelem)
        L.mfem_get_area(_d(self.area))
        self.g = self.ngauss // self.nelem

    # --- ops -------------------------------------------------------------------------
    def _coo(self, fn, coef, D):
        self._activate()
        N = self.ngauss * D * D
        ind = np.zeros(2 * N, dtype=np.int64)
        vv = np.zeros(N)
        fn(_l(ind), _d(vv), _d(_f64(coef)))
        return ind.reshape(N, 2), vv

    def laplace_fwd(self, kappa):
        return self._coo(lib().FemLaplaceScalar_forward_Julia, kappa, self.elem_ndof)

    def laplace_bwd(self, grad_vv):
        self._activate()
        g = np.zeros(self.ngauss)
        lib().oracle_FemLaplaceScalar_backward(_d(g), _d(_f64(grad_vv)))
        return g

    def laplace_jacobian(self):
        self._activate()
        N = self.ngauss * self.elem_ndof ** 2
        H = np.zeros(self.ngauss * N)
        lib().pcl_FemLaplaceScalar_Jacobian(_d(H))
        return H.reshape(N, self.ngauss).T    # column-major ngauss x N

    def mass_fwd(self, rho):
        return self._coo(lib().oracle_ComputeFemMassMatrix1_forward, rho, self.elem_ndof)

    def mass_bwd(self, grad_vv):
        self._activate()
        g = np.zeros(self.ngauss)
        lib().oracle_ComputeFemMassMatrix1_backward(_d(g), _d(_f64(grad_vv)))
        return g

    def stiffness_fwd(self, hmat):
        return self._coo(lib().ComputeFemStiffnessMatrixMfem_forward_Julia, hmat, 2 * self.elem_ndof)

    def stiffness_bwd(self, grad_vv):
        self._activate()
        g = np.zeros(9 * self.ngauss)
        lib().oracle_ComputeFemStiffnessMatrixMfem_backward(_d(g), _d(_f64(grad_vv)))
        return g

    def source_fwd(self, f):
        self._activate()
        rhs = np.zeros(self.ndof)
        lib().FemSourceScalar_forward_Julia(_d(rhs), _d(_f64(f)))
        return rhs

    def source_bwd(self, grad_rhs):
        self._activate()
        g = np.zeros(self.ngauss)
        lib().oracle_FemSourceScalar_backward(_d(g), _d(_f64(grad_rhs)))
        return g

    # --- Gauss-point operators / matrix-free terms ------------------------
    def _vec(self, name, nout, *ins):
        """Calls `name(out, *ins)` on a zero-filled output (the reference's callers zero-fill where the body accumulates)."""
        self._activate()
        out = np.zeros(nout)
        getattr(lib(), name)(_d(out), *[_d(_f64(a)) for a in ins])
        return out

    def fem_to_gauss_fwd(self, u):
        return self._vec("oracle_FemToGaussPointsMfem_forward", self.ngauss, u)

    def fem_to_gauss_bwd(self, grad_out):
        return self._vec("oracle_FemToGaussPointsMfem_backward", self.nnode, grad_out)

    def dof_to_gauss_fwd(self, u):
        return self._vec("oracle_DofToGaussPointsMfem_forward", self.ngauss, u)

    def dof_to_gauss_bwd(self, grad_out):
        return self._vec("oracle_DofToGaussPointsMfem_backward", self.ndof, grad_out)

    def grad_fwd(self, u):
        return self._vec("oracle_FemGradMfem_forward", 2 * self.ngauss, u)

    def grad_bwd(self, grad_out):
        return self._vec("oracle_FemGradMfem_backward", self.ndof, grad_out)

    def strain_fwd(self, u):
        return self._vec("oracle_EvalStrainOnGaussPts_forward", 3 * self.ngauss, u)

    def strain_bwd(self, grad_eps):
        return self._vec("oracle_EvalStrainOnGaussPts_backward", 2 * self.ndof, grad_eps)

    def strain_energy_fwd(self, sigma):
        return self._vec("oracle_ComputeStrainEnergyTermMfem_forward", 2 * self.ndof, sigma)

    def strain_energy_bwd(self, grad_out):
        return self._vec("oracle_ComputeStrainEnergyTermMfem_backward", 3 * self.ngauss, grad_out)

    def laplace_term_fwd(self, nu, u):
        return self._vec("oracle_ComputeLaplaceTermMfem_forward", self.ndof, nu, u)

    def laplace_term_bwd(self, grad_out, nu, u):
        self._activate()
        gnu, gu = np.zeros(self.ngauss), np.zeros(self.ndof)
        lib().oracle_ComputeLaplaceTermMfem_backward(_d(gnu), _d(gu), _d(_f64(grad_out)), _d(_f64(nu)), _d(_f64(u)))
        return gnu, gu


class Mesh3D(_MeshBase):
    """3-D tetrahedral mesh tables: deps/MFEM3/Common.cpp:9-148."""
    dim = 3

    def __init__(self, coords, elems, order=-1, degree=1):
        self.coords = _f64(coords)
        self.elems_in = np.ascontiguousarray(elems, dtype=np.int32)
        if order == -1:
            order = 2 if degree == 1 else 4          # src/MFEM3/MFEM.jl:49-55
        self.order, self.degree = order, degree
        self.nnode, self.nelem = self.coords.shape[0], self.elems_in.shape[0]
        _current[1] = None
        self._activate()

    def _init(self):
        L = lib()
        ne = C.c_longlong(0)
        p = L.init_nnfem_mesh3(_d(self.coords), C.c_int(self.nnode), self.elems_in.ctypes.data_as(c_ip), C.c_int(self.nelem),
                               C.c_int(self.order), C.c_int(self.degree), C.byref(ne))
        self.nedge = ne.value
        e = np.ctypeslib.as_array(p, shape=(2 * max(self.nedge, 1),)).copy()
        L.oracle_free(p)
        self.edges = e[:2 * self.nedge].reshape(2, self.nedge).T - 1
        self.ngauss = L.mfem_get_ngauss3()
        self.elem_ndof = L.mfem_get_elem_ndof3()
        self.ndof = L.mfem_get_ndof3()
        conn = np.zeros(self.nelem * self.elem_ndof, dtype=np.int64)
        L.mfem_get_connectivity3(_l(conn))
        self.conn = conn.reshape(self.nelem, self.elem_ndof) - 1
        ev = np.zeros(4 * self.nelem, dtype=np.int64)
        L.mfem_get_element_to_vertices3(_l(ev))
        self.elems = ev.reshape(4, self.nelem).T - 1
        x, y, z = np.zeros(self.ngauss), np.zeros(self.ngauss), np.zeros(self.ngauss)
        L.mfem_get_gauss3(_d(x), _d(y), _d(z))
        self.gauss = np.stack([x, y, z], axis=1)
        self.weights = np.zeros(self.ngauss)
        L.mfem_get_gauss_weights3(_d(self.weights))
        self.volume = np.zeros(self.nelem)
        L.mfem_get_volume3(_d(self.volume))
        self.g = self.ngauss // self.nelem

    def _coo(self, fn, coef, N):
        self._activate()
        ind = np.zeros(2 * N, dtype=np.int64)
        vv = np.zeros(N)
        fn(_l(ind), _d(vv), _d(_f64(coef)))
        return ind.reshape(N, 2), vv

    def laplace_fwd(self, kappa):
        return self._coo(lib().FemLaplaceScalarT_forward_Julia, kappa, self.ngauss * self.elem_ndof ** 2)

    def laplace_bwd(self, grad_vv):
        self._activate()
        g = np.zeros(self.ngauss)
        lib().oracle_FemLaplaceScalarT_backward(_d(g), _d(_f64(grad_vv)))
        return g

    def mass_fwd(self, rho):     # one slot per (e,p,q): N = nelem*d^2 (quirk Q5)
        return self._coo(lib().oracle_ComputeFemMassMatrixMfemT_forward, rho, self.nelem * self.elem_ndof ** 2)

    def mass_bwd(self, grad_vv):
        self._activate()
        g = np.zeros(self.ngauss)
        lib().oracle_ComputeFemMassMatrixMfemT_backward(_d(g), _d(_f64(grad_vv)))
        return g

    def stiffness_fwd(self, hmat):   # extension N2
        return self._coo(lib().oracle_ComputeFemStiffnessMatrixMfemT_forward, hmat, self.ngauss * (3 * self.elem_ndof) ** 2)

    def stiffness_bwd(self, grad_vv):
        self._activate()
        g = np.zeros(36 * self.ngauss)
        lib().oracle_ComputeFemStiffnessMatrixMfemT_backward(_d(g), _d(_f64(grad_vv)))
        return g

    def source_fwd(self, f):
        self._activate()
        rhs = np.zeros(self.ndof)
        lib().FemSourceScalarT_forward_Julia(_d(rhs), _d(_f64(f)))
        return rhs

    def source_bwd(self, grad_rhs):
        self._activate()
        g = np.zeros(self.ngauss)
        lib().oracle_FemSourceScalarT_backward(_d(g), _d(_f64(grad_rhs)))
        return g

    def laplace_term_fwd(self, nu, u):
        self._activate()
        out = np.zeros(self.ndof)
        lib().oracle_ComputeLaplaceTermMfemT_forward(_d(out), _d(_f64(nu)), _d(_f64(u)))
        return out

    def laplace_term_bwd(self, grad_out, nu, u):
        self._activate()
        gnu, gu = np.zeros(self.ngauss), np.zeros(self.ndof)
        lib().oracle_ComputeLaplaceTermMfemT_backward(_d(gnu), _d(gu), _d(_f64(grad_out)), _d(_f64(nu)), _d(_f64(u)))
        return gnu, gu

    def shape_tables(self):
        """h, hx, hy, hz of every element, each [nelem, elem_ndof, g] (extension helper, see oracle_shape_tables3)."""
        self._activate()
        n = self.nelem * self.elem_ndof * self.g
        t = [np.zeros(n) for _ in range(4)]
        lib().oracle_shape_tables3(*[_d(a) for a in t])
        return [a.reshape(self.nelem, self.elem_ndof, self.g) for a in t]


# --- mesh-free ops -----------------------------------------------------------------------
def dirichlet_bd_fwd(ii, jj, vv, bd, m, n):
    """deps/DirichletBd/DirichletBd.h:8-60 -> (ii1, jj1, vv1, ii2, jj2, vv2); bd int32 in the index base of ii / jj."""
    ii, jj = np.ascontiguousarray(ii, dtype=np.int64), np.ascontiguousarray(jj, dtype=np.int64)
    vv, bd = _f64(vv), np.ascontiguousarray(bd, dtype=np.int32)
    n1, n2 = C.c_longlong(0), C.c_longlong(0)
    lib().oracle_DirichletBd_forward(_l(ii), _l(jj), _d(vv), C.c_int(len(vv)), bd.ctypes.data_as(c_ip), C.c_int(len(bd)), C.c_int(m), C.c_int(n),
                                     C.byref(n1), C.byref(n2))
    out = [np.zeros(n1.value, dtype=np.int64), np.zeros(n1.value, dtype=np.int64), np.zeros(n1.value),
           np.zeros(n2.value, dtype=np.int64), np.zeros(n2.value, dtype=np.int64), np.zeros(n2.value)]
    lib().oracle_DirichletBd_copy(_l(out[0]), _l(out[1]), _d(out[2]), _l(out[3]), _l(out[4]), _d(out[5]))
    return out


def dirichlet_bd_bwd(ii, jj, g1, g2, bd, m, n):
    ii, jj = np.ascontiguousarray(ii, dtype=np.int64), np.ascontiguousarray(jj, dtype=np.int64)
    bd = np.ascontiguousarray(bd, dtype=np.int32)
    g = np.zeros(len(ii))
    lib().oracle_DirichletBd_backward(_d(g), _l(ii), _l(jj), _d(_f64(g1)), _d(_f64(g2)), C.c_int(len(ii)), bd.ctypes.data_as(c_ip), C.c_int(len(bd)),
                                      C.c_int(m), C.c_int(n))
    return g


def impose_dirichlet_fwd(indices, vv, bd0, rhs, bdval):
    """deps/MFEM/ImposeDirichlet/ImposeDirichlet.h:27-60. `bd0` is 0-based here (the C symbol takes 1-based)."""
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    vv, rhs, bdval = _f64(vv), _f64(rhs), _f64(bdval)
    bd = np.ascontiguousarray(np.asarray(bd0, dtype=np.int64) + 1)
    S = lib().oracle_ImposeDirichlet_forward(_l(indices), _d(vv), _l(bd), _d(rhs), _d(bdval), C.c_int(len(rhs)), C.c_int(len(bd)),
                                             C.c_int(len(vv)))
    oi, ov, orhs = np.zeros(2 * S, dtype=np.int64), np.zeros(S), np.zeros(len(rhs))
    lib().oracle_ImposeDirichlet_copy(_l(oi), _d(ov), _d(orhs))
    return oi.reshape(S, 2), ov, orhs


def impose_dirichlet_bwd(grad_ov, grad_orhs, indices, vv, bd0, bdval, N):
    indices = np.ascontiguousarray(indices, dtype=np.int64)
    vv, bdval, grad_ov, grad_orhs = _f64(vv), _f64(bdval), _f64(grad_ov), _f64(grad_orhs)
    bd = np.ascontiguousarray(np.asarray(bd0, dtype=np.int64) + 1)
    gv, gr, gb = np.zeros(len(vv)), np.zeros(N), np.zeros(len(bd))
    lib().oracle_ImposeDirichlet_backward(_d(gv), _d(gr), _d(gb), _d(grad_ov), _d(grad_orhs), _l(indices), _d(vv), _l(bd), _d(bdval),
                                          C.c_int(N), C.c_int(len(bd)), C.c_int(len(vv)))
    return gv, gr, gb


def plane_matrix_fwd(E, nu, mode):
    """deps/MFEM/PlaneStrainAndStress/PlaneStrainAndStress.h: mode 0 = PlaneStrainMatrix, 1 = PlaneStressMatrix; [N,3,3]."""
    E, nu = _f64(E), _f64(nu)
    out = np.zeros(9 * len(E))
    lib().oracle_PlaneMatrix_forward(_d(out), _d(E), _d(nu), C.c_int(len(E)), C.c_int(mode))
    return out.reshape(-1, 3, 3)


def plane_matrix_bwd(grad_out, E, nu, mode):
    E, nu, grad_out = _f64(E), _f64(nu), _f64(grad_out)
    gnu, gE = np.zeros(len(E)), np.zeros(len(E))
    lib().oracle_PlaneMatrix_backward(_d(gnu), _d(gE), _d(grad_out), _d(E), _d(nu), C.c_int(len(E)), C.c_int(mode))
    return gE, gnu


def _quad(fn_name, nslot, hmat, m, n, h, extra=()):
    ii, jj, vv = np.zeros(nslot, dtype=np.int64), np.zeros(nslot, dtype=np.int64), np.zeros(nslot)
    getattr(lib(), fn_name)(_l(ii), _l(jj), _d(vv), _d(_f64(hmat)), C.c_int(m), C.c_int(n), C.c_double(h), *extra)
    return ii, jj, vv


def univariate_stiffness_fwd(hmat, m, n, h):
    """deps/FemStiffness1/UnivariateFemStiffness.h; hmat [4mn,2,2] or [2,2]; ii/jj are 1-based like the op."""
    hmat = _f64(hmat)
    rank3 = 1 if hmat.ndim == 3 else 0
    return _quad("oracle_UnivariateFemStiffness_forward", 64 * m * n, hmat, m, n, h, (C.c_int(rank3),))


def univariate_stiffness_bwd(grad_vv, m, n, h, rank3):
    g = np.zeros(16 * m * n if rank3 else 4)
    lib().oracle_UnivariateFemStiffness_backward(_d(g), _d(_f64(grad_vv)), C.c_int(m), C.c_int(n), C.c_double(h), C.c_int(int(rank3)))
    return g


def fem_stiffness_fwd(hmat, m, n, h):
    return _quad("oracle_FemStiffness_forward", 64 * m * n, hmat, m, n, h)


def fem_stiffness_bwd(grad_vv, m, n, h):
    g = np.zeros(9)
    lib().oracle_FemStiffness_backward(_d(g), _d(_f64(grad_vv)), C.c_int(m), C.c_int(n), C.c_double(h))
    return g


def spatial_stiffness_fwd(hmat, m, n, h):
    return _quad("oracle_SpatialFemStiffness_forward", 256 * m * n, hmat, m, n, h)


def spatial_stiffness_bwd(grad_vv, m, n, h):
    g = np.zeros(36 * m * n)
    lib().oracle_SpatialFemStiffness_backward(_d(g), _d(_f64(grad_vv)), C.c_int(m), C.c_int(n), C.c_double(h))
    return g


def svt_fwd(mu, m, n, type_):
    out = np.zeros(16 * m * n)
    lib().oracle_SVT_forward(_d(out), _d(_f64(mu)), C.c_longlong(m), C.c_longlong(n), C.c_int(type_))
    return out


def svt_bwd(grad_hmat, m, n, type_):
    g = np.zeros(4 * m * n * type_)
    lib().oracle_SVT_backward(_d(g), _d(_f64(grad_hmat)), C.c_longlong(m), C.c_longlong(n), C.c_int(type_))
    return g


def segment_rule(order):
    n = C.c_int(0)
    p, w = np.zeros(64), np.zeros(64)
    lib().oracle_segment_rule(C.c_int(order), C.byref(n), _d(p), _d(w))
    return p[:n.value].copy(), w[:n.value].copy()


# --- the "reference CSR" (SURVEY.md §8c working definition) -----------------------------------
def canonical_csr(indices, vv, n):
    """Sort COO by (row, col), sum duplicates, keep structural zeros.  Returns (rowptr, colind, vals)."""
    indices = np.asarray(indices, dtype=np.int64)
    key = indices[:, 0] * np.int64(n) + indices[:, 1]
    order = np.argsort(key, kind="stable")
    ks = key[order]
    first = np.concatenate([[True], ks[1:] != ks[:-1]]) if len(ks) else np.zeros(0, dtype=bool)
    starts = np.flatnonzero(first)
    vals = np.add.reduceat(np.asarray(vv, dtype=np.float64)[order], starts) if len(ks) else np.zeros(0)
    uk = ks[starts]
    rows, cols = uk // n, uk % n
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, rows + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, cols.astype(np.int64), vals


def csr_adjoint_to_slots(rowptr, colind, dvals, indices, n):
    """Pull a per-nnz upstream gradient back to per-COO-slot gradients (d vals[nnz] / d vv[slot] = 1 on its slot)."""
    key_nnz = np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr)) * np.int64(n) + colind
    key = np.asarray(indices[:, 0], dtype=np.int64) * np.int64(n) + indices[:, 1]
    pos = np.searchsorted(key_nnz, key)
    return np.asarray(dvals)[pos]


def _quad_scalar(name, coef, m, n, h):
    """FemLaplace / FemMass (deps/FemLaplace/FemLaplace.h, deps/FemMass/FemMass.h): coef [4mn]; ii / jj 0-based like the ops emit them."""
    N = 64 * m * n
    ii, jj, vv = np.zeros(N, dtype=np.int64), np.zeros(N, dtype=np.int64), np.zeros(N)
    getattr(lib(), name)(_l(ii), _l(jj), _d(vv), _d(_f64(coef)), C.c_int(m), C.c_int(n), C.c_double(h))
    return ii, jj, vv


def _quad_vec(name, x, nout, m, n, h):
    out = np.zeros(nout)
    getattr(lib(), name)(_d(out), _d(_f64(x)), C.c_int(m), C.c_int(n), C.c_double(h))
    return out


def quad_laplace_fwd(K, m, n, h):
    return _quad_scalar("oracle_FemLaplace_forward", K, m, n, h)


def quad_laplace_bwd(grad_vv, m, n, h):
    return _quad_vec("oracle_FemLaplace_backward", grad_vv, 4 * m * n, m, n, h)


def quad_mass_fwd(rho, m, n, h):
    return _quad_scalar("oracle_FemMass_forward", rho, m, n, h)


def quad_mass_bwd(grad_vv, m, n, h):
    return _quad_vec("oracle_FemMass_backward", grad_vv, 4 * m * n, m, n, h)


def quad_source_fwd(f, m, n, h):
    """deps/FemSource/FemSource.h:8-26."""
    return _quad_vec("oracle_FemSource_forward", f, (m + 1) * (n + 1), m, n, h)


def quad_source_bwd(grad_rhs, m, n, h):
    return _quad_vec("oracle_FemSource_backward", grad_rhs, 4 * m * n, m, n, h)
