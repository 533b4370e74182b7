"""Exact-integration golden matrices, independent of oracle/ and of the library (test infrastructure).

The reference's quantitative checks (deps/MFEM/FemLaplace1/ftest.jl:6-33 against FEniCS dumps `fenics/A.txt`, `fenics/A2.txt`
made by ftest.py:1-19 / ftest2.py; test/pcl.jl:34-50) rely on data files that are not shipped.  The matrices themselves are
exactly computable, so this script integrates them in RATIONAL arithmetic (sympy Rational / Poly) from nothing but the mesh
arrays (node coordinates, element vertex lists):

  * Lagrange P1 / P2 bases written in barycentric coordinates (vertex functions L_i(2L_i - 1), edge functions 4 L_i L_j);
  * polynomial coefficients kappa(x), rho, f(x), H(x) of a degree the reference's quadrature integrates exactly (order 2 for
    P1, order 4 for P2: src/MFEM/MFEM.jl:71-77, src/MFEM3/MFEM.jl:49-55), so the exact integral IS what the reference computes
    up to rounding;
  * monomial integrals over the reference simplex, int xi^a eta^b zeta^c = a! b! c! / (a + b + c + dim)!.

Dofs are named geometrically, so nothing here depends on MFEM's orientation fix, edge numbering or Gauss-point order:
vertex dof = node id, edge dof = nnode + rank of (min, max) vertex pair among all edges in lexicographic order.  A test maps the
oracle's / library's own edge table (`edges`) onto these names.  Elasticity uses the reference's component-blocked dofs
(deps/MFEM/ComputeFemStiffnessMatrixMfem/ComputeFemStiffnessMatrixMfem.h:17-36) and its Voigt rows [exx, eyy, 2exy]; 3-D is the
extension N2 with rows [xx, yy, zz, yz, xz, xy].

Run:  python tests/golden/make_exact_goldens.py      (writes tests/golden/exact_*.npz, ~1 minute)
"""
import itertools
import os
import sys
from fractions import Fraction
from math import factorial

import numpy as np
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

XI = sp.symbols("xi0:3")
X = sp.symbols("x0:3")


def ref_monomial_integral(exps, dim):
    num = 1
    for a in exps:
        num *= factorial(a)
    return sp.Rational(num, factorial(sum(exps) + dim))


def integrate_ref(expr, dim):
    """Exact integral of a polynomial in XI[:dim] over the reference simplex."""
    p = sp.Poly(sp.expand(expr), *XI[:dim])
    return sum((c * ref_monomial_integral(m, dim) for m, c in p.terms()), sp.Integer(0))


def local_basis(dim, degree):
    """[(label, poly in XI)]: label = ('v', i) or ('e', i, j) with local vertex numbers."""
    L = [1 - sum(XI[:dim])] + list(XI[:dim])
    if degree == 1:
        return [(("v", i), L[i]) for i in range(dim + 1)]
    out = [(("v", i), L[i] * (2 * L[i] - 1)) for i in range(dim + 1)]
    for i, j in itertools.combinations(range(dim + 1), 2):
        out.append((("e", i, j), 4 * L[i] * L[j]))
    return out


def edge_table(elems):
    pairs = set()
    for el in elems:
        for i, j in itertools.combinations(range(len(el)), 2):
            pairs.add((min(el[i], el[j]), max(el[i], el[j])))
    pairs = sorted(pairs)
    return pairs, {p: k for k, p in enumerate(pairs)}


class Element:
    def __init__(self, verts, dim, degree, nnode, edge_id):
        self.dim = dim
        v0 = sp.Matrix(verts[0])
        self.J = sp.Matrix.hstack(*[sp.Matrix(verts[i + 1]) - v0 for i in range(dim)])   # x = v0 + J xi
        self.det = abs(self.J.det())
        Jinv = self.J.inv()
        self.xmap = {X[a]: v0[a] + sum(self.J[a, b] * XI[b] for b in range(dim)) for a in range(dim)}
        self.basis = local_basis(dim, degree)
        self.phi = [b for _, b in self.basis]
        # physical gradient: d/dx_a = sum_b dxi_b/dx_a d/dxi_b
        self.grad = [[sum(Jinv[b, a] * sp.diff(p, XI[b]) for b in range(dim)) for a in range(dim)] for p in self.phi]

    def to_ref(self, expr):
        return sp.sympify(expr).subs(self.xmap, simultaneous=True)

    def integral(self, expr):
        return integrate_ref(expr, self.dim) * self.det


def assemble(coords, elems, dim, degree, kind, coef):
    """Returns (ndof, {(I, J): Rational}) or (ndof, {I: Rational}) for kind == 'source'.  coef is a sympy expression in X[:dim]
    (scalar ops) or a matrix of such (stiffness)."""
    nnode = len(coords)
    pairs, edge_id = edge_table(elems)
    n = nnode + (len(pairs) if degree == 2 else 0)
    out = {}
    for el in elems:
        E = Element([coords[v] for v in el], dim, degree, nnode, edge_id)
        gid = []
        for lab, _ in E.basis:
            if lab[0] == "v":
                gid.append(el[lab[1]])
            else:
                a, b = el[lab[1]], el[lab[2]]
                gid.append(nnode + edge_id[(min(a, b), max(a, b))])
        d = len(gid)
        if kind == "source":
            f = E.to_ref(coef)
            for p in range(d):
                out[gid[p]] = out.get(gid[p], 0) + E.integral(f * E.phi[p])
            continue
        if kind in ("laplace", "mass"):
            c = E.to_ref(coef)
            for p in range(d):
                for q in range(p, d):
                    if kind == "laplace":
                        integrand = c * sum(E.grad[p][a] * E.grad[q][a] for a in range(dim))
                    else:
                        integrand = c * E.phi[p] * E.phi[q]
                    val = E.integral(integrand)
                    for (r, s) in {(p, q), (q, p)}:
                        key = (gid[r], gid[s])
                        out[key] = out.get(key, 0) + val
            continue
        assert kind == "stiffness"
        H = coef.applyfunc(E.to_ref)
        ns = H.shape[0]

        def brow(p, comp):   # column of B for dof (p, component comp): strain rows
            g = E.grad[p]
            if dim == 2:
                return [g[0], 0, g[1]] if comp == 0 else [0, g[1], g[0]]
            # [xx, yy, zz, yz, xz, xy]
            return ([g[0], 0, 0, 0, g[2], g[1]] if comp == 0 else
                    [0, g[1], 0, g[2], 0, g[0]] if comp == 1 else [0, 0, g[2], g[1], g[0], 0])
        for p in range(d):
            for a in range(dim):
                bp = brow(p, a)
                for q in range(d):
                    for b in range(dim):
                        bq = brow(q, b)
                        integrand = sum(bp[i] * H[i, j] * bq[j] for i in range(ns) for j in range(ns) if bp[i] != 0 and bq[j] != 0)
                        key = (gid[p] + a * n, gid[q] + b * n)
                        out[key] = out.get(key, 0) + E.integral(integrand)
    return n, out


def rational_coords(c, denom):
    return [[Fraction(int(round(float(x) * denom)), denom) for x in row] for row in c]


def poly_terms(expr, dim):
    """Store a polynomial as (exponent rows, float coefficients) so that the test can evaluate it with numpy."""
    p = sp.Poly(sp.expand(expr), *X[:dim])
    exps = np.array([m for m, _ in p.terms()], dtype=np.int64).reshape(-1, dim)
    cf = np.array([float(Fraction(int(c.p), int(c.q))) for _, c in p.terms()])
    return exps, cf


def to_float(r):
    r = sp.Rational(r)
    return float(Fraction(int(r.p), int(r.q)))     # correctly rounded


def pack_matrix(d):
    keys = sorted(d)
    rows = np.array([k[0] for k in keys], dtype=np.int64)
    cols = np.array([k[1] for k in keys], dtype=np.int64)
    vals = np.array([to_float(d[k]) for k in keys])
    keep = np.array([d[k] != 0 for k in keys])
    return rows, cols, vals, keep


def make_case(name, coords, elems, dim, denom):
    rc = rational_coords(coords, denom)
    cf = np.array([[float(x) for x in row] for row in rc])
    elems = [list(map(int, e)) for e in elems]
    pairs, _ = edge_table(elems)
    x = X[:dim]
    R = sp.Rational
    # polynomial coefficients, degree <= 2 (exact under the reference's default quadrature for the Laplace / stiffness operators)
    kappa = 1 + R(1, 2) * x[0] + R(1, 3) * x[1] ** 2 + R(1, 5) * x[0] * x[-1] + (R(1, 7) * x[2] if dim == 3 else 0)
    f1 = 1 + 2 * x[0] - R(3, 4) * x[-1]                                   # degree 1: exact with P1 test functions at order 2
    f2 = f1 + R(1, 2) * x[0] * x[1] - R(1, 3) * x[-1] ** 2                # degree 2: exact with P2 test functions at order 4
    ns = 3 if dim == 2 else 6
    rng = np.random.default_rng(7)
    H = sp.zeros(ns, ns)
    for i in range(ns):
        for j in range(ns):      # unsymmetric on purpose: the operator must not symmetrise H
            c0, c1, c2 = (R(int(v), 8) for v in rng.integers(-8, 9, 3))
            H[i, j] = (4 if i == j else 0) + c0 + c1 * x[0] + c2 * x[-1] * x[0]
    out = dict(coords=cf, elems=np.array(elems, dtype=np.int64), edge_pairs=np.array(pairs, dtype=np.int64).reshape(-1, 2),
               dim=np.int64(dim))
    for nm, ex in (("kappa", kappa), ("f1", f1), ("f2", f2)):
        out[nm + "_exp"], out[nm + "_cf"] = poly_terms(ex, dim)
    hexp, hcf = [], []
    for i in range(ns):
        for j in range(ns):
            e_, c_ = poly_terms(H[i, j], dim)
            hexp.append(e_)
            hcf.append(c_)
    out["H_exp"] = np.concatenate(hexp)
    out["H_cf"] = np.concatenate(hcf)
    out["H_len"] = np.array([len(c_) for c_ in hcf], dtype=np.int64)
    for degree in (1, 2):
        tag = "P%d" % degree
        n, lap = assemble(rc, elems, dim, degree, "laplace", kappa)
        out[tag + "_ndof"] = np.int64(n)
        for nm, d in (("laplace", lap), ("mass", assemble(rc, elems, dim, degree, "mass", sp.Integer(1) * R(3, 2))[1])):
            r_, c_, v_, k_ = pack_matrix(d)
            out["%s_%s_rows" % (tag, nm)], out["%s_%s_cols" % (tag, nm)], out["%s_%s_vals" % (tag, nm)] = r_, c_, v_
        _, src = assemble(rc, elems, dim, degree, "source", f1 if degree == 1 else f2)
        rhs = np.zeros(n)
        for k, v in src.items():
            rhs[k] = to_float(v)
        out[tag + "_source"] = rhs
        # P2 tetrahedral elasticity is exact-integrated on the smaller meshes only (30 x 30 blocks of 6 x 6 polynomial tangents)
        if not (dim == 3 and degree == 2 and len(elems) > 12):
            _, st = assemble(rc, elems, dim, degree, "stiffness", H)
            r_, c_, v_, k_ = pack_matrix(st)
            out[tag + "_stiffness_rows"], out[tag + "_stiffness_cols"], out[tag + "_stiffness_vals"] = r_, c_, v_
        print(name, tag, "ndof", n, "done", flush=True)
    np.savez_compressed(os.path.join(HERE, "exact_%s.npz" % name), **out)


def main():
    from importlib import import_module
    mg = import_module("adfem_jl_b200").meshgen       # mesh GENERATORS only (numpy index arithmetic, no FEM)
    # the reference's own test mesh: deps/MFEM/FemLaplace1/ftest.jl:7,18  Mesh(8, 8, 1/8), P1 and P2
    c, e = mg.tri_grid(8, 8, 1 / 8)
    make_case("tri_grid8", c, e, 2, 8)
    # a distorted, renumbered triangulation (both orientations occur, so the orientation fix is exercised)
    c, e = mg.jitter_unstructured(4, 3, 0.25, seed=5)
    make_case("tri_jitter", c, e, 2, 256)
    # Mesh3(2, 2, 2, 1/2) (src/MFEM3/MFEM.jl:124-185): both cube splittings
    c, e = mg.tet_grid(2, 2, 2, 0.5)
    make_case("tet_grid2", c, e, 3, 2)
    # one distorted cube of 5 tetrahedra with a mirrored element
    c, e = mg.tet_grid(1, 1, 1, 1.0)
    rng = np.random.default_rng(3)
    c = c + rng.integers(-20, 21, c.shape) / 128.0
    e = e.copy()
    e[2] = e[2][[1, 0, 2, 3]]
    make_case("tet_jitter", c, e, 3, 128)


if __name__ == "__main__":
    main()
