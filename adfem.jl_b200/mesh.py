"""Host-side mirror of the reference's `Mesh` / `Mesh3` constructors and mesh utilities.

Reference: src/MFEM/MFEM.jl:60-170 (`Mesh`), src/MFEM3/MFEM.jl:40-185 (`Mesh3`), src/MFEM/MCore.jl:385-450
(`bcedge`, `bcnode`), src/MFEM/MUtils.jl:32-50 (file constructor).  Same names and argument meaning; index
arrays are 0-based here (Julia's are 1-based).  The tables come from libadfem_cuda.so's mesh handle, which
replaces the reference's process-global `mmesh` / `mmesh3` singletons (deps/MFEM/Common.cpp:7).
"""
import ctypes as C

import numpy as np

from . import _lib, meshgen
from ._lib import check, lib

P1, P2 = "P1", "P2"


def _degree_of(degree):
    if degree in (P1, 1):
        return 1
    if degree in (P2, 2):
        return 2
    raise ValueError("Only degree = 1 or 2 is supported.")       # src/MFEM/MFEM.jl:67-69 (BDM1 is outside the path)


class _MeshBase:
    dim = 0

    def _create(self, coords, elems, order, degree, lorder, host_only):
        L = lib()
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        elems = np.ascontiguousarray(elems)
        assert coords.ndim == 2 and coords.shape[1] == self.dim
        assert elems.ndim == 2 and elems.shape[1] == self.dim + 1
        e32 = np.ascontiguousarray(elems, dtype=np.int32)
        h = C.c_void_p()
        check(L.adfem_mesh_create(C.byref(h), C.c_int(self.dim), coords.ctypes.data_as(_lib.c_dp), C.c_int(self.dim),
                                  C.c_int(coords.shape[0]), e32.ctypes.data_as(_lib.c_ip), C.c_int(e32.shape[0]),
                                  C.c_int(order), C.c_int(degree), C.c_int(lorder), C.c_int(_lib.HOST_ONLY if host_only else 0)))
        self.handle = h
        self.host_only = host_only
        info = lambda w: int(L.adfem_mesh_info(h, w))
        self.nodes = coords
        self.nnode, self.nelem, self.ndof = info(_lib.INFO_NV), info(_lib.INFO_NE), info(_lib.INFO_NDOF)
        self.elem_ndof, self.ngauss = info(_lib.INFO_ELEM_NDOF), info(_lib.INFO_NGAUSS)
        self.gauss_per_elem = info(_lib.INFO_GAUSS_PER_ELEM)
        self.elem_type = P1 if degree == 1 else P2
        self.degree = degree
        self._lazy = {}                                            # edges / conn / elems: ne x d int64 copies, fetched on first use only
        self._csr = {}

    def _fetch(self, name):
        if name not in self._lazy:
            L, h = lib(), self.handle
            if name == "edges":
                edges = np.zeros(2 * max(self.nedge, 1), dtype=np.int64)
                check(L.adfem_mesh_edges(h, edges.ctypes.data_as(_lib.c_lp)))
                self._lazy[name] = edges[:2 * self.nedge].reshape(2, self.nedge).T - 1
            elif name == "conn":
                conn = np.zeros(self.nelem * self.elem_ndof, dtype=np.int64)
                check(L.adfem_mesh_connectivity(h, conn.ctypes.data_as(_lib.c_lp)))
                self._lazy[name] = conn.reshape(self.nelem, self.elem_ndof) - 1
            else:
                ev = np.zeros(self.nelem * (self.dim + 1), dtype=np.int64)
                check(L.adfem_mesh_element_to_vertices(h, ev.ctypes.data_as(_lib.c_lp)))
                self._lazy[name] = ev.reshape(self.dim + 1, self.nelem).T - 1     # post orientation fix, like src/MFEM/MFEM.jl:106
        return self._lazy[name]

    nedge = property(lambda self: int(lib().adfem_mesh_info(self.handle, _lib.INFO_NEDGES)))      # P1 meshes number their edges on first use
    edges = property(lambda self: self._fetch("edges"))            # nedge x 2, 0-based (src/MFEM/MFEM.jl:99-100)
    conn = property(lambda self: self._fetch("conn"))              # ne x d global dofs, 0-based
    elems = property(lambda self: self._fetch("elems"))            # ne x (dim+1) vertices after the orientation fix

    def __del__(self):
        h = getattr(self, "handle", None)
        if h is not None and _lib._lib is not None:
            _lib._lib.adfem_mesh_destroy(h)
            self.handle = None

    def set_option(self, key, value):
        check(lib().adfem_set_option(self.handle, key.encode(), C.c_longlong(int(value))))

    # --- mesh-static symbolic data ---------------------------------------------------------------
    def csr_pattern(self, ncomp=1):
        """(rowptr int64[n+1], colind int32[nnz]) of the scalar (ncomp=1) or elasticity (ncomp=dim) operator."""
        if ncomp not in self._csr:
            L = lib()
            nnz = L.adfem_csr_nnz(self.handle, C.c_int(ncomp))
            if nnz < 0:
                raise _lib.AdfemError(_lib.last_error())
            rowptr = np.zeros(ncomp * self.ndof + 1, dtype=np.int64)
            colind = np.zeros(nnz, dtype=np.int32)
            check(L.adfem_csr_pattern(self.handle, C.c_int(ncomp), rowptr.ctypes.data_as(_lib.c_lp), colind.ctypes.data_as(_lib.c_ip)))
            self._csr[ncomp] = (rowptr, colind)
        return self._csr[ncomp]

    def slot_to_nnz(self):
        out = np.zeros(self.nelem * self.elem_ndof ** 2, dtype=np.uint32)
        check(lib().adfem_slot_to_nnz(self.handle, out.ctypes.data_as(C.POINTER(C.c_uint))))
        return out

    def plan_array(self, which, ncomp, array_id, dtype):
        L = lib()
        n = L.adfem_plan_array(self.handle, C.c_int(which), C.c_int(ncomp), C.c_int(array_id), None)
        if n < 0:
            raise _lib.AdfemError(_lib.last_error())
        out = np.zeros(n, dtype=dtype)
        L.adfem_plan_array(self.handle, C.c_int(which), C.c_int(ncomp), C.c_int(array_id), out.ctypes.data_as(C.c_void_p))
        return out


class Mesh(_MeshBase):
    """`Mesh(coords, elems, order=-1, degree=1, lorder=-1)`, `Mesh(m, n, h; order, degree, lorder, version)` or
    `Mesh(filename)` — src/MFEM/MFEM.jl:60-170, src/MFEM/MUtils.jl:32-50."""
    dim = 2

    def __init__(self, *args, order=-1, degree=1, lorder=-1, version=1, host_only=False):
        if len(args) == 1 and isinstance(args[0], str):
            coords, elems = read_mesh_file(args[0])
            if order == -1:
                order = 2                      # file constructor defaults (quirk Q13): order=2, lorder=2
            if lorder == -1:
                lorder = 2
        elif len(args) >= 3 and np.isscalar(args[0]) and np.isscalar(args[1]):
            m, n, h = int(args[0]), int(args[1]), float(args[2])
            coords, elems = meshgen.tri_grid(m, n, h, version=version, dtype=np.int32 if (m + 1) * (n + 1) < 2 ** 31 else np.int64)
        else:
            coords, elems = args[0], args[1]
            if len(args) > 2:
                order = args[2]
            if len(args) > 3:
                degree = args[3]
            if len(args) > 4:
                lorder = args[4]
        self.lorder = 6 if lorder == -1 else lorder
        self._create(coords, elems, order, _degree_of(degree), lorder, host_only)


class Mesh3(_MeshBase):
    """`Mesh3(coords, elems, order=-1, degree=1)` or `Mesh3(m, n, l, h; order, degree)` — src/MFEM3/MFEM.jl:40-185."""
    dim = 3

    def __init__(self, *args, order=-1, degree=1, lorder=-1, host_only=False):
        if len(args) >= 4 and np.isscalar(args[0]):
            m, n, l, h = int(args[0]), int(args[1]), int(args[2]), float(args[3])
            coords, elems = meshgen.tet_grid(m, n, l, h, dtype=np.int32 if (m + 1) * (n + 1) * (l + 1) < 2 ** 31 else np.int64)
        else:
            coords, elems = args[0], args[1]
            if len(args) > 2:
                order = args[2]
            if len(args) > 3:
                degree = args[3]
        self.lorder = lorder
        self._create(coords, elems, order, _degree_of(degree), -1, host_only)


def read_mesh_file(filename):
    """`.npz` (nodes, elems) or ASCII `.stl` read the way meshio 4.2 feeds `Mesh(filename)`: one triangle per
    facet, bit-identical vertices merged in first-appearance order, z dropped (src/MFEM/MUtils.jl:39-49)."""
    if filename.endswith(".npz"):
        d = np.load(filename)
        return d["nodes"], d["elems"]
    pts, index, tris, cur = [], {}, [], []
    with open(filename) as fh:
        for line in fh:
            t = line.split()
            if len(t) == 4 and t[0] == "vertex":
                key = (float(t[1]), float(t[2]), float(t[3]))
                if key not in index:
                    index[key] = len(pts)
                    pts.append(key)
                cur.append(index[key])
                if len(cur) == 3:
                    tris.append(cur)
                    cur = []
    if not tris:
        raise ValueError("No triangles found in the mesh file.")      # src/MFEM/MUtils.jl:46-48
    return np.array(pts)[:, :2], np.array(tris, dtype=np.int64)


# ------------------------------------------------------------------------------------------------
def get_ngauss(mesh):
    """src/MFEM/MFEM.jl `get_ngauss` -> mfem_get_ngauss (deps/MFEM/API.cpp:17-19)."""
    return mesh.ngauss


def gauss_nodes(mesh):
    """ngauss x dim Gauss-point coordinates (deps/MFEM/API.cpp:21-24)."""
    out = np.zeros(mesh.dim * mesh.ngauss)
    check(lib().adfem_mesh_gauss(mesh.handle, out.ctypes.data_as(_lib.c_dp)))
    return out.reshape(mesh.dim, mesh.ngauss).T.copy()


def gauss_nodes_soa(mesh):
    """(dim, ngauss) array: row c = coordinate c of every Gauss point — the layout of mfem_get_gauss(x, y) itself, without the transposed copy
    that `gauss_nodes` makes for the Julia-shaped (ngauss, dim) result."""
    out = np.zeros(mesh.dim * mesh.ngauss)
    check(lib().adfem_mesh_gauss(mesh.handle, out.ctypes.data_as(_lib.c_dp)))
    return out.reshape(mesh.dim, mesh.ngauss)


def gauss_weights(mesh):
    out = np.zeros(mesh.ngauss)
    check(lib().adfem_mesh_gauss_weights(mesh.handle, out.ctypes.data_as(_lib.c_dp)))
    return out


def get_area(mesh):
    """Heron areas (2-D, deps/MFEM/Common.cpp:9-15) / tet volumes (3-D, mfem_get_volume3)."""
    out = np.zeros(mesh.nelem)
    check(lib().adfem_mesh_measure(mesh.handle, out.ctypes.data_as(_lib.c_dp)))
    return out


get_volume = get_area


def fem_nodes(mesh):
    """Coordinates of all dofs: vertices, then (P2) edge mid-points — src/MFEM/MCore.jl:47-60."""
    if mesh.elem_type == P1:
        return mesh.nodes.copy()
    mid = 0.5 * (mesh.nodes[mesh.edges[:, 0]] + mesh.nodes[mesh.edges[:, 1]])
    return np.concatenate([mesh.nodes, mid], 0)


def bcedge(mesh):
    """Boundary edges = edges seen by exactly one triangle (src/MFEM/MCore.jl:385-406); rows sorted (lo, hi)."""
    e = mesh.elems
    pairs = np.concatenate([e[:, [0, 1]], e[:, [2, 1]], e[:, [0, 2]]], 0)
    pairs = np.sort(pairs, 1)
    # rows as one integer key lo * nnode + hi: the ascending keys are the rows in (lo, hi) order, and a 1-D unique is several times faster than
    # the row-wise one (this is most of the setup of an element-block partition of a 16 M-triangle mesh)
    nn = np.int64(max(int(mesh.nnode), 1))
    key, cnt = np.unique(pairs[:, 0].astype(np.int64) * nn + pairs[:, 1], return_counts=True)
    key = key[cnt % 2 == 1]
    return np.stack([key // nn, key % nn], 1).astype(pairs.dtype, copy=False)


def get_edge_dof(edges, mesh):
    """Edge ids of (lo, hi) vertex pairs (src/MFEM/MUtils.jl `get_edge_dof`)."""
    key = np.minimum(mesh.edges[:, 0], mesh.edges[:, 1]) * np.int64(mesh.nnode) + np.maximum(mesh.edges[:, 0], mesh.edges[:, 1])
    order = np.argsort(key)
    edges = np.atleast_2d(edges)
    q = np.minimum(edges[:, 0], edges[:, 1]) * np.int64(mesh.nnode) + np.maximum(edges[:, 0], edges[:, 1])
    pos = np.searchsorted(key[order], q)
    return order[pos]


def bcnode(mesh, by_dof=True):
    """All boundary dofs (src/MFEM/MCore.jl:439-450). The reference returns them unordered (quirk Q12); sorted here."""
    bd = bcedge(mesh)
    nodes = np.unique(bd.reshape(-1))
    if by_dof and mesh.elem_type == P2:
        return np.concatenate([nodes, np.sort(get_edge_dof(bd, mesh)) + mesh.nnode])
    return nodes
