"""CPU tests of the product's host logic (no GPU): mesh tables, the symbolic CSR pattern and a replay of the
forward / adjoint tile plans against the oracle.  The replay executes exactly the index arithmetic the CUDA
kernels k_tile_fwd / k_tile_adj perform, with the oracle supplying the per-element local matrices."""
import numpy as np
import pytest

import adfem_jl_b200 as A
from adfem_jl_b200 import meshgen


def _meshes():
    out = {}
    out["tri_struct"] = (2, *meshgen.tri_grid(7, 5, 0.25))
    out["tri_unstruct"] = (2, *meshgen.jitter_unstructured(9, 8, 0.1, seed=3))
    out["tet"] = (3, *meshgen.tet_grid(3, 3, 2, 0.5))
    return out


MESHES = _meshes()
CASES = [(name, deg) for name in MESHES for deg in (1, 2)]


def make(name, degree, oracle):
    dim, c, e = MESHES[name]
    if dim == 2:
        return A.Mesh(c, e, degree=degree, host_only=True), oracle.Mesh2D(c, e, degree=degree)
    return A.Mesh3(c, e, degree=degree, host_only=True), oracle.Mesh3D(c, e, degree=degree)


@pytest.mark.parametrize("name,degree", CASES)
def test_mesh_tables_match_oracle(oracle, name, degree):
    m, o = make(name, degree, oracle)
    assert (m.nnode, m.nelem, m.nedge, m.ndof, m.ngauss, m.elem_ndof) == (o.nnode, o.nelem, o.nedge, o.ndof, o.ngauss, o.elem_ndof)
    assert np.array_equal(m.elems, o.elems)          # orientation fix
    assert np.array_equal(m.edges, o.edges)          # first-appearance edge numbering
    assert np.array_equal(m.conn, o.conn)
    assert np.array_equal(A.gauss_nodes(m), o.gauss)
    assert np.allclose(A.gauss_weights(m), o.weights, rtol=1e-15, atol=0)
    assert np.allclose(A.get_area(m), o.area if m.dim == 2 else o.volume, rtol=1e-15, atol=0)


@pytest.mark.parametrize("name,degree", CASES)
def test_csr_pattern_bit_exact(oracle, name, degree):
    m, o = make(name, degree, oracle)
    ind, vv = o.laplace_fwd(np.ones(o.ngauss))
    rp, ci, _ = oracle.canonical_csr(ind, vv, o.ndof)
    rowptr, colind = m.csr_pattern(1)
    assert np.array_equal(rowptr, rp) and np.array_equal(colind, ci)
    # slot -> nnz map: the Gauss-point-0 block of every element lists its (row, col) pairs
    g, dd = o.g, o.elem_ndof ** 2
    s2n = m.slot_to_nnz().reshape(o.nelem, dd)
    blk = ind.reshape(o.nelem, g, dd, 2)[:, 0]
    rows_of_nnz = np.repeat(np.arange(o.ndof), np.diff(rp))
    assert np.array_equal(rows_of_nnz[s2n], blk[..., 0]) and np.array_equal(ci[s2n], blk[..., 1])
    # elasticity pattern (component-blocked dofs)
    if degree == 1 or m.dim == 2:
        ns = 3 if m.dim == 2 else 6
        ind_e, vv_e = o.stiffness_fwd(np.tile(np.eye(ns).reshape(-1), o.ngauss))
        rp_e, ci_e, _ = oracle.canonical_csr(ind_e, vv_e, m.dim * o.ndof)
        rowptr_e, colind_e = m.csr_pattern(m.dim)
        assert np.array_equal(rowptr_e, rp_e) and np.array_equal(colind_e, ci_e)


def a16(x):
    return (x + 15) & ~15


def parse_blob(head, body, dim, d, fwd, sym):
    """Decode one tile (head + body) exactly as the kernels do (layout: adfem.jl_b200/csrc/plan.h)."""
    hdr = np.frombuffer(head, dtype=np.int32, count=8)
    nrows, nel, nvt = (int(x) for x in hdr[:3])
    out = {"nrows": nrows, "nel": nel, "nvt": nvt}
    o, buf = 32, head

    def take(dtype, count):
        nonlocal o
        a = np.frombuffer(buf, dtype=dtype, count=count, offset=o)
        o += a16(count * np.dtype(dtype).itemsize)
        return a
    if fwd:
        ndst, nsrc, ncls, ent32, nnz = int(hdr[3]), int(hdr[4]), int(hdr[5]), int(hdr[6]) & 1, int(hdr[7])
        out.update(nnz=nnz, nsrc=nsrc)
        out["elems"] = take(np.int32, nel)
        assert o == len(head)
        o, buf = 0, body
        out["rstart"] = take(np.uint32, nrows).astype(np.int64)
        if not sym:
            out["rlen"] = take(np.uint16, nrows).astype(np.int64)
        out["tv"] = take(np.uint16, (dim + 1) * nel).reshape(dim + 1, nel)
        out["xy"] = take(np.float64, dim * nvt).reshape(nvt, dim)
        out["cls"] = take(np.int32, 4 * ncls).reshape(ncls, 4)
        e = take(np.uint32 if ent32 else np.uint16, ndst).astype(np.int64)
        out["dst_lr"], out["dst_j"] = (e & 0xffff, e >> 16) if ent32 else (e & 0xff, e >> 8)
        out["src"] = take(np.uint16, nsrc).astype(np.int64)
        assert ent32 or nrows <= 256
    else:
        nnz, flags = int(hdr[3]), int(hdr[4])
        out["nnz"] = nnz
        out["delta"] = take(np.uint32, nrows)
        out["roff"] = take(np.uint16, nrows + 1)
        out["rstart"] = ((out["delta"].astype(np.uint64) + out["roff"][:-1]) & 0xffffffff).astype(np.int64)
        out["lrow"] = take(np.uint16 if flags & 1 else np.uint8, nnz)
        assert (flags & 1) or nrows <= 256
        assert o == len(head)
        o, buf = 0, body
        out["elems"] = take(np.int32, nel)
        out["tv"] = take(np.uint16, (dim + 1) * nel).reshape(dim + 1, nel)
        out["xy"] = take(np.float64, dim * nvt).reshape(nvt, dim)
        out["td"] = take(np.uint16, d * nel).reshape(d, nel) if flags & 2 else out["tv"]
        W = (d + 3) // 4
        gpk = take(np.uint32, d * W * nel).reshape(d, W, nel)
        out["gpos"] = np.stack([(gpk[p, q // 4] >> (8 * (q % 4))) & 0xff for p in range(d) for q in range(d)])      # [dd, nel]
    assert o == len(buf)
    return out


def tiles_of(m, which, ncomp):
    ptr = m.plan_array(which, ncomp, 0, np.int64)
    blob = m.plan_array(which, ncomp, 1, np.uint8).tobytes()
    assert len(ptr) % 2 == 1
    assert (ptr % 16 == 0).all()                                              # TMA bulk copy alignment rules (address and size)
    nt = (len(ptr) - 1) // 2
    return [parse_blob(blob[ptr[2 * t]:ptr[2 * t + 1]], blob[ptr[2 * t + 1]:ptr[2 * t + 2]], m.dim, m.elem_ndof, which == 0, ncomp == 1)
            for t in range(nt)]


def _check_tile_geometry(m, T):
    """tv/xy must reproduce the (orientation-fixed) vertex coordinates of the tile's elements."""
    assert np.array_equal(T["xy"][T["tv"].T], m.nodes[m.elems[T["elems"]]])


def _fwd_replay(m, local, ncomp, n_out):
    """local: [nelem, Dt*Dt] pre-summed local matrices. Mirrors k_tile_fwd."""
    d = m.elem_ndof
    dd, Dt = d * d, ncomp * d
    rowptr, _ = m.csr_pattern(1)
    nnz_s = rowptr[-1]
    sym = {}
    i = 0
    for p in range(d):
        for q in range(p, d):
            sym[i] = (p, q)
            i += 1
    vals = np.full(n_out, np.nan)
    written = np.zeros(n_out, dtype=np.int32)
    ntiles = npaired = 0
    for T in tiles_of(m, 0, ncomp):
        ntiles += 1
        _check_tile_geometry(m, T)
        nel = T["nel"]
        loc = local[T["elems"]]                               # [nel, S]
        rlen = rowptr[1:][np.searchsorted(rowptr[:-1], T["rstart"])] - T["rstart"]
        if ncomp > 1:
            assert np.array_equal(T["rlen"], rlen)
        assert rlen.sum() == T["nnz"] and (np.diff(T["cls"][:, 0]) > 0).all()
        for key, n, so, do in T["cls"]:
            cnt, paired = key & 0xffff, key >> 16
            assert not (paired and ncomp > 1)
            for i in range(n):
                cs = T["src"][so + np.arange(cnt) * n + i]
                dests = [do + i] + ([do + n + i] if paired else [])
                npaired += paired
                if ncomp == 1:
                    v = 0.0
                    for c in cs:
                        s, le = divmod(int(c), nel)
                        p, q = sym[s]
                        assert abs(loc[le, p * d + q] - loc[le, q * d + p]) <= 1e-14 * abs(loc[le]).max()
                        v += loc[le, p * d + q]
                    for dk in dests:
                        lr, j = int(T["dst_lr"][dk]), int(T["dst_j"][dk])
                        assert j < rlen[lr]
                        vals[T["rstart"][lr] + j] = v
                        written[T["rstart"][lr] + j] += 1
                else:
                    lr, j = int(T["dst_lr"][do + i]), int(T["dst_j"][do + i])
                    rs, ln = int(T["rstart"][lr]), int(rlen[lr])
                    assert j < ln
                    le, pq = cs // dd, cs % dd
                    p, q = pq // d, pq % d
                    for a in range(ncomp):
                        for b in range(ncomp):
                            v = 0.0
                            for k in range(len(cs)):
                                v += loc[le[k], ((a * d + p[k]) * Dt + b * d + q[k])]
                            dest = ncomp * (a * nnz_s + rs) + b * ln + j
                            vals[dest] = v
                            written[dest] += 1
    assert (written == 1).all()                                # every CSR entry is produced exactly once
    assert ncomp > 1 or npaired > 0                            # symmetric pairs are really merged
    return vals, ntiles


def _close(a, b, rel=1e-12):
    scale = max(np.abs(b).max(), 1e-300)
    return np.all(np.abs(a - b) <= rel * np.maximum(np.maximum(np.abs(a), np.abs(b)), scale * 1e-3))


@pytest.mark.parametrize("R", [24, 300])
@pytest.mark.parametrize("name,degree", CASES)
def test_forward_plan_replay_scalar(oracle, name, degree, R):
    m, o = make(name, degree, oracle)
    m.set_option("rows_per_tile", R)                           # several tiles even on these small meshes
    rng = np.random.default_rng(0)
    kappa = rng.random(o.ngauss) + 0.5
    ind, vv = o.laplace_fwd(kappa)
    rp, ci, ref = oracle.canonical_csr(ind, vv, o.ndof)
    dd = o.elem_ndof ** 2
    local = vv.reshape(o.nelem, o.g, dd).sum(1)
    vals, ntiles = _fwd_replay(m, local, 1, len(ref))
    assert ntiles >= (2 if R == 24 else 1)
    assert _close(vals, ref)


@pytest.mark.parametrize("name,degree", [("tri_unstruct", 1), ("tri_struct", 2), ("tet", 1)])
def test_forward_plan_replay_elasticity(oracle, name, degree):
    m, o = make(name, degree, oracle)
    m.set_option("rows_per_tile", 16)
    rng = np.random.default_rng(1)
    ns = 3 if m.dim == 2 else 6
    H = rng.random(o.ngauss * ns * ns)
    ind, vv = o.stiffness_fwd(H)
    rp, ci, ref = oracle.canonical_csr(ind, vv, m.dim * o.ndof)
    Dt = m.dim * o.elem_ndof
    local = vv.reshape(o.nelem, o.g, Dt * Dt).sum(1)
    vals, _ = _fwd_replay(m, local, m.dim, len(ref))
    assert _close(vals, ref)


def test_forward_plan_replay_elasticity_3d_with_coef_presum(oracle):
    """Option "coef_presum" changes the shared-memory slots per element of the 3-D P1 elasticity tiles (Gauss-summed blocks are staged): the
    plan is rebuilt and must still replay to the oracle's matrix."""
    m, o = make("tet", 1, oracle)
    rng = np.random.default_rng(2)
    H = rng.random(o.ngauss * 36)
    ind, vv = o.stiffness_fwd(H)
    rp, ci, ref = oracle.canonical_csr(ind, vv, 3 * o.ndof)
    local = vv.reshape(o.nelem, o.g, 144).sum(1)
    v0, n0 = _fwd_replay(m, local, 3, len(ref))
    m.set_option("coef_presum", 1)                             # clears the cached 3-D plan
    v1, n1 = _fwd_replay(m, local, 3, len(ref))
    assert _close(v0, ref) and _close(v1, ref) and n0 >= 1 and n1 >= 1


@pytest.mark.parametrize("name,degree", CASES)
def test_adjoint_plan_replay(oracle, name, degree):
    """Mirrors k_tile_adj: staged row segments + gidx must deliver dvals[slot_nnz[e,p,q]] to every owned element."""
    m, o = make(name, degree, oracle)
    m.set_option("elems_per_tile", 20)
    d = o.elem_ndof
    dd = d * d
    rowptr, _ = m.csr_pattern(1)
    s2n = m.slot_to_nnz().reshape(o.nelem, dd)
    dvals = np.random.default_rng(2).standard_normal(rowptr[-1])
    seen = np.zeros(o.nelem, dtype=np.int32)
    for T in tiles_of(m, 1, 1):
        _check_tile_geometry(m, T)
        staged = np.concatenate([dvals[rs:rs + ln] for rs, ln in zip(T["rstart"], np.diff(T["roff"].astype(int)))])
        assert len(staged) == T["nnz"]
        lr = T["lrow"].astype(int)
        assert np.array_equal(staged, dvals[(np.arange(T["nnz"], dtype=np.uint64) + T["delta"][lr]) & 0xffffffff])   # the kernel's staging loop
        tel = T["elems"]
        # the kernel's gather: staged[roff[td_p] + gpos[p*d+q]]
        idx = T["roff"].astype(int)[T["td"].astype(int)][np.repeat(np.arange(d), d)] + T["gpos"].astype(int)      # [dd, nel]
        assert np.array_equal(staged[idx.T], dvals[s2n[tel]])
        if degree == 1:
            assert T["td"] is T["tv"]                          # P1: the staged rows are the tile vertices (no td section)
        seen[tel] += 1
    assert (seen == 1).all()


@pytest.mark.parametrize("dim,degree,ncomp", [(2, 2, 1), (2, 1, 2), (3, 1, 3), (3, 2, 1)])
def test_symbolic_phase_is_thread_count_independent(oracle, dim, degree, ncomp):
    """The threaded symbolic phase (adjacency claimed with atomic increments and sorted per row, per-thread column buffers, per-thread tile
    blobs, early stop of rejected tile sizes) on meshes large enough to be split over the host threads (>= 4096 rows / tiles of 65536+
    elements), randomly renumbered: pattern, slot map and both tile plans are the same bytes for 1, 2 and 7 threads, and the pattern is the
    oracle's."""
    rng = np.random.default_rng(11)
    if dim == 2:
        c, e = meshgen.jitter_unstructured(190, 180, 1.0 / 190, seed=4, permute=True)
        mk, o = (lambda: A.Mesh(c, e, degree=degree, host_only=True)), oracle.Mesh2D(c, e, degree=degree)
    else:
        c, e = meshgen.tet_grid(26, 26, 22, 1.0 / 24)
        c = c + rng.uniform(-0.1 / 24, 0.1 / 24, c.shape)
        perm = rng.permutation(len(c))
        inv = np.empty_like(perm); inv[perm] = np.arange(len(c))
        c, e = np.ascontiguousarray(c[perm]), np.ascontiguousarray(inv[e][rng.permutation(len(e))]).astype(e.dtype)
        mk, o = (lambda: A.Mesh3(c, e, degree=degree, host_only=True)), oracle.Mesh3D(c, e, degree=degree)
    assert o.nelem >= 65536 and o.ndof >= 4096
    got = []
    for threads in (1, 2, 7):
        m = mk()
        m.set_option("host_threads", threads)
        rowptr, colind = m.csr_pattern(1)
        got.append([rowptr, colind, m.slot_to_nnz()] + [m.plan_array(which, ncomp, a, np.int64 if a == 0 else np.uint8) for which in (0, 1) for a in (0, 1)])
    for other in got[1:]:
        for a, b in zip(got[0], other):
            assert a.shape == b.shape and np.array_equal(a, b)
    # the component-blocked pattern handed to the caller (row blocks over the threads) is the documented expansion of the scalar one
    rowptr, colind = got[0][0], got[0][1]
    nd, nnz, lens = o.ndof, len(colind), np.diff(rowptr)
    rp_v, ci_v = m.csr_pattern(dim)
    within = np.arange(nnz) - np.repeat(rowptr[:-1], lens)
    want = np.empty(dim * dim * nnz, dtype=np.int64)
    for a in range(dim):
        assert np.array_equal(rp_v[a * nd:(a + 1) * nd], dim * (a * nnz + rowptr[:-1]))
        for b in range(dim):
            want[np.repeat(dim * (a * nnz + rowptr[:-1]) + b * lens, lens) + within] = colind + b * nd
    assert rp_v[-1] == dim * dim * nnz and np.array_equal(ci_v, want)
    ind, vv = o.laplace_fwd(np.ones(o.ngauss))
    rp, ci, _ = oracle.canonical_csr(ind, vv, o.ndof)
    assert np.array_equal(got[0][0], rp) and np.array_equal(got[0][1], ci)
    dd = o.elem_ndof ** 2
    blk = ind.reshape(o.nelem, o.g, dd, 2)[:, 0]
    rows_of_nnz = np.repeat(np.arange(o.ndof), np.diff(rp))
    s2n = got[0][2].reshape(o.nelem, dd)
    assert np.array_equal(rows_of_nnz[s2n], blk[..., 0]) and np.array_equal(ci[s2n], blk[..., 1])


@pytest.mark.parametrize("dim", [2, 3])
def test_threaded_edge_numbering_is_the_sequential_walk(oracle, dim):
    """P2 meshes of 32768+ elements number their edges by sorting (edge, appearance) records instead of walking the elements; ids, end points and
    the P2 connectivity must be those of the first-appearance walk (deps/MFEM/Common.cpp:70-74,133-140): against the oracle's walk and
    against the library's own sequential path (ADFEM_HOST_THREADS=1), on a randomly renumbered mesh."""
    import os
    rng = np.random.default_rng(dim)
    if dim == 2:
        c, e = meshgen.jitter_unstructured(140, 130, 1.0 / 140, seed=6, permute=True)
        mk, o = (lambda: A.Mesh(c, e, degree=2, host_only=True)), oracle.Mesh2D(c, e, degree=2)
        mk1 = lambda: A.Mesh(c, e, host_only=True)
    else:
        c, e = meshgen.tet_grid(19, 19, 20, 1.0 / 19)
        c = c + rng.uniform(-0.1 / 19, 0.1 / 19, c.shape)
        perm = rng.permutation(len(c))
        inv = np.empty_like(perm); inv[perm] = np.arange(len(c))
        c, e = np.ascontiguousarray(c[perm]), np.ascontiguousarray(inv[e][rng.permutation(len(e))]).astype(e.dtype)
        mk, o = (lambda: A.Mesh3(c, e, degree=2, host_only=True)), oracle.Mesh3D(c, e, degree=2)
        mk1 = lambda: A.Mesh3(c, e, host_only=True)
    assert o.nelem >= 32768
    old = os.environ.get("ADFEM_HOST_THREADS")
    try:
        got = []
        for threads in ("1", "4", "7"):
            os.environ["ADFEM_HOST_THREADS"] = threads
            m = mk()
            got.append((m.nedge, m.edges.copy(), m.conn.copy()))
            m1 = mk1()                       # P1: the edge list is numbered on first use, the same way
            assert m1.nedge == o.nedge and np.array_equal(m1.edges, o.edges)
    finally:
        if old is None:
            os.environ.pop("ADFEM_HOST_THREADS", None)
        else:
            os.environ["ADFEM_HOST_THREADS"] = old
    for nedge, edges, conn in got:
        assert nedge == o.nedge and np.array_equal(edges, o.edges) and np.array_equal(conn, o.conn)


def test_device_path_host_helpers(tmp_path):
    """Host helpers the library only reaches with a GPU (adfem_mesh_create's struct-of-arrays copies of the element tables), built for the host
    with g++ from the product's own source (tests/host_emul/host_check.cpp + csrc/host_mesh.cpp) and checked against numpy, below and above the
    size where they split the work over threads, for 1 and 5 threads."""
    import ctypes as C
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cxx = os.environ.get("CXX") or shutil.which("g++")
    if not cxx:
        pytest.skip("no C++ compiler")
    so = str(tmp_path / "libhost_check.so")
    csrc = os.path.join(root, "adfem.jl_b200", "csrc")
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-I", csrc, os.path.join(root, "tests", "host_emul", "host_check.cpp"),
                           os.path.join(csrc, "host_mesh.cpp"), "-o", so])
    L = C.CDLL(so)
    rng = np.random.default_rng(3)
    old = os.environ.get("ADFEM_HOST_THREADS")
    try:
        for threads in ("1", "5"):
            os.environ["ADFEM_HOST_THREADS"] = threads
            for ne, k in ((0, 3), (1, 4), (1000, 3), (70001, 6), (400003, 4), (150000, 10)):
                aos = rng.integers(0, 2 ** 31 - 1, size=(ne, k), dtype=np.int32)
                out = np.full(ne * k, -7, dtype=np.int32)
                assert L.check_soa_copy(aos.ctypes.data_as(C.c_void_p), C.c_longlong(ne), C.c_int(k), out.ctypes.data_as(C.c_void_p)) == 0
                assert np.array_equal(out.reshape(k, ne), aos.T)
    finally:
        if old is None:
            os.environ.pop("ADFEM_HOST_THREADS", None)
        else:
            os.environ["ADFEM_HOST_THREADS"] = old


@pytest.mark.parametrize("degree", [1, 2])
def test_high_valence_vertex(oracle, degree):
    """A fan of 3000 triangles around one vertex: rows far longer than the small-row fast paths of the symbolic phase (sorted adjacency rows,
    sorted column sets).  Pattern and slot map are the oracle's for 1 and 4 host threads; the tile plans refuse such a mesh with an error
    (a row of 3001 entries fits no tile; the COO-compatible kernels and the gather adjoint take it) — never with a crash."""
    K = 3000
    th = np.linspace(0, 2 * np.pi, K, endpoint=False)
    c = np.concatenate([[[0.0, 0.0]], np.stack([np.cos(th), np.sin(th)], 1)])
    e = np.stack([np.zeros(K, dtype=np.int64), 1 + np.arange(K), 1 + (np.arange(K) + 1) % K], 1)
    e = e[np.random.default_rng(0).permutation(K)]
    o = oracle.Mesh2D(c, e, degree=degree)
    ind, vv = o.laplace_fwd(np.ones(o.ngauss))
    rp, ci, _ = oracle.canonical_csr(ind, vv, o.ndof)
    dd = o.elem_ndof ** 2
    blk = ind.reshape(o.nelem, o.g, dd, 2)[:, 0]
    rows_of_nnz = np.repeat(np.arange(o.ndof), np.diff(rp))
    for threads in (1, 4):
        m = A.Mesh(c, e, degree=degree, host_only=True)
        m.set_option("host_threads", threads)
        rowptr, colind = m.csr_pattern(1)
        assert np.array_equal(rowptr, rp) and np.array_equal(colind, ci) and np.diff(rp).max() >= K
        s2n = m.slot_to_nnz().reshape(o.nelem, dd)
        assert np.array_equal(rows_of_nnz[s2n], blk[..., 0]) and np.array_equal(ci[s2n], blk[..., 1])
        for which in (0, 1):
            with pytest.raises(A._lib.AdfemError):
                m.plan_array(which, 1, 0, np.int64)


def test_degenerate_elements_do_not_break_the_symbolic_phase(oracle):
    """Elements that list a vertex twice (zero area: their values are meaningless, as in the reference) must not corrupt the symbolic phase:
    a CSR entry then receives several contributions from one (element, local dof) pair.  Pattern = the oracle's index set; slot map consistent;
    plans build and are the same bytes for 1 and 4 threads."""
    c, e = meshgen.jitter_unstructured(40, 30, 0.1, seed=5)
    e = e.copy()
    for at in (0, 17, len(e) // 2, len(e) - 1):
        e[at, 1] = e[at, 0]
    o = oracle.Mesh2D(c, e)
    ind, _ = o.laplace_fwd(np.ones(o.ngauss))
    rp, ci, _ = oracle.canonical_csr(ind, np.zeros(len(ind)), o.ndof)
    blk = ind.reshape(o.nelem, o.g, 9, 2)[:, 0]
    rows_of_nnz = np.repeat(np.arange(o.ndof), np.diff(rp))
    got = []
    for threads in (1, 4):
        m = A.Mesh(c, e, host_only=True)
        m.set_option("host_threads", threads)
        rowptr, colind = m.csr_pattern(1)
        assert np.array_equal(rowptr, rp) and np.array_equal(colind, ci)
        s2n = m.slot_to_nnz().reshape(o.nelem, 9)
        assert np.array_equal(rows_of_nnz[s2n], blk[..., 0]) and np.array_equal(ci[s2n], blk[..., 1])
        got.append([m.plan_array(w, nc, a, np.int64 if a == 0 else np.uint8) for nc in (1, 2) for w in (0, 1) for a in (0, 1)])
    for x, y in zip(*got):
        assert np.array_equal(x, y)


def test_plan_manifest_unchanged():
    """The symbolic phase is a byte-exact contract with the kernels (tile blobs are decoded on the device): digests of pattern, slot map and
    both tile plans over every element family / numbering / plan kind (scripts/host_plan_manifest.py) against the manifest committed when
    the kernels were last validated on B200 (tests/golden/host_plan_manifest.txt).  A deliberate change of the blob layout regenerates the
    manifest together with a GPU validation pass."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "host_plan_manifest.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    want = open(os.path.join(root, "tests", "golden", "host_plan_manifest.txt")).read().split("\n")
    got = out.stdout.split("\n")
    assert [l for l in got if l.strip()] == [l for l in want if l.strip()]


def test_product_has_no_cpu_compute_path():
    """A host-only handle (and any box without CUDA) must refuse to compute: no CPU fallback."""
    import ctypes as C
    m = A.Mesh(3, 3, 0.5, host_only=True)
    L = A._lib.lib()
    rc = L.adfem_assemble_csr(m.handle, 0, None, None, None)
    assert rc != 0 and "no CPU fallback" in A._lib.last_error()
    if L.adfem_device_count() == 0:
        with pytest.raises(A._lib.AdfemError):
            A.Mesh(3, 3, 0.5)
