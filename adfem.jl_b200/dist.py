"""Multi-GPU assembly: element-block partition + interface-row exchange (one process per GPU).

The reference has no parallelism at all; this is the new §8(e) design:

  * each rank holds a contiguous block of ELEMENTS and a local mesh over the dofs they touch (local ids keep
    the global order), so the coefficient arrays (`kappa`, `H`, `rho`, `f`: element-major, e*g+k) split with the
    elements at zero cost and the assembly / adjoint kernels run unchanged and rank-local;
  * a dof row touched by several ranks is OWNED by the lowest such rank.  Forward: a rank packs the CSR row
    segments of the rows it does not own and one `all_to_all_single` (NCCL over NVLink) delivers them to the
    owners, which add them into their rows — entries whose column the owner already has are summed in place,
    entries with a ghost column form a small off-process block (`ghost_*`, like PETSc's MPIAIJ off-diagonal part);
  * adjoint: the same lists run backwards (`replicate_interface`) so every rank sees d loss / d K for all entries
    its own elements contributed to; the element-level adjoint then needs no reduction.

Setup (numpy + one collective of index lists) is mesh-static.  The per-step exchange works on any
`torch.distributed` backend: NCCL on the GPUs, gloo in the CPU unit tests (which inject oracle values).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, meshgen
from .mesh import Mesh, Mesh3, bcedge, get_edge_dof


def make_nccl_comm(rank, world, group=None):
    """An `ncclComm_t` owned by libadfem_cuda (adfem_dist_comm_create): rank 0 draws the ncclUniqueId, torch.distributed carries its
    128 bytes to the other ranks (any backend), every rank then joins with ncclCommInitRank on its current CUDA device."""
    L = _lib.lib()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        _lib.check(L.adfem_dist_nccl_unique_id(buf))
    box = [bytes(buf)]
    if world > 1:
        dist.broadcast_object_list(box, src=0, group=group)
    comm = C.c_void_p()
    _lib.check(L.adfem_dist_comm_create(C.byref(comm), box[0], C.c_int(rank), C.c_int(world)))
    return comm


def _boundary_faces(elems):
    f = np.concatenate([elems[:, [0, 1, 2]], elems[:, [0, 1, 3]], elems[:, [0, 2, 3]], elems[:, [1, 2, 3]]], 0)
    f = np.sort(f, 1).astype(np.int64, copy=False)
    # rows as integer keys (a 1-D unique is several times faster than the row-wise one): first (a, b) -> its rank among the distinct pairs, then
    # rank * nv + c; ascending keys are the rows in (a, b, c) order
    nv = np.int64(int(f.max()) + 1 if f.size else 1)
    ab, rank_ab = np.unique(f[:, 0] * nv + f[:, 1], return_inverse=True)
    key, cnt = np.unique(rank_ab.reshape(-1).astype(np.int64) * nv + f[:, 2], return_counts=True)
    key = key[cnt == 1]
    k_ab = ab[key // nv]
    return np.stack([k_ab // nv, k_ab % nv, key % nv], 1).astype(elems.dtype, copy=False)


class Partition:
    """One rank's share of a distributed mesh.

    coords/elems: the rank's own elements in LOCAL vertex numbering; gvid[i] = global id of local vertex i
    (ascending); nv_global = number of global vertices.  `mesh_kwargs` go to Mesh/Mesh3 (degree, order, host_only)."""

    def __init__(self, coords, elems, gvid, nv_global, rank, world, group=None, device=None, interface_vertices=None, **mesh_kwargs):
        self.rank, self.world, self.group = rank, world, group
        dim = coords.shape[1]
        self.mesh = (Mesh if dim == 2 else Mesh3)(coords, elems, **mesh_kwargs)
        m = self.mesh
        self.device = device if device is not None else (torch.device("cpu") if m.host_only else torch.device("cuda", torch.cuda.current_device()))
        gvid = np.asarray(gvid, dtype=np.int64)
        # globally consistent dof ids: vertices keep their global id; a P2 edge dof is keyed by its global end points
        gid = np.empty(m.ndof, dtype=np.int64)
        gid[:m.nnode] = gvid
        if m.ndof > m.nnode:
            lo, hi = gvid[m.edges[:, 0]], gvid[m.edges[:, 1]]
            gid[m.nnode:] = nv_global + np.minimum(lo, hi) * np.int64(nv_global) + np.maximum(lo, hi)
        self.gid = gid
        self._gid_order = np.argsort(gid, kind="stable")
        self._gid_sorted = gid[self._gid_order]
        # dofs that can be shared: those on facets of the block boundary (`interface_vertices`: the caller knows the candidate
        # vertices, e.g. the two end planes of a structured slab — skips the facet sort, which dominates the setup of large P1 meshes)
        if interface_vertices is not None and m.ndof == m.nnode:
            cand = np.unique(np.asarray(interface_vertices, dtype=np.int64))
        elif dim == 2:
            bd = bcedge(m)
            cand = np.unique(bd.reshape(-1))
            if m.ndof > m.nnode:
                cand = np.concatenate([cand, np.unique(get_edge_dof(bd, m)) + m.nnode])
        else:
            bf = _boundary_faces(m.elems)
            cand = np.unique(bf.reshape(-1))
            if m.ndof > m.nnode:
                be = np.concatenate([bf[:, [0, 1]], bf[:, [0, 2]], bf[:, [1, 2]]], 0)
                cand = np.concatenate([cand, np.unique(get_edge_dof(be, m)) + m.nnode])
        cand_gid = np.sort(gid[cand])
        all_cand = self._all_gather(cand_gid)
        # owner of every local dof = lowest rank that has it
        owner = np.full(m.ndof, rank, dtype=np.int64)
        shared_with = {}
        for q in range(world):
            if q == rank or len(all_cand[q]) == 0:
                continue
            common = np.intersect1d(cand_gid, all_cand[q], assume_unique=True)
            if len(common) == 0:
                continue
            loc = self.local_of_gid(common)
            shared_with[q] = loc
            owner[loc] = np.minimum(owner[loc], q)
        self.owner = owner
        self.owned = owner == rank
        rowptr, colind = m.csr_pattern(1)
        self.rowptr, self.colind = rowptr, colind
        # ---- forward send lists: every entry of a row I do not own goes to its owner
        send_pos, send_keys = [], []
        for q in range(world):
            rows = np.flatnonzero(owner == q) if q != rank else np.zeros(0, dtype=np.int64)
            if len(rows):
                lens = rowptr[rows + 1] - rowptr[rows]
                pos = np.repeat(rowptr[rows] - np.cumsum(np.r_[0, lens[:-1]]), lens) + np.arange(lens.sum())
                keys = np.stack([gid[np.repeat(rows, lens)], gid[colind[pos]]], 1)
            else:
                pos, keys = np.zeros(0, dtype=np.int64), np.zeros((0, 2), dtype=np.int64)
            send_pos.append(pos)
            send_keys.append(keys.reshape(-1))
        self.send_counts = [len(p) for p in send_pos]
        self._send_pos_np = np.ascontiguousarray(np.concatenate(send_pos), dtype=np.int64)
        self._dist = None                                      # adfem_dist handle once use_library() was called (GPU runs)
        self.send_pos = torch.from_numpy(np.concatenate(send_pos)).to(self.device)
        recv_keys = self._all_to_all([k for k in send_keys])
        self.recv_counts = [len(k) // 2 for k in recv_keys]
        rk = np.concatenate(recv_keys).reshape(-1, 2) if sum(self.recv_counts) else np.zeros((0, 2), dtype=np.int64)
        # match received (global row, global col) against my local pattern
        lrow = self.local_of_gid(rk[:, 0])
        assert (lrow >= 0).all() and self.owned[lrow].all(), "received a row this rank does not own"
        lcol = self.local_of_gid(rk[:, 1], missing_ok=True)
        pos = np.full(len(rk), -1, dtype=np.int64)
        have = lcol >= 0
        if have.any():
            urows = np.unique(lrow[have])
            lens = rowptr[urows + 1] - rowptr[urows]
            ent = np.repeat(rowptr[urows] - np.cumsum(np.r_[0, lens[:-1]]), lens) + np.arange(lens.sum())
            key_have = np.repeat(urows, lens) * np.int64(m.ndof) + colind[ent]          # ascending (rows asc, cols asc)
            q_keys = lrow[have] * np.int64(m.ndof) + lcol[have]
            at = np.searchsorted(key_have, q_keys)
            at = np.minimum(at, len(key_have) - 1)
            hit = key_have[at] == q_keys
            ph = np.full(have.sum(), -1, dtype=np.int64)
            ph[hit] = ent[at[hit]]
            pos[have] = ph
        matched = pos >= 0
        self._recv_pos_np = np.ascontiguousarray(pos, dtype=np.int64)
        self.recv_match_idx = torch.from_numpy(np.flatnonzero(matched)).to(self.device)
        self.recv_match_pos = torch.from_numpy(pos[matched]).to(self.device)
        self.ghost_idx = torch.from_numpy(np.flatnonzero(~matched)).to(self.device)
        self.ghost_rows = lrow[~matched]                       # local row ids (owned)
        self.ghost_gcols = rk[~matched, 1]                     # global column ids
        self.ghost_vals = torch.zeros(int((~matched).sum()), dtype=torch.float64, device=self.device)
        self.interface_bytes = 8 * (sum(self.send_counts) + sum(self.recv_counts))
        # ---- dof-vector exchange lists (source term, Laplace / strain-energy terms, dof fields): the dofs I hold but do not own go to their
        # owner; the owner's list towards rank q is the dofs it shares with q and owns.  Both sides order by global id.
        by_gid = lambda loc: loc[np.argsort(gid[loc], kind="stable")] if len(loc) else np.zeros(0, dtype=np.int64)
        vsend = [by_gid(np.flatnonzero(owner == q)) if q != rank else np.zeros(0, dtype=np.int64) for q in range(world)]
        vrecv = [by_gid(shared_with[q][owner[shared_with[q]] == rank]) if q in shared_with else np.zeros(0, dtype=np.int64) for q in range(world)]
        self.vsend_counts, self.vrecv_counts = [len(a) for a in vsend], [len(a) for a in vrecv]
        their = self._all_to_all([gid[a] for a in vsend])                      # what the others will send me, as global ids
        for q in range(world):
            assert np.array_equal(their[q], gid[vrecv[q]]), "dof-vector exchange lists of two ranks disagree"
        self.vsend_idx = torch.from_numpy(np.concatenate(vsend)).to(self.device)
        self.vrecv_idx = [torch.from_numpy(a).to(self.device) for a in vrecv]

    # ---- exchange inside the library (include/adfem_cuda.h group 3) ------------------------------------------
    def use_library(self, comm=None, max_ncomp=None):
        """Route reduce_interface / replicate_interface through libadfem_cuda's own pack / ncclSend+ncclRecv / unpack path
        (csrc/dist.cu, deterministic owner sums).  `comm`: an ncclComm_t as c_void_p, default one made by make_nccl_comm()."""
        if self.world == 1 or self.mesh.host_only:
            return self
        L = _lib.lib()
        self._comm = comm if comm is not None else make_nccl_comm(self.rank, self.world, self.group)
        sc = np.asarray(self.send_counts, dtype=np.int64)
        rc = np.asarray(self.recv_counts, dtype=np.int64)
        h = C.c_void_p()
        _lib.check(L.adfem_dist_create(C.byref(h), self.mesh.handle, self._comm, C.c_int(self.rank), C.c_int(self.world),
                                       C.c_int(max_ncomp or self.mesh.dim), sc.ctypes.data_as(_lib.c_lp), self._send_pos_np.ctypes.data_as(_lib.c_lp),
                                       rc.ctypes.data_as(_lib.c_lp), self._recv_pos_np.ctypes.data_as(_lib.c_lp)))
        self._dist = h
        self._ghost_bufs = {}
        return self

    def _ghost_buf(self, ncomp, dtype, device):
        if ncomp not in self._ghost_bufs:
            self._ghost_bufs[ncomp] = torch.zeros(len(self.ghost_idx) * ncomp * ncomp, dtype=dtype, device=device)
        return self._ghost_bufs[ncomp]

    def __del__(self):
        if getattr(self, "_dist", None) is not None and _lib._lib is not None:
            _lib._lib.adfem_dist_destroy(self._dist)
            self._dist = None

    # ---- helpers ---------------------------------------------------------------------------------
    def local_of_gid(self, g, missing_ok=False):
        at = np.searchsorted(self._gid_sorted, g)
        at = np.minimum(at, len(self._gid_sorted) - 1)
        ok = self._gid_sorted[at] == g
        out = np.where(ok, self._gid_order[at], -1)
        if not missing_ok:
            assert ok.all()
        return out

    def _all_gather(self, arr):
        """all_gather of variable-length int64 arrays."""
        if self.world == 1:
            return [arr]
        n = torch.tensor([len(arr)], dtype=torch.int64, device=self.device)
        ns = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(ns, n, group=self.group)
        mx = max(int(x.item()) for x in ns)
        buf = torch.zeros(max(mx, 1), dtype=torch.int64, device=self.device)
        buf[:len(arr)] = torch.from_numpy(arr).to(self.device)
        bufs = [torch.zeros_like(buf) for _ in range(self.world)]
        dist.all_gather(bufs, buf, group=self.group)
        return [b[:int(k.item())].cpu().numpy() for b, k in zip(bufs, ns)]

    def _all_to_all(self, arrs):
        """all_to_all of variable-length int64 arrays (setup only)."""
        if self.world == 1:
            return [arrs[0]]
        cnt = torch.tensor([len(a) for a in arrs], dtype=torch.int64, device=self.device)
        rc = torch.zeros_like(cnt)
        dist.all_to_all_single(rc, cnt, group=self.group)
        rcl = [int(x) for x in rc.cpu()]
        send = torch.from_numpy(np.concatenate(arrs) if sum(len(a) for a in arrs) else np.zeros(0, dtype=np.int64)).to(self.device)
        recv = torch.zeros(sum(rcl), dtype=torch.int64, device=self.device)
        dist.all_to_all_single(recv, send, output_split_sizes=rcl, input_split_sizes=[len(a) for a in arrs], group=self.group)
        out, o = [], 0
        r = recv.cpu().numpy()
        for c in rcl:
            out.append(r[o:o + c])
            o += c
        return out

    # ---- per-step exchanges ------------------------------------------------------------------------
    def _block_lists(self, ncomp):
        """Index lists of the interface exchange for the ncomp x ncomp block operator (elasticity): scalar entry `pos` of row r (length len,
        start rs, j = pos - rs) holds its ncomp^2 values at ncomp*(a*nnz + rs) + b*len + j (the layout of adfem_assemble_csr, = the
        canonical CSR of the component-blocked matrix).  The exchange moves the ncomp^2 values of an entry together (entry-major)."""
        cache = self.__dict__.setdefault("_block_cache", {})
        if ncomp not in cache:
            nc2, nnz = ncomp * ncomp, int(self.rowptr[-1])

            def expand(pos_t):
                pos = pos_t.cpu().numpy()
                rows = np.searchsorted(self.rowptr, pos, side="right") - 1
                rs, ln = self.rowptr[rows], self.rowptr[rows + 1] - self.rowptr[rows]
                out = np.empty((len(pos), ncomp, ncomp), dtype=np.int64)
                for a in range(ncomp):
                    for b in range(ncomp):
                        out[:, a, b] = ncomp * (a * nnz + rs) + b * ln + (pos - rs)
                return torch.from_numpy(out.reshape(-1)).to(self.device)

            def widen(idx_t):      # indices into the exchanged buffer: entry k -> k*nc2 .. k*nc2 + nc2 - 1
                return (idx_t.reshape(-1, 1) * nc2 + torch.arange(nc2, device=idx_t.device).reshape(1, -1)).reshape(-1)

            cache[ncomp] = dict(send_pos=expand(self.send_pos), match_pos=expand(self.recv_match_pos), match_idx=widen(self.recv_match_idx),
                                ghost_idx=widen(self.ghost_idx), send_counts=[c * nc2 for c in self.send_counts],
                                recv_counts=[c * nc2 for c in self.recv_counts])
        return cache[ncomp]

    def reduce_interface(self, vals, ncomp=1):
        """Forward: sum the partial interface rows into their owners (in place on `vals`, ghost-column part into
        `self.ghost_vals`).  Rows this rank does not own keep their partial sums (they are not part of its result).
        ncomp > 1: the component-blocked elasticity operator (vals in the layout of adfem_assemble_csr, ncomp^2 * nnz values;
        ghost_vals then holds ncomp^2 values per ghost entry, (a, b)-minor)."""
        if self.world == 1:
            return vals
        if self._dist is not None:
            self.ghost_vals = self._ghost_buf(ncomp, vals.dtype, vals.device)
            _lib.check(_lib.lib().adfem_dist_reduce(self._dist, C.c_int(ncomp), C.c_void_p(vals.data_ptr()), C.c_void_p(self.ghost_vals.data_ptr()),
                                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            return vals
        if ncomp == 1:
            L = dict(send_pos=self.send_pos, match_pos=self.recv_match_pos, match_idx=self.recv_match_idx, ghost_idx=self.ghost_idx,
                     send_counts=self.send_counts, recv_counts=self.recv_counts)
        else:
            L = self._block_lists(ncomp)
        send = vals.index_select(0, L["send_pos"])
        recv = torch.empty(sum(L["recv_counts"]), dtype=vals.dtype, device=vals.device)
        dist.all_to_all_single(recv, send, output_split_sizes=L["recv_counts"], input_split_sizes=L["send_counts"], group=self.group)
        vals.index_add_(0, L["match_pos"], recv.index_select(0, L["match_idx"]))
        self.ghost_vals = recv.index_select(0, L["ghost_idx"])
        return vals

    def replicate_interface(self, dvals, dghost=None, ncomp=1):
        """Adjoint: owners send d loss / d K of the interface entries back, so `dvals` becomes valid on every entry
        this rank's elements contribute to (rows it does not own included)."""
        if self.world == 1:
            return dvals
        if self._dist is not None:
            _lib.check(_lib.lib().adfem_dist_replicate(self._dist, C.c_int(ncomp), C.c_void_p(dvals.data_ptr()),
                                                       C.c_void_p(dghost.data_ptr()) if dghost is not None and dghost.numel() else None,
                                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            return dvals
        if ncomp == 1:
            L = dict(send_pos=self.send_pos, match_pos=self.recv_match_pos, match_idx=self.recv_match_idx, ghost_idx=self.ghost_idx,
                     send_counts=self.send_counts, recv_counts=self.recv_counts)
        else:
            L = self._block_lists(ncomp)
        back = torch.empty(sum(L["recv_counts"]), dtype=dvals.dtype, device=dvals.device)
        back.index_copy_(0, L["match_idx"], dvals.index_select(0, L["match_pos"]))
        if len(L["ghost_idx"]):
            back.index_copy_(0, L["ghost_idx"], dghost if dghost is not None else torch.zeros(len(L["ghost_idx"]), dtype=dvals.dtype, device=dvals.device))
        got = torch.empty(sum(L["send_counts"]), dtype=dvals.dtype, device=dvals.device)
        dist.all_to_all_single(got, back, output_split_sizes=L["send_counts"], input_split_sizes=L["recv_counts"], group=self.group)
        dvals.index_copy_(0, L["send_pos"], got)
        return dvals

    def _comp_index(self, idx, ncomp):
        """local dof indices -> indices into a component-blocked vector (dof + c*ndof), component-major"""
        if ncomp == 1:
            return idx
        return torch.cat([idx + c * self.mesh.ndof for c in range(ncomp)])

    def reduce_interface_vector(self, vec, ncomp=1):
        """Forward of the scatter-type operators (source term, Laplace term, strain-energy term): the partial sums of the dofs
        this rank holds but does not own go to their owners and are added there (in place; non-owned entries keep their partial sums and
        are not part of this rank's result).  `vec` has ncomp*ndof entries, component-blocked like the reference (dof + c*ndof).
        Contributions are added rank by rank in ascending order, so the result is reproducible."""
        if self.world == 1:
            return vec
        send = vec.index_select(0, self._comp_index(self.vsend_idx, ncomp)) if ncomp == 1 else \
            torch.cat([vec.index_select(0, self._comp_index(self.vsend_idx[o:o + c], ncomp)) for o, c in self._spans(self.vsend_counts)])
        recv = torch.empty(ncomp * sum(self.vrecv_counts), dtype=vec.dtype, device=vec.device)
        dist.all_to_all_single(recv, send, output_split_sizes=[ncomp * c for c in self.vrecv_counts],
                               input_split_sizes=[ncomp * c for c in self.vsend_counts], group=self.group)
        o = 0
        for q in range(self.world):
            c = ncomp * self.vrecv_counts[q]
            if c:
                vec.index_add_(0, self._comp_index(self.vrecv_idx[q], ncomp), recv[o:o + c])
            o += c
        return vec

    def replicate_interface_vector(self, vec, ncomp=1):
        """The other direction: every owner sends its values of the shared dofs to the ranks that hold copies (in place).  Used for dof
        fields that feed the gather-type operators (u at the Gauss points, strain, gradients) and for the upstream gradient of the
        scatter-type operators in the adjoint."""
        if self.world == 1:
            return vec
        back = torch.cat([vec.index_select(0, self._comp_index(self.vrecv_idx[q], ncomp)) for q in range(self.world)]) \
            if sum(self.vrecv_counts) else torch.empty(0, dtype=vec.dtype, device=vec.device)
        got = torch.empty(ncomp * sum(self.vsend_counts), dtype=vec.dtype, device=vec.device)
        dist.all_to_all_single(got, back, output_split_sizes=[ncomp * c for c in self.vsend_counts],
                               input_split_sizes=[ncomp * c for c in self.vrecv_counts], group=self.group)
        o = 0
        for off, c in self._spans(self.vsend_counts):
            if c:
                vec.index_copy_(0, self._comp_index(self.vsend_idx[off:off + c], ncomp), got[o:o + ncomp * c])
            o += ncomp * c
        return vec

    @staticmethod
    def _spans(counts):
        o = 0
        for c in counts:
            yield o, c
            o += c

    def owned_rows_coo(self, vals, ncomp=1):
        """(global row, global col, value) triplets of the rows this rank owns — for tests / hand-off to a solver.  ncomp > 1: also the
        component of the row and of the column, i.e. (gr, gc, gv, a, b)."""
        rows = np.repeat(np.arange(self.mesh.ndof), np.diff(self.rowptr))
        keep = self.owned[rows]
        v = vals.detach().cpu().numpy()
        if ncomp == 1:
            gr = np.concatenate([self.gid[rows[keep]], self.gid[self.ghost_rows]])
            gc = np.concatenate([self.gid[self.colind[keep]], self.ghost_gcols])
            gv = np.concatenate([v[keep], self.ghost_vals.cpu().numpy()])
            return gr, gc, gv
        nnz = int(self.rowptr[-1])
        pos = np.flatnonzero(keep)
        rs, ln = self.rowptr[rows[pos]], self.rowptr[rows[pos] + 1] - self.rowptr[rows[pos]]
        gh = self.ghost_vals.cpu().numpy().reshape(-1, ncomp, ncomp)
        out = [[], [], [], [], []]
        for a in range(ncomp):
            for b in range(ncomp):
                at = ncomp * (a * nnz + rs) + b * ln + (pos - rs)
                out[0] += [self.gid[rows[pos]], self.gid[self.ghost_rows]]
                out[1] += [self.gid[self.colind[pos]], self.ghost_gcols]
                out[2] += [v[at], gh[:, a, b]]
                out[3] += [np.full(len(pos) + len(self.ghost_rows), a)]
                out[4] += [np.full(len(pos) + len(self.ghost_rows), b)]
        return tuple(np.concatenate(o) for o in out)


def partition_elements(coords, elems, rank, world, **kw):
    """Element-block partition of a global mesh given as arrays (setup helper for moderately sized meshes)."""
    ne = elems.shape[0]
    e0, e1 = ne * rank // world, ne * (rank + 1) // world
    own = np.asarray(elems[e0:e1])
    gv = np.unique(own.reshape(-1))
    loc = np.searchsorted(gv, own)
    return Partition(np.asarray(coords)[gv], loc, gv, coords.shape[0], rank, world, **kw), (e0, e1)


def structured_slab(m, n_total, h, rank, world, **kw):
    """Rank `rank`'s row slab of Mesh(m, n_total, h) (src/MFEM/MFEM.jl:134-170) without ever forming the global mesh:
    element blocks of the row-major cell numbering are slabs of n_total/world cell rows."""
    assert n_total % world == 0
    nl = n_total // world
    j0 = rank * nl
    coords, elems = meshgen.tri_grid(m, nl, h)
    coords[:, 1] += j0 * h
    gvid = np.arange(coords.shape[0], dtype=np.int64) + j0 * (m + 1)
    iface = np.concatenate([np.arange(m + 1), np.arange(m + 1) + nl * (m + 1)])          # first and last node row of the slab
    return Partition(coords, elems, gvid, (m + 1) * (n_total + 1), rank, world, interface_vertices=iface, **kw)


def structured_slab3(n, l_total, h, rank, world, **kw):
    """Rank `rank`'s z-slab of the tetrahedral grid Mesh3(n, n, l_total, h) (src/MFEM3/MFEM.jl:124-185: 5 tets per cube, the two splittings
    alternate with the parity of i+j+k) without forming the global mesh: l_total/world layers of cubes per rank.  The layer count per rank
    must be even so that every slab starts on the same parity and is tet_grid(n, n, l, h) shifted in z; the vertex numbering is k-major,
    so a slab's vertices are a contiguous range of global ids."""
    assert l_total % world == 0
    l = l_total // world
    assert world == 1 or l % 2 == 0, "layers per rank must be even (parity-alternating tetrahedral splitting)"
    coords, elems = meshgen.tet_grid(n, n, l, h)
    coords[:, 2] += rank * l * h
    gvid = np.arange(coords.shape[0], dtype=np.int64) + rank * l * (n + 1) * (n + 1)
    plane = (n + 1) * (n + 1)
    iface = np.concatenate([np.arange(plane), np.arange(plane) + l * plane])          # bottom and top node planes of the slab
    return Partition(coords, elems, gvid, (n + 1) * (n + 1) * (l_total + 1), rank, world, interface_vertices=iface, **kw)

