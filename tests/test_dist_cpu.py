"""world_size-2/3 gloo tests (CPU) of the multi-GPU path: element-block partition, interface-row reduce
(forward) and replicate (adjoint).  The rank-local kernel results are injected from the oracle; the exchange
code is the same torch.distributed code that runs over NCCL on the GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import adfem_jl_b200 as A
from adfem_jl_b200 import dist as adist
from adfem_jl_b200 import meshgen


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, degree, ret):
    from oracle import oracle as O
    try:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        if case == "tri":
            coords, elems = meshgen.jitter_unstructured(10, 9, 0.1, seed=4, permute=False)
            OM = O.Mesh2D
        elif case == "tri_perm":
            coords, elems = meshgen.jitter_unstructured(8, 8, 0.1, seed=5, permute=True)    # blocks are scattered: many neighbours
            OM = O.Mesh2D
        else:
            coords, elems = meshgen.tet_grid(4, 4, 3, 0.25)
            OM = O.Mesh3D
        # global reference (oracle on the whole mesh)
        og = OM(coords, elems, degree=degree)
        rng = np.random.default_rng(0)
        kappa = rng.random(og.ngauss) + 0.5
        gi, gv = og.laplace_fwd(kappa)
        grp, gci, gref = O.canonical_csr(gi, gv, og.ndof)
        gdof = np.zeros((og.ndof, coords.shape[1]))
        gdof[:og.nnode] = coords
        if degree == 2:
            gdof[og.nnode:] = 0.5 * (coords[og.edges[:, 0]] + coords[og.edges[:, 1]])

        if case == "slab":
            raise AssertionError
        part, (e0, e1) = adist.partition_elements(coords, elems, rank, world, degree=degree, host_only=True)
        m = part.mesh
        ol = OM(m.nodes, np.asarray(elems[e0:e1]) * 0 + np.searchsorted(np.unique(elems[e0:e1].reshape(-1)), elems[e0:e1]), degree=degree)
        assert np.array_equal(ol.conn, m.conn)
        g = og.g
        li, lv = ol.laplace_fwd(kappa[e0 * g:e1 * g])                  # coefficient arrays split with the elements
        lrp, lci, lvals = O.canonical_csr(li, lv, ol.ndof)
        assert np.array_equal(lrp, part.rowptr) and np.array_equal(lci, part.colind)
        vals = torch.from_numpy(lvals.copy())
        part.reduce_interface(vals)
        # compare the rows I own, by dof POSITION (global edge numbering differs from the local one for P2)
        ldof = np.zeros((m.ndof, coords.shape[1]))
        ldof[:m.nnode] = m.nodes
        if degree == 2:
            ldof[m.nnode:] = 0.5 * (m.nodes[m.edges[:, 0]] + m.nodes[m.edges[:, 1]])
        def keyify(p):
            return [tuple(np.round(x, 9)) for x in p]
        gpos = {k: i for i, k in enumerate(keyify(gdof))}
        l2g = np.array([gpos[k] for k in keyify(ldof)])
        gid_to_g = {int(part.gid[i]): int(l2g[i]) for i in range(m.ndof)}
        gr, gc, gvv = part.owned_rows_coo(vals)
        # ghost columns are dofs of other ranks: translate their gid through an all_gather'd dictionary
        objs = [None] * world
        dist.all_gather_object(objs, gid_to_g)
        full = {}
        for o in objs:
            full.update(o)
        import scipy.sparse as sp
        R = np.array([full[int(x)] for x in gr]); Cc = np.array([full[int(x)] for x in gc])
        mine = sp.coo_matrix((gvv, (R, Cc)), shape=(og.ndof, og.ndof)).tocsr()
        ref = sp.csr_matrix((gref, gci, grp), shape=(og.ndof, og.ndof))
        owned_g = l2g[part.owned]
        diff = abs(mine[owned_g] - ref[owned_g]).max()
        assert diff < 1e-12 * np.abs(gref).max(), diff
        # every global row is owned by exactly one rank
        cnt = torch.zeros(og.ndof, dtype=torch.int64)
        cnt[torch.from_numpy(owned_g)] += 1
        dist.all_reduce(cnt)
        assert (cnt == 1).all()
        # adjoint exchange: dK defined on owned entries (+ghost) must reach every contributing rank unchanged
        dKg = sp.csr_matrix((np.random.default_rng(1).standard_normal(len(gref)), gci, grp), shape=(og.ndof, og.ndof))
        rows_l = np.repeat(np.arange(m.ndof), np.diff(part.rowptr))
        want = np.asarray(dKg[l2g[rows_l], l2g[part.colind]]).reshape(-1)
        dv = torch.from_numpy(np.where(part.owned[rows_l], want, np.nan))
        dghost = torch.from_numpy(np.asarray(dKg[l2g[part.ghost_rows], [full[int(x)] for x in part.ghost_gcols]]).reshape(-1)) \
            if len(part.ghost_rows) else torch.zeros(0, dtype=torch.float64)
        part.replicate_interface(dv, dghost)
        assert np.array_equal(dv.numpy(), want)
        # ---- the component-blocked elasticity operator (configs 3 and 5): ncomp^2 values per scalar entry, layout of adfem_assemble_csr
        if degree == 1:
            nc = coords.shape[1]
            ns2 = 9 if nc == 2 else 36
            Hg = rng.random(og.ngauss * ns2) + 0.1
            sgi, sgv = og.stiffness_fwd(Hg)
            srp, sci, sref = O.canonical_csr(sgi, sgv, nc * og.ndof)
            sli, slv = ol.stiffness_fwd(Hg[ns2 * e0 * g:ns2 * e1 * g])
            lrp2, lci2, lsv = O.canonical_csr(sli, slv, nc * ol.ndof)
            sv = torch.from_numpy(lsv.copy())
            part.reduce_interface(sv, ncomp=nc)
            gr, gc, gvv, ca, cb = part.owned_rows_coo(sv, ncomp=nc)
            R = np.array([full[int(x)] for x in gr]) + ca * og.ndof
            Cc = np.array([full[int(x)] for x in gc]) + cb * og.ndof
            mine = sp.coo_matrix((gvv, (R, Cc)), shape=(nc * og.ndof, nc * og.ndof)).tocsr()
            ref = sp.csr_matrix((sref, sci, srp), shape=(nc * og.ndof, nc * og.ndof))
            own_rows = np.concatenate([owned_g + a * og.ndof for a in range(nc)])
            assert abs(mine[own_rows] - ref[own_rows]).max() < 1e-12 * np.abs(sref).max()
            # adjoint direction: dK of the global block matrix reaches every local entry
            dKb = sp.csr_matrix((np.random.default_rng(2).standard_normal(len(sref)), sci, srp), shape=ref.shape)
            lrows = np.repeat(np.arange(nc * m.ndof), np.diff(lrp2))
            la, lr = lrows // m.ndof, lrows % m.ndof
            lb, lc = lci2 // m.ndof, lci2 % m.ndof
            want = np.asarray(dKb[l2g[lr] + la * og.ndof, l2g[lc] + lb * og.ndof]).reshape(-1)
            dv2 = torch.from_numpy(np.where(part.owned[lr], want, np.nan))
            if len(part.ghost_rows):
                gR, gC = l2g[part.ghost_rows], np.array([full[int(x)] for x in part.ghost_gcols])
                dgh = np.stack([np.asarray(dKb[gR + a * og.ndof, gC + b * og.ndof]).reshape(-1) for a in range(nc) for b in range(nc)], 1).reshape(-1)
            else:
                dgh = np.zeros(0)
            part.replicate_interface(dv2, torch.from_numpy(dgh), ncomp=nc)
            assert np.array_equal(dv2.numpy(), want)
        # ---- dof-vector exchange ("source term: same exchange on a vector"), scalar and component-blocked
        f = rng.standard_normal(og.ngauss)
        gsrc = og.source_fwd(f)
        lsrc = torch.from_numpy(ol.source_fwd(f[e0 * g:e1 * g]))
        part.reduce_interface_vector(lsrc)
        assert np.abs(lsrc.numpy()[part.owned] - gsrc[owned_g]).max() < 1e-13 * max(1.0, np.abs(gsrc).max())
        part.replicate_interface_vector(lsrc)                                   # now every local copy holds the global value
        assert np.abs(lsrc.numpy() - gsrc[l2g]).max() < 1e-13 * max(1.0, np.abs(gsrc).max())
        if case != "tet":
            # matrix-free Laplace term (gather u -> scatter): u replicated from the owners, partial sums reduced to the owners
            ug = rng.standard_normal(og.ndof)
            ul = torch.from_numpy(np.where(part.owned, ug[l2g], np.nan))
            part.replicate_interface_vector(ul)
            assert np.array_equal(ul.numpy(), ug[l2g])
            term = torch.from_numpy(ol.laplace_term_fwd(kappa[e0 * g:e1 * g], ul.numpy()))
            part.reduce_interface_vector(term)
            gterm = og.laplace_term_fwd(kappa, ug)
            assert np.abs(term.numpy()[part.owned] - gterm[owned_g]).max() < 1e-12 * np.abs(gterm).max()
            # two components (strain-energy term): vector of 2*ndof entries, dof + c*ndof
            sig = rng.standard_normal(3 * og.ngauss)
            ge_ = og.strain_energy_fwd(sig)
            le_ = torch.from_numpy(ol.strain_energy_fwd(sig[3 * e0 * g:3 * e1 * g]))
            part.reduce_interface_vector(le_, ncomp=2)
            own2 = np.concatenate([part.owned, part.owned])
            gsel = np.concatenate([owned_g, owned_g + og.ndof])
            assert np.abs(le_.numpy()[own2] - ge_[gsel]).max() < 1e-12 * np.abs(ge_).max()
            part.replicate_interface_vector(le_, ncomp=2)
            assert np.abs(le_.numpy() - ge_[np.concatenate([l2g, l2g + og.ndof])]).max() < 1e-12 * np.abs(ge_).max()
        ret[rank] = "ok"
    except Exception as ex:  # surface the failure to the parent
        import traceback
        ret[rank] = traceback.format_exc()
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


@pytest.mark.parametrize("case,degree,world", [("tri", 1, 2), ("tri", 2, 2), ("tri_perm", 1, 3), ("tet", 1, 2), ("tet", 2, 2)])
def test_partition_and_interface_exchange(case, degree, world):
    from oracle import oracle as O
    O.build()
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, case, degree, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == "ok", ret.get(r)


def test_structured_slab_matches_global_numbering():
    """structured_slab() never forms the global mesh; its local slab must be the element block of Mesh(m, n, h)."""
    m, n, h, world = 6, 8, 0.5, 4
    gc, ge = meshgen.tri_grid(m, n, h)
    for rank in range(world):
        nl = n // world
        c, e = meshgen.tri_grid(m, nl, h)
        c[:, 1] += rank * nl * h
        gv = np.arange(c.shape[0]) + rank * nl * (m + 1)
        blk = ge[2 * m * nl * rank:2 * m * nl * (rank + 1)]
        assert np.array_equal(gv[e], blk) and np.allclose(gc[gv], c)


def test_structured_slab3_is_a_block_of_the_global_tet_grid():
    """structured_slab3(): the z-slab of rank r is the set of tetrahedra of Mesh3(n, n, l_total, h) whose cubes lie in its layers, with the
    same vertex coordinates and the same splitting (parity) — the local element order is the slab's own tet_grid order."""
    n, l_total, h, world = 3, 8, 0.5, 4
    gc, ge = meshgen.tet_grid(n, n, l_total, h)
    gset = {tuple(sorted(t)) for t in ge.tolist()}
    seen = set()
    for rank in range(world):
        l = l_total // world
        c, e = meshgen.tet_grid(n, n, l, h)
        c[:, 2] += rank * l * h
        gv = np.arange(c.shape[0]) + rank * l * (n + 1) * (n + 1)
        assert np.allclose(gc[gv], c)
        mine = {tuple(sorted(t)) for t in gv[e].tolist()}
        assert mine <= gset and not (mine & seen)
        seen |= mine
    assert seen == gset

