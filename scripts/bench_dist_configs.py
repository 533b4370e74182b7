#!/usr/bin/env python
"""Multi-GPU timings of BASELINE.json configs 3-5 (one process per GPU, launched with torchrun like bench.py): every rank assembles its
element block with the unchanged kernels, interface rows are summed at their owners with one all_to_all (NCCL over NVLink) and the adjoint
runs the same lists backwards (adfem.jl_b200/dist.py).  Weak scaling: the per-GPU mesh is fixed, the global mesh grows with the ranks.

  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_dist_configs.py --cases 5 --steps 10
      config 5 at N = 8: Mesh3(215, 215, 26 * 8, h), 48 M tetrahedra in total

  --dry-run  CPU only (gloo, host-only meshes): the partition setup and both exchanges run on random values, no kernel is launched —
             this is how the host logic of this script is checked where there is no GPU (tests/test_bench_cli.py).
One JSON line per case from rank 0.  These are not bench.py lines (bench.py measures config 2).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import adfem_jl_b200 as A
from adfem_jl_b200 import _lib, meshgen
from adfem_jl_b200 import dist as adist


morton_order = meshgen.morton_element_order


def build_case(case, scale, rank, world, host_only):
    kw = dict(host_only=host_only)
    if case == "3":
        m, nl = max(4, int(4096 * scale)), max(2, int(2048 * scale))
        part = adist.structured_slab(m, nl * world, 1.0 / m, rank, world, **kw)
        return part, 2, 9, "config 3: P1 elasticity, row slab of Mesh(%d, %d, h) per GPU" % (m, nl)
    if case in ("4l", "4m"):
        n = max(4, int(1000 * scale))                          # the global mesh grows with the ranks: n x (n * world) cells
        coords, elems = meshgen.jitter_unstructured(n, n * world, 1.0 / n, seed=2)
        elems = elems[morton_order(coords, elems)]             # the random numbering stays; blocks are made compact first
        part, _ = adist.partition_elements(coords, elems, rank, world, degree=2, **kw)
        return part, 0 if case == "4l" else 1, 1, "config 4: P2 %s, Morton element blocks of a jittered, randomly renumbered %d x %d grid" % (
            "Laplace" if case == "4l" else "mass", n, n * world)
    if case == "5g":
        # STRONG scaling: the global mesh Mesh3(160, 160, 160, h) (20.5 M tetrahedra) is fixed, every rank takes 160 / world cube layers
        n = max(2, int(160 * scale)) // (2 * world) * (2 * world)
        part = adist.structured_slab3(n, n, 1.0 / n, rank, world, **kw)
        return part, 2, 36, "config 5, strong scaling: P1 tetrahedral elasticity, z-slabs of the fixed global Mesh3(%d, %d, %d, h)" % (n, n, n)
    if case == "5":
        n = max(2, int(215 * scale))
        l = max(2, int(26 * scale)) // 2 * 2
        part = adist.structured_slab3(n, l * world, 1.0 / n, rank, world, **kw)
        return part, 2, 36, "config 5: P1 tetrahedral elasticity, z-slab of Mesh3(%d, %d, %d, h) per GPU" % (n, n, l)
    raise SystemExit("unknown case " + case)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="3,4l,5")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="exchanges on the kernels' stream (their time is then visible between forward and adjoint)")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE")
    ap.add_argument("--library", type=int, default=1, help="1: interface exchange inside libadfem_cuda (adfem_dist_*: pack kernel, ncclSend/Recv group, "
                    "deterministic unpack); 0: the torch.distributed reference path (index_select / all_to_all_single / index_add_)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dry = args.dry_run
    if not dry:
        if not torch.cuda.is_available():
            raise SystemExit("needs CUDA devices (libadfem_cuda has no CPU fallback); --dry-run checks the host logic only")
        torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo" if dry else "nccl", **({} if dry else {"device_id": torch.device("cuda", local)}))
    L = _lib.lib()
    dev = torch.device("cpu") if dry else torch.device("cuda", local)
    for case in args.cases.split(","):
        t0 = time.perf_counter()
        part, op, cpg, note = build_case(case, args.scale, rank, world, dry)
        mesh = part.mesh
        if args.library and not dry and world > 1:
            part.use_library()
        for kv in args.opt:
            k, v = kv.split("=")
            mesh.set_option(k, int(v))
        nc = mesh.dim if op == 2 else 1
        nnz = nc * nc * int(part.rowptr[-1])
        G, E = mesh.ngauss, mesh.nelem
        gen = torch.Generator(device=dev).manual_seed(rank)
        coef = torch.rand(G * cpg, dtype=torch.float64, device=dev, generator=gen) + 0.5
        dK = torch.rand(nnz, dtype=torch.float64, device=dev, generator=gen) - 0.5
        vals = torch.zeros(nnz, dtype=torch.float64, device=dev)
        grad = torch.zeros(G * cpg, dtype=torch.float64, device=dev)
        dghost = None
        if not dry:
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            pk, pv, pd, pg = (C.c_void_p(t.data_ptr()) for t in (coef, vals, dK, grad))

        # Both interface exchanges run on a high-priority side stream, as in bench.py: replicate(dK) (input of the adjoint) overlaps the forward
        # kernel, reduce(vals) overlaps the adjoint kernel; the streams join at the start of every step.  --no-overlap keeps everything on one stream
        # (then the exchange time is what the events between forward and adjoint measure).
        overlap = (not dry) and world > 1 and not args.no_overlap
        main = torch.cuda.current_stream() if not dry else None
        side = torch.cuda.Stream(priority=-1) if overlap else None

        def step(ev=None):
            nonlocal dghost
            if dghost is None:
                dghost = torch.zeros(len(part.ghost_idx) * nc * nc, dtype=torch.float64, device=dev)
            if overlap:
                side.wait_stream(main)
                main.wait_stream(side)
                with torch.cuda.stream(side):
                    part.replicate_interface(dK, dghost, ncomp=nc)
            if ev:
                ev[0].record()
            if not dry:
                _lib.check(L.adfem_assemble_csr(mesh.handle, op, pk, pv, st))
            if ev:
                ev[1].record()
            if overlap:
                fwd_done = torch.cuda.Event()
                fwd_done.record(main)
                main.wait_stream(side)                               # the adjoint needs the replicated dK
            else:
                part.reduce_interface(vals, ncomp=nc)                # interface-row partial sums to their owners
                part.replicate_interface(dK, dghost, ncomp=nc)       # d loss / d K of the interface rows back to every contributor
            if ev:
                ev[2].record()
            if not dry:
                _lib.check(L.adfem_assemble_csr_adjoint(mesh.handle, op, pd, pg, st))
            if ev:
                ev[3].record()
            if overlap:
                side.wait_event(fwd_done)
                with torch.cuda.stream(side):
                    part.reduce_interface(vals, ncomp=nc)

        def join():
            if overlap:
                main.wait_stream(side)

        step()
        if not dry:
            torch.cuda.synchronize()
        setup = time.perf_counter() - t0
        for _ in range(args.warmup):
            step()
        join()
        K = args.steps
        if dry:
            tw = time.perf_counter()
            for _ in range(K):
                step()
            total_ms = (time.perf_counter() - tw) * 1e3
            fwd_ms = xch_ms = adj_ms = 0.0
        else:
            ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            for i in range(K):
                step(ev[i])
            join()
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            total_ms = ev[0][0].elapsed_time(end)
            fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / K
            xch_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / K
            adj_ms = sum(e[2].elapsed_time(e[3]) for e in ev) / K
        t = torch.tensor([total_ms, fwd_ms, xch_ms, adj_ms, float(E), float(part.interface_bytes * nc * nc)], dtype=torch.float64, device=dev)
        tmax, tsum = t.clone(), t.clone()
        if world > 1:
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        if rank == 0:
            ms = tmax[0].item() / K
            print(json.dumps({"case": "config" + case, "note": note, "n_gpus": world, "scaling": "strong" if case == "5g" else "weak", "dry_run": dry,
                              "exchange": "library (adfem_dist_*)" if (args.library and not dry and world > 1) else "torch.distributed",
                              "elements_total": int(tsum[4].item()), "elements_per_gpu_max": int(tmax[4].item()),
                              "ms_per_step": ms, "Melem_per_s": tsum[4].item() / (ms * 1e-3) / 1e6 if not dry else None,
                              "fwd_ms_max": tmax[1].item(), "exchange_ms_max": tmax[2].item(), "adj_ms_max": tmax[3].item(),
                              "interface_bytes_per_step_total": int(tsum[5].item()), "setup_s": round(setup, 1), "overlap": bool(overlap),
                              "options": args.opt}), flush=True)
        del part, mesh
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
