#!/usr/bin/env python
"""Benchmark of the differentiable FEM assembly hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm's CPU path (oracle port), rank 0 only

A "step" is one forward assembly (kappa -> CSR values) plus one adjoint (dK -> grad kappa) of
`compute_fem_laplace_matrix1` on BASELINE config 2: the structured P1 mesh Mesh(4096, 4096, 1/4096)
(33 554 432 triangles per GPU; weak scaling: rank r assembles row-slab r of Mesh(4096, 4096*N, h) and the
interface rows are summed over NCCL).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fwd+adjoint FEM assembly throughput (fp64)"
UNIT = "Melem/s"
# algorithmic bytes per element, SURVEY.md §8(d) config 2 (P1 triangles, scalar Laplace):
#   fwd = 4*d conn + 8*dim*nnode/E coords + 8*g kappa + 8*nnz/E values   = 12 + 8 + 24 + 28 = 72 B
#   adj = 12 + 8 + 28 (dK) + 24 (grad kappa)                             = 72 B


def alg_bytes_per_elem(nelem, nnode, nnz, d=3, dim=2, g=3, c=1):
    per = 4 * d + 8 * dim * nnode / nelem + 8 * c * g + 8 * nnz / nelem
    return per, per


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
                time.sleep(0.002)
        except Exception as ex:  # NVML missing: report it, do not fail the bench
            self.reasons.add("nvml_unavailable:" + type(ex).__name__)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_run(size, reps):
    """The reference algorithm on the host: oracle port of FemLaplaceScalar_forward/_backward (single thread — the
    reference has no threading), COO layout with duplicates exactly like the reference ops."""
    import numpy as np

    from adfem_jl_b200 import meshgen
    from oracle import oracle as O
    O.build()
    c, e = meshgen.tri_grid(size, size, 1.0 / size)
    M = O.Mesh2D(c, e)
    xy = M.gauss
    kappa = 1 + 0.5 * np.sin(2 * np.pi * xy[:, 0]) * np.cos(2 * np.pi * xy[:, 1])
    gv = np.random.default_rng(0).uniform(-1, 1, M.ngauss * 9)
    L = O.lib()
    N = M.ngauss * 9
    ind, vv, gk = np.zeros(2 * N, dtype=np.int64), np.zeros(N), np.zeros(M.ngauss)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        L.FemLaplaceScalar_forward_Julia(O._l(ind), O._d(vv), O._d(kappa))
        L.oracle_FemLaplaceScalar_backward(O._d(gk), O._d(gv))
        times.append(time.perf_counter() - t0)
    return M.nelem, times


def run_reference(args, rank):
    if rank != 0:
        return
    # bounded sample: the oracle port needs 0.25-0.5 us per triangle and step on one core; keep the whole run (warm-up + steps) within
    # about two minutes whatever K the caller asks for
    reps = max(1, args.warmup + args.steps)
    size = min(args.cpu_size, max(128, int((120.0 / reps / 0.5e-6 / 2) ** 0.5)))
    nelem, times = cpu_reference_run(size, reps)
    t = times[args.warmup:]
    ms = 1e3 * sum(t) / len(t)
    val = nelem / (ms * 1e-3) / 1e6
    sample = f"Mesh({size},{size},1/{size}) = {nelem} triangles per step (bounded sample of config 2), fwd+bwd COO ops"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config 2: structured P1 Laplace fwd+adjoint, Mesh(4096,4096,1/4096) per GPU", "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local):
    """Pin this rank (and the pinned host buffers it allocates afterwards, first touch) to the NUMA node its GPU hangs off, so that the
    end-to-end host<->device copies of several ranks do not cross sockets.  Best effort: silently skipped when sysfs has no answer."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
        if bus is None:
            import pynvml as nv
            nv.nvmlInit()
            bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(local)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = str(bus).lower()
        if len(bus.split(":")[0]) == 8:                      # NVML prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=4096, help="cells per side of the per-GPU mesh (config 2: 4096)")
    ap.add_argument("--cpu-size", type=int, default=2048, help="cells per side of the CPU baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rows-per-tile", type=int, default=0)
    ap.add_argument("--elems-per-tile", type=int, default=0)
    ap.add_argument("--adjoint-tiled", type=int, default=1)
    ap.add_argument("--tile-threads", type=int, default=0)
    ap.add_argument("--smem-budget", type=int, default=0)
    ap.add_argument("--pipeline", type=int, default=-1)
    ap.add_argument("--coef-prefetch", type=int, default=-1)
    ap.add_argument("--structured", type=int, default=1, help="0: force the general tile kernels on the structured mesh")
    ap.add_argument("--general-steps", type=int, default=20, help="extra timed steps of the general (unstructured-mesh) tile kernels, N=1 only")
    ap.add_argument("--grid-rows", type=int, default=0)
    ap.add_argument("--grid-occupancy", type=int, default=0)
    ap.add_argument("--host-chunks", type=int, default=0, help="node-row chunks of the pipelined host-buffer calls (end-to-end path)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist

    import adfem_jl_b200 as A
    from adfem_jl_b200 import _lib
    import ctypes as C
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libadfem_cuda has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    n = args.size
    t_setup = time.perf_counter()
    if world == 1:
        part = None
        mesh = A.Mesh(n, n, 1.0 / n)
    else:
        from adfem_jl_b200 import dist as adist
        part = adist.structured_slab(n, n * world, 1.0 / n, rank, world)
        mesh = part.mesh
    if args.rows_per_tile:
        mesh.set_option("rows_per_tile", args.rows_per_tile)
    if args.elems_per_tile:
        mesh.set_option("elems_per_tile", args.elems_per_tile)
    mesh.set_option("adjoint_tiled", args.adjoint_tiled)
    if args.tile_threads:
        mesh.set_option("tile_threads", args.tile_threads)
    if args.smem_budget:
        mesh.set_option("smem_budget", args.smem_budget)
    if args.pipeline >= 0:
        mesh.set_option("pipeline", args.pipeline)
    if args.coef_prefetch >= 0:
        mesh.set_option("coef_prefetch", args.coef_prefetch)
    mesh.set_option("structured", args.structured)
    if args.grid_rows:
        mesh.set_option("grid_rows", args.grid_rows)
    if args.grid_occupancy:
        mesh.set_option("grid_occupancy", args.grid_occupancy)
    if args.host_chunks:
        mesh.set_option("host_chunks", args.host_chunks)
    structured = bool(args.structured) and L.adfem_mesh_info(mesh.handle, _lib.INFO_STRUCTURED) == 1
    rowptr, colind = mesh.csr_pattern(1)
    nnz, G, E = int(rowptr[-1]), mesh.ngauss, mesh.nelem
    xy = A.gauss_nodes(mesh)
    kappa_h = 1 + 0.5 * np.sin(2 * np.pi * xy[:, 0]) * np.cos(2 * np.pi * xy[:, 1])
    del xy
    dK_h = np.random.default_rng(rank).uniform(-1, 1, nnz)
    kappa, dK = torch.from_numpy(kappa_h).cuda(), torch.from_numpy(dK_h).cuda()
    vals = torch.empty(nnz, dtype=torch.float64, device="cuda")
    grad = torch.empty(G, dtype=torch.float64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    h = mesh.handle
    pk, pv, pd, pg = (C.c_void_p(t.data_ptr()) for t in (kappa, vals, dK, grad))

    # Multi-GPU step: both interface exchanges run on a side stream so that they overlap the kernels.  replicate(dK) (owners send
    # d loss / d K of the interface rows back, input of the adjoint) overlaps the forward kernel; reduce(vals) (interface-row partial
    # sums to their owners) overlaps the adjoint kernel.  Streams join at the start of every step.
    main = torch.cuda.current_stream()
    side = torch.cuda.Stream(priority=-1) if part is not None else None      # high priority: its small kernels slip in between the CTAs of the big ones

    def step(ev=None):
        if part is not None:
            side.wait_stream(main)             # previous step's adjoint has read dK; previous reduce has finished with vals
            main.wait_stream(side)
            with torch.cuda.stream(side):
                part.replicate_interface(dK)
        if ev:
            ev[0].record()
        _lib.check(L.adfem_assemble_csr(h, 0, pk, pv, st))
        if ev:
            ev[1].record()
        if part is not None:
            fwd_done = torch.cuda.Event()
            fwd_done.record(main)
            main.wait_stream(side)             # adjoint needs the replicated dK
        _lib.check(L.adfem_assemble_csr_adjoint(h, 0, pd, pg, st))
        if ev:
            ev[2].record()
        if part is not None:
            side.wait_event(fwd_done)
            with torch.cuda.stream(side):
                part.reduce_interface(vals)

    def join():
        if part is not None:
            main.wait_stream(side)

    step()                                     # builds the mesh-static plans (untimed, reused by every later call)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    for _ in range(args.warmup):
        step()
    join()
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev_begin, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_begin.record()
    for i in range(K):
        step(ev[i])
    join()
    ev_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join()
    total_ms = ev_begin.elapsed_time(ev_end)
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    adj_ms = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    tt = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = tt.item()
    ms_per_step = total_ms / K
    value = world * E / (ms_per_step * 1e-3) / 1e6

    # ---- the general (any-mesh) tile kernels on the same mesh, for reference next to the structured fast path
    general = None
    if structured and world == 1 and args.general_steps > 0:
        mesh.set_option("structured", 0)
        step(); step(); step()
        torch.cuda.synchronize()
        gev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.general_steps)]
        for i in range(args.general_steps):
            gev[i][0].record()
            _lib.check(L.adfem_assemble_csr(h, 0, pk, pv, st))
            gev[i][1].record()
            _lib.check(L.adfem_assemble_csr_adjoint(h, 0, pd, pg, st))
            gev[i][2].record()
        torch.cuda.synchronize()
        gf = sum(e[0].elapsed_time(e[1]) for e in gev) / len(gev)
        ga = sum(e[1].elapsed_time(e[2]) for e in gev) / len(gev)
        general = {"fwd_ms": gf, "adj_ms": ga, "ms_per_step": gf + ga, "value": E / ((gf + ga) * 1e-3) / 1e6, "unit": UNIT,
                   "kernels": ["k_tile_fwd<2,1,LAPLACE,1,0>", "k_tile_adj<2,1,LAPLACE>"],
                   "plan_bytes_per_elem": L.adfem_mesh_info(h, _lib.INFO_PLAN_BYTES) / E}
        mesh.set_option("structured", 1)

    # ---- end-to-end through the host-buffer C-ABI calls (pinned host memory, copies inside the timed region)
    e2e = None
    if args.e2e_steps > 0:
        hk, hv = torch.from_numpy(kappa_h).pin_memory(), torch.empty(nnz, dtype=torch.float64).pin_memory()
        hd, hg = torch.from_numpy(dK_h).pin_memory(), torch.empty(G, dtype=torch.float64).pin_memory()
        ph = [C.c_void_p(t.data_ptr()) for t in (hk, hv, hd, hg)]

        def e2e_step():
            _lib.check(L.adfem_assemble_csr_host(h, 0, ph[0], ph[1]))
            _lib.check(L.adfem_assemble_csr_adjoint_host(h, 0, ph[2], ph[3]))
        e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        te = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * E / te.item() / 1e6, "unit": UNIT, "h2d_bytes_per_step": 8 * (G + nnz), "d2h_bytes_per_step": 8 * (nnz + G),
               "ms_per_step": te.item() * 1e3, "api": "adfem_assemble_csr_host + adfem_assemble_csr_adjoint_host (pinned host buffers)"}
        assert torch.equal(hv.cuda(), vals) or part is not None
        del hk, hv, hd, hg

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = read_peaks()
    bf72, ba72 = alg_bytes_per_elem(E, mesh.nnode, nnz)
    if structured:
        # the structured-mesh kernels take connectivity and coordinates as index arithmetic (they are implicit inputs of
        # Mesh(m,n,h)), so their compulsory streams are the coefficients and the values only: 8*g + 8*nnz/E per element
        bf = ba = 8 * 3 + 8 * nnz / E
    else:
        bf, ba = bf72, ba72
    fname = "k_grid_fwd<LAPLACE>" if structured else "k_tile_fwd<2,1,LAPLACE,1,0>"
    aname = "k_grid_adj<LAPLACE>" if structured else ("k_tile_adj<2,1,LAPLACE>" if args.adjoint_tiled else "k_csr_adj_gather<2,1,LAPLACE>")
    kern = {"fwd": {"name": fname, "ms": fwd_ms, "alg_bytes": bf * E, "GBps": bf * E / (fwd_ms * 1e-3) / 1e9},
            "adj": {"name": aname, "ms": adj_ms, "alg_bytes": ba * E, "GBps": ba * E / (adj_ms * 1e-3) / 1e9}}
    dom = "fwd" if fwd_ms >= adj_ms else "adj"
    roofline = {"bound": "hbm", "kernel": kern[dom]["name"], "achieved": kern[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["GBps"] / peak, "traffic": None, "peak_source": peak_src,
                "alg_bytes_per_elem": {"fwd": bf, "adj": ba}, "step_frac": (bf + ba) * E / (ms_per_step * 1e-3) / 1e9 / peak,
                "kernels": kern}
    if structured:
        roofline["general_mesh_accounting"] = {
            "note": "SURVEY 8(d)'s figure for a general mesh (12 B connectivity + 8 B coordinates + 24 B coefficients + 28 B values per element "
                    "and direction); the structured kernels do not move the first 20 B, so this ratio can exceed 1 and is NOT the roofline fraction",
            "bytes_per_elem": bf72, "step_ratio": (bf72 + ba72) * E / (ms_per_step * 1e-3) / 1e9 / peak}
        if general:
            general["roofline"] = {"alg_bytes_per_elem": bf72, "fwd_GBps": bf72 * E / (general["fwd_ms"] * 1e-3) / 1e9,
                                   "adj_GBps": ba72 * E / (general["adj_ms"] * 1e-3) / 1e9,
                                   "step_frac": (bf72 + ba72) * E / (general["ms_per_step"] * 1e-3) / 1e9 / peak}
    prof = os.path.join(ROOT, "profiles", "traffic.json")     # dram bytes per launch from the committed ncu capture
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get(kern[dom]["name"])
        except Exception:
            pass
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ne_cpu, times = cpu_reference_run(args.cpu_size, 3)
        best = min(times)
        cpu = {"value": ne_cpu / best / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"oracle port of FemLaplaceScalar_forward+_backward (COO with duplicates) on Mesh({args.cpu_size},{args.cpu_size}) = "
                         f"{ne_cpu} triangles, best of 3, 1 thread (the reference has no threading)"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"config 2: structured P1 Laplace fwd+adjoint (CSR mode), Mesh({n},{n},1/{n}) per GPU",
                       "elements_per_gpu": E, "nodes_per_gpu": mesh.nnode, "nnz_per_gpu": nnz, "gauss_points_per_gpu": G,
                       "l2_policy": "inputs larger than L2 (kappa %.2f GB, values %.2f GB per pass; no flush needed)" % (8 * G / 1e9, 8 * nnz / 1e9),
                       "setup_s_untimed": round(t_setup, 1), "numa_node_rank0": numa,
                       "path": "structured triangulation kernels (tri_grid.cuh): no mesh-static index data" if structured else "general tile kernels",
                       "plan_bytes_per_elem": 0.0 if structured else L.adfem_mesh_info(h, _lib.INFO_PLAN_BYTES) / E,
                       "parallelism": "element slabs; NCCL all_to_all of interface rows on a side stream (reduce(vals) overlaps the adjoint kernel, "
                                      "replicate(dK) the forward kernel)" if world > 1 else "single GPU"},
            "roofline": roofline, "general_path": general, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": 2 * K, "clocks": sampler.result()}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
