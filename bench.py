#!/usr/bin/env python
"""Benchmark of the differentiable FEM assembly hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm's CPU path (oracle port), rank 0 only
  python bench.py --config 3|4|5 ...                        # the same line for another BASELINE config (default 2 = the metric's config)

A "step" is one forward assembly (coefficients -> CSR values) plus one adjoint (dK -> coefficient gradient).  The headline
line is BASELINE config 2: `compute_fem_laplace_matrix1` on the structured P1 mesh Mesh(4096, 4096, 1/4096) (33 554 432
triangles per GPU; weak scaling: rank r assembles row-slab r of Mesh(4096, 4096*N, h) and the interface rows are summed at
their owners over NCCL inside libadfem_cuda).  The line also carries `extra.configs`: BASELINE configs 3, 4 and 5 at their
full sizes on the same GPUs (config 5 = the fixed 48 M-tetrahedra mesh Mesh3(215, 215, 208, h), STRONG scaling over the
ranks), each with its roofline fraction on SURVEY 8(d)'s algorithmic bytes.  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fwd+adjoint FEM assembly throughput (fp64)"
UNIT = "Melem/s"
# algorithmic bytes per element and direction, SURVEY.md 8(d):
#   4*d connectivity + 8*dim*nnode/E coordinates + 8*c*g coefficients + 8*nnz/E values
#   config 2 (P1 triangles, scalar):      12 + 8 + 24 + 28 = 72 B
#   config 3 (P1 triangles, 3x3 H):       12 + 8 + 216 + 112 = 348 B
#   config 4 (P2 triangles, scalar):      24 + 8 + 48 + 184 = 264 B
#   config 5 (P1 tetrahedra, 6x6 H):      16 + 5 + 1152 + 190 = 1363 B


def alg_bytes_per_elem(mesh, nnz_values, cpg):
    return 4 * mesh.elem_ndof + 8 * mesh.dim * mesh.nnode / mesh.nelem + 8 * cpg * mesh.gauss_per_elem + 8 * nnz_values / mesh.nelem


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def read_traffic():
    """dram bytes per launch from the committed ncu captures (profiles/traffic.json): either a plain number (capture taken at the
    bench size) or {"bytes": B, "elements": E} (capture at E elements, scaled linearly to the bench size)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def traffic_of(tab, name, E):
    v = tab.get(name)
    if v is None:
        return None
    if isinstance(v, dict):
        return v["bytes"] * E / v["elements"]
    return v


class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons; samples taken while `armed` belong to the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.armed, self.samples, self.reasons, self.max_mhz, self.err = index, False, False, [], set(), None, None
        self.timed = False          # inside the K timed steps (a subset of `armed`, which also covers the warm-up steps right before them)
        self.n_timed = 0
        self.ready = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                     "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
            self.ready.set()
            while not self.stop_flag:
                if self.armed:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                    self.n_timed += 1 if self.timed else 0
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for k, bit in names.items():
                        if r & bit:
                            self.reasons.add(k)
                time.sleep(0.0005)
        except Exception as ex:  # NVML missing: report it, do not fail the bench
            self.err = type(ex).__name__
            self.reasons.add("nvml_unavailable:" + self.err)
            self.ready.set()

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "samples_in_timed_region": self.n_timed,
                "window": "warm-up steps + the K timed steps of the headline config (same kernels back to back; an NVML query takes 1-20 ms, the timed region 60 ms)"}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm on the host cores (oracle port; the reference itself cannot be built here)
# ------------------------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_reference_run(size_m, size_n, reps, threads, single_reps=0):
    """Oracle port of FemLaplaceScalar_forward / _backward (COO layout with duplicates, exactly the reference ops' outputs) on
    Mesh(size_m, size_n, 1/size_m).  threads == 1 is the reference as it is (it has no threading); threads > 1 runs the same two loops
    over contiguous element blocks, one per host thread (every slot written by exactly one thread: bit-identical outputs).
    Returns (elements, times with `threads`, times with one thread (single_reps of them))."""
    import numpy as np

    from adfem_jl_b200 import meshgen
    from oracle import oracle as O
    O.build()
    c, e = meshgen.tri_grid(size_m, size_n, 1.0 / size_m)
    M = O.Mesh2D(c, e)
    xy = M.gauss
    kappa = 1 + 0.5 * np.sin(2 * np.pi * xy[:, 0]) * np.cos(2 * np.pi * xy[:, 1])
    gv = np.random.default_rng(0).uniform(-1, 1, M.ngauss * 9)
    L = O.lib()
    N = M.ngauss * 9
    ind, vv, gk = np.zeros(2 * N, dtype=np.int64), np.zeros(N), np.zeros(M.ngauss)

    def run(nt):
        t0 = time.perf_counter()
        if nt > 1:
            L.oracle_FemLaplaceScalar_forward_backward_mt(C.c_int(nt), O._l(ind), O._d(vv), O._d(kappa), O._d(gk), O._d(gv))
        else:
            L.FemLaplaceScalar_forward_Julia(O._l(ind), O._d(vv), O._d(kappa))
            L.oracle_FemLaplaceScalar_backward(O._d(gk), O._d(gv))
        return time.perf_counter() - t0
    times = [run(threads) for _ in range(reps)]
    single = [run(1) for _ in range(single_reps)]
    return M.nelem, times, single


def run_reference(args, rank):
    """`--impl reference`: rank 0 alone; all host threads; config 2 at its full size when the whole run fits the time budget
    (mesh tables ~70 s + ~1 s per step with 16 threads), otherwise a slab of fewer cell rows of the same mesh (stated in `sample`)."""
    if rank != 0:
        return
    threads = host_threads()
    reps = max(1, args.warmup + args.steps)
    m = args.size
    # measured: ~0.5 us per triangle and step on one core, about 1/6 of that per step with >= 8 threads (memory bound); table build ~2 us per triangle
    per_tri = 0.5e-6 / min(threads, 6)
    budget = float(args.cpu_budget_s)
    rows = int(min(m, max(64, budget / (2 * m * (2.0e-6 + reps * per_tri)))))
    nelem, times, _ = cpu_reference_run(m, rows, reps, threads)
    t = times[args.warmup:] or times
    ms = 1e3 * sum(t) / len(t)
    val = nelem / (ms * 1e-3) / 1e6
    full = rows == m
    sample = (f"Mesh({m},{rows},1/{m}) = {nelem} triangles per step (" + ("the full config-2 mesh" if full else "a row slab of the config-2 mesh, bounded by the time budget")
              + f"), fwd+bwd COO ops, {threads} host threads over element blocks")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"config 2: structured P1 Laplace fwd+adjoint, Mesh({m},{m},1/{m}) per GPU", "sample": sample, "same_mesh_as_gpu_arm": full},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local):
    """Pin this rank (and the pinned host buffers it allocates afterwards, first touch) to the NUMA node its GPU hangs off, so that the
    end-to-end host<->device copies of several ranks do not cross sockets.  Returns (node, why): node None when the box has a single
    NUMA node or sysfs has no answer (`why` says which)."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(local)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = str(bus).lower()
        if len(bus.split(":")[0]) == 8:                      # NVML prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0 or len(nodes) <= 1:
            return None, "single NUMA node (%d in sysfs, device reports %d)" % (len(nodes), node)
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node, "bound to %d cpus" % len(cpus)
        return None, "no allowed cpu on node %d" % node
    except Exception as ex:
        return None, "unavailable: " + type(ex).__name__


# ------------------------------------------------------------------------------------------------------------------
# cases = BASELINE.json configs
# ------------------------------------------------------------------------------------------------------------------
CASE_KERNELS = {          # (forward kernels, adjoint kernels) of the default path, names as ncu prints them (keys of profiles/traffic.json)
    "2": (["k_grid_fwd<LAPLACE>"], ["k_grid_adj<LAPLACE>"]),
    "2g": (["k_tile_fwd<2,1,LAPLACE,1,0>"], ["k_tile_adj<2,1,LAPLACE>"]),
    "2m": (["k_grid_fwd<LAPLACE,MAPPED>"], ["k_grid_adj<LAPLACE,MAPPED>"]),
    "3": (["k_grid_elast_fwd<LAPLACE>"], ["k_grid_elast_adj<LAPLACE>"]),
    "3m": (["k_grid_elast_fwd<MAPPED>"], ["k_grid_elast_adj<MAPPED>"]),
    "4l": (["k_tile_fwd<2,2,LAPLACE,0,1>"], ["k_tile_adj<2,2,LAPLACE>"]),
    "4": (["k_tile_fwd<2,2,LAPLACE,0,1>"], ["k_tile_adj<2,2,LAPLACE>"]),
    "4o": (["k_tile_fwd<2,2,LAPLACE,0,1>"], ["k_tile_adj<2,2,LAPLACE>"]),
    "4m": (["k_tile_fwd<2,2,MASS,0,1>"], ["k_tile_adj<2,2,MASS>"]),
    "5": (["k_tet_presum_xg<4>", "k_tet_node_fwd"], ["k_tet_grid_elast_adj<3>"]),
}


def build_case(case, rank, world, scale=1.0, size=None, numbering="random", host_only=False):
    """-> (mesh, partition or None, op, coefficients per Gauss point, description, scaling).  world > 1: element blocks, one per rank."""
    import numpy as np

    import adfem_jl_b200 as A
    from adfem_jl_b200 import dist as adist
    from adfem_jl_b200 import meshgen
    kw = dict(host_only=True) if host_only else {}
    if case in ("2", "2g"):
        n = size or max(4, int(4096 * scale))
        if world == 1:
            return A.Mesh(n, n, 1.0 / n, **kw), None, 0, 1, f"config 2: structured P1 Laplace fwd+adjoint (CSR mode), Mesh({n},{n},1/{n}) per GPU", "weak"
        part = adist.structured_slab(n, n * world, 1.0 / n, rank, world, **kw)
        return part.mesh, part, 0, 1, f"config 2: structured P1 Laplace fwd+adjoint (CSR mode), Mesh({n},{n},1/{n}) per GPU", "weak"
    if case == "2m":
        # config 2's connectivity on mapped + jittered node positions: what one moved node does to the structured fast path (index-free kernels
        # with positions from the coordinate array instead of the general tile kernels)
        import numpy as np
        n = size or max(4, int(4096 * scale))
        coords, elems = meshgen.tri_grid(n, n, 1.0 / n)
        rng = np.random.default_rng(3)
        coords = np.stack([coords[:, 0] + 0.02 * np.sin(3.0 * coords[:, 1]), coords[:, 1] + 0.02 * np.cos(2.0 * coords[:, 0])], 1) + rng.uniform(-0.2 / n, 0.2 / n, coords.shape)
        note = f"config 2 connectivity, Mesh({n},{n},1/{n}), on mapped + jittered node positions (structured kernels, positions from the coordinate array)"
        return A.Mesh(coords, elems, **kw), None, 0, 1, note, "weak"
    if case == "3m":
        # config 3's connectivity on mapped + jittered node positions (MAPPED instantiations of the structured elasticity kernels)
        import numpy as np
        m, nl = max(4, int(4096 * scale)), max(2, int(2048 * scale))
        coords, elems = meshgen.tri_grid(m, nl, 1.0 / m)
        rng = np.random.default_rng(4)
        coords = np.stack([coords[:, 0] + 0.02 * np.sin(3.0 * coords[:, 1]), coords[:, 1] + 0.02 * np.cos(2.0 * coords[:, 0])], 1) + rng.uniform(-0.2 / m, 0.2 / m, coords.shape)
        note = f"config 3 connectivity, Mesh({m},{nl},1/{m}), on mapped + jittered node positions (structured elasticity kernels, positions from the coordinate array)"
        return A.Mesh(coords, elems, **kw), None, 2, 9, note, "weak"
    if case == "3":
        m, nl = max(4, int(4096 * scale)), max(2, int(2048 * scale))
        note = f"config 3: P1 linear elasticity, per-Gauss-point 3x3 tangent H, Mesh({m},{nl},1/{m}) per GPU ({2 * m * nl} triangles)"
        if world == 1:
            return A.Mesh(m, nl, 1.0 / m, **kw), None, 2, 9, note, "weak"
        part = adist.structured_slab(m, nl * world, 1.0 / m, rank, world, **kw)
        return part.mesh, part, 2, 9, note, "weak"
    if case in ("4", "4l", "4m", "4o"):
        # jittered grid, random diagonals, nodes AND elements randomly renumbered, 16 M triangles.  On N > 1 GPUs (and in case
        # "4o" on one GPU) the elements are first put in Morton order of their centroids — SURVEY 8(e)'s space-filling-curve renumbering that makes
        # contiguous element blocks compact; the node numbering stays random.  The mesh is fixed, N ranks take one block each: strong scaling.
        n = max(4, int(2828 * scale))
        what = {"4": "Laplace and mass", "4l": "Laplace", "4m": "mass", "4o": "Laplace"}[case]
        coords, elems = meshgen.jitter_unstructured(n, n, 1.0 / n, seed=2, permute=(numbering == "random"))
        morton = world > 1 or case == "4o"
        note = (f"config 4: P2 {what}, jittered triangulation with random diagonals, {n} x {n} cells ({2 * n * n} triangles), "
                + ("nodes randomly renumbered" if numbering == "random" else "generator (row-major) node numbering") + ", elements "
                + ("in Morton order of their centroids (SURVEY 8e: compact contiguous element blocks)" if morton else
                   ("randomly renumbered" if numbering == "random" else "in generator order")))
        if morton:
            elems = np.take(elems, meshgen.morton_element_order(coords, elems), axis=0)
        op = 1 if case == "4m" else 0
        if world == 1:
            return A.Mesh(coords, elems, degree=2, **kw), None, op, 1, note, "strong"
        part, _ = adist.partition_elements(coords, elems, rank, world, degree=2, **kw)
        return part.mesh, part, op, 1, note + "; one contiguous element block per GPU", "strong"
    if case == "5":
        n = max(2, int(215 * scale))
        l = max(2 * world, int(208 * scale) // (2 * world) * (2 * world))
        note = (f"config 5: P1 tetrahedral elasticity, per-Gauss-point 6x6 Voigt tangent, the fixed global mesh Mesh3({n},{n},{l},1/{n}) = {5 * n * n * l} tetrahedra "
                f"in z-slabs of {l // world} cube layers per GPU")
        if world == 1:
            return A.Mesh3(n, n, l, 1.0 / n, **kw), None, 2, 36, note, "strong"
        part = adist.structured_slab3(n, l, 1.0 / n, rank, world, **kw)
        return part.mesh, part, 2, 36, note, "strong"
    raise SystemExit("unknown case " + case)


def link_probe(world, nbytes=1 << 29, reps=4):
    """What the box's host <-> device path gives when every rank copies at once: `reps` x (H2D + D2H of `nbytes` each, concurrently on two
    streams, pinned host memory).  Context for the end-to-end number: the host-buffer calls cannot be faster than this."""
    import torch
    import torch.distributed as dist
    h_in, h_out = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory(), torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
    h_in.fill_(1.0)
    d_in, d_out = torch.empty_like(h_in, device="cuda"), torch.ones(nbytes // 8, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for timed in (False, True):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps if timed else 1):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"duplex_GBps_all_ranks": world * 2 * reps * nbytes / t.item() / 1e9, "bytes_each_way_per_rank": reps * nbytes,
            "how": "all ranks at once, H2D and D2H concurrently, pinned memory, max time over ranks"}


class DeviceStep:
    """One forward + adjoint CSR assembly of a case on this rank's GPU, device-resident inputs, with the interface exchanges of the
    multi-GPU path on a high-priority side stream: replicate(dK) (owners send d loss / d K of the interface rows back, input of the
    adjoint) overlaps the forward kernel; reduce(vals) (interface-row partial sums to their owners) overlaps the adjoint kernel."""

    def __init__(self, mesh, part, op, cpg, coef_h=None, dK_h=None, seed=0, library=True):
        import torch

        from adfem_jl_b200 import _lib
        self.torch, self._lib, self.L = torch, _lib, _lib.lib()
        self.mesh, self.part, self.op, self.cpg = mesh, part, op, cpg
        self.nc = mesh.dim if op == 2 else 1
        self.nnz_s = int(self.L.adfem_csr_nnz(mesh.handle, 1))       # builds the symbolic pattern; no host copy of rowptr / colind
        if self.nnz_s < 0:
            raise RuntimeError(_lib.last_error())
        self.nnz = self.nc * self.nc * self.nnz_s
        G = mesh.ngauss
        gen = torch.Generator(device="cuda").manual_seed(seed)
        if coef_h is not None:
            self.coef = coef_h if torch.is_tensor(coef_h) else torch.from_numpy(coef_h).cuda()
            self.dK = dK_h if torch.is_tensor(dK_h) else torch.from_numpy(dK_h).cuda()
        else:
            self.coef = torch.rand(G * cpg, dtype=torch.float64, device="cuda", generator=gen) + 0.5
            self.dK = torch.rand(self.nnz, dtype=torch.float64, device="cuda", generator=gen) - 0.5
        self.vals = torch.empty(self.nnz, dtype=torch.float64, device="cuda")
        self.grad = torch.empty(G * cpg, dtype=torch.float64, device="cuda")
        self.main = torch.cuda.current_stream()
        self.st = C.c_void_p(self.main.cuda_stream)
        self.p = [C.c_void_p(t.data_ptr()) for t in (self.coef, self.vals, self.dK, self.grad)]
        # The exchanges overlap the kernels only on the structured paths (many short CTAs: the high-priority side stream's kernels slip in
        # between them).  The tile kernels are PERSISTENT (every SM slot is held for the whole launch), so a concurrent ncclSend/ncclRecv
        # kernel either waits for them to end or sits on an SM waiting for a peer whose own NCCL kernel cannot start — measured on config 4 at
        # N = 8: forward 0.38 -> 0.63 ms and the adjoint waiting 0.29 ms for an exchange that takes 0.06 ms alone.  There the exchanges run on
        # the main stream between the kernels.
        self.overlap = part is not None and int(self.L.adfem_mesh_info(mesh.handle, _lib.INFO_STRUCTURED)) != 0
        self.side = torch.cuda.Stream(priority=-1) if self.overlap else None
        self.dghost = None
        if part is not None:
            if library:
                try:
                    part.use_library()
                except Exception as ex:        # e.g. libnccl.so.2 cannot be bound: keep measuring, through torch.distributed, and say so
                    library = False
                    self.exchange_note = "library exchange unavailable (%s: %s)" % (type(ex).__name__, str(ex)[:120])
            self.dghost = torch.zeros(len(part.ghost_idx) * self.nc * self.nc, dtype=torch.float64, device="cuda")
        self.exchange = None if part is None else ("library (adfem_dist_*: pack kernel, ncclSend/ncclRecv group, deterministic unpack)" if library else
                                                   "torch.distributed all_to_all_single" + ("; " + self.exchange_note if getattr(self, "exchange_note", None) else ""))

    def forward(self):
        self._lib.check(self.L.adfem_assemble_csr(self.mesh.handle, self.op, self.p[0], self.p[1], self.st))

    def adjoint(self):
        self._lib.check(self.L.adfem_assemble_csr_adjoint(self.mesh.handle, self.op, self.p[2], self.p[3], self.st))

    def __call__(self, ev=None):
        torch, part, main, side = self.torch, self.part, self.main, self.side
        if part is not None and not self.overlap:       # in order on the main stream
            if ev:
                ev[0].record()
            self.forward()
            if ev:
                ev[1].record()
            part.reduce_interface(self.vals, ncomp=self.nc)
            part.replicate_interface(self.dK, self.dghost, ncomp=self.nc)
            if ev:
                ev[2].record()
            self.adjoint()
            if ev:
                ev[3].record()
            return
        if part is not None:
            side.wait_stream(main)             # previous step's adjoint has read dK; previous reduce has finished with vals
            main.wait_stream(side)
            with torch.cuda.stream(side):
                part.replicate_interface(self.dK, self.dghost, ncomp=self.nc)
        if ev:
            ev[0].record()
        self.forward()
        if ev:
            ev[1].record()
        if part is not None:
            fwd_done = torch.cuda.Event()
            fwd_done.record(main)
            main.wait_stream(side)             # the adjoint needs the replicated dK
        if ev:
            ev[2].record()
        self.adjoint()
        if ev:
            ev[3].record()
        if part is not None:
            side.wait_event(fwd_done)
            with torch.cuda.stream(side):
                part.reduce_interface(self.vals, ncomp=self.nc)

    def join(self):
        if self.side is not None:
            self.main.wait_stream(self.side)

    def exchange_alone_ms(self, reps=10):
        """the two interface exchanges of a step by themselves on the main stream (no kernel to hide behind): pack kernel -> ncclSend/ncclRecv
        group -> unpack kernel, once for reduce(vals) and once for replicate(dK); max over ranks"""
        if self.part is None:
            return None
        import torch.distributed as dist
        torch = self.torch
        self.join()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for timed in (False, True):
            dist.barrier()
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(reps if timed else 2):
                self.part.reduce_interface(self.vals, ncomp=self.nc)
                self.part.replicate_interface(self.dK, self.dghost, ncomp=self.nc)
            ev1.record()
            torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1) / reps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()


def time_steps(step, K, warmup, world, sampler=None):
    """W untimed steps, then exactly K steps between barrier + synchronize on both sides; CUDA events on the launching stream;
    returns (ms per step = max over ranks, forward ms, adjoint ms, wait-for-exchange ms) — the last three averaged on this rank."""
    import torch
    import torch.distributed as dist
    if sampler is not None:
        sampler.armed = True
    for _ in range(max(warmup, 50 if sampler is not None else 0)):      # the headline run warms up for at least 50 steps so that the clock sampler sees the load
        step()
    step.join()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    ev_begin, ev_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.timed = True
    ev_begin.record()
    for i in range(K):
        step(ev[i])
    step.join()
    ev_end.record()
    torch.cuda.synchronize()
    if sampler is not None:
        sampler.armed = sampler.timed = False
    if world > 1:
        dist.barrier()
    total_ms = ev_begin.elapsed_time(ev_end)
    f = sum(e[0].elapsed_time(e[1]) for e in ev) / K
    x = sum(e[1].elapsed_time(e[2]) for e in ev) / K
    a = sum(e[2].elapsed_time(e[3]) for e in ev) / K
    tt = torch.tensor([total_ms, f, a, x], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return tt[0].item() / K, tt[1].item(), tt[2].item(), tt[3].item()


def case_records(case, rank, world, steps, warmup, scale, peak, traffic, library=True, numbering="random", opts=()):
    """Builds a case, times it, frees it; returns the records for `extra.configs` (rank 0) or [] (config 4: Laplace and mass on one mesh)."""
    import torch
    import torch.distributed as dist

    from adfem_jl_b200 import _lib
    L = _lib.lib()
    t0 = time.perf_counter()
    mesh, part, op0, cpg, note, scaling = build_case(case, rank, world, scale, numbering=numbering)
    for k, v in opts:
        mesh.set_option(k, v)
    recs = []
    for sub, op in ((("4l", 0), ("4m", 1)) if case == "4" else ((case, op0),)):
        step = DeviceStep(mesh, part, op, cpg, seed=rank, library=library)
        step()
        step.join()
        torch.cuda.synchronize()
        setup = time.perf_counter() - t0
        ms, f_ms, a_ms, x_ms = time_steps(step, steps, warmup, world)
        xa_ms = step.exchange_alone_ms()
        E = mesh.nelem
        b = alg_bytes_per_elem(mesh, step.nnz, cpg)
        if case in ("2m", "3m"):
            b -= 4 * mesh.elem_ndof        # structured connectivity is index arithmetic: coordinates (8 B), coefficients (24 B) and values (28 B) per element are what moves
        cnt = torch.tensor([float(E), float(part.interface_bytes * step.nc * step.nc) if part is not None else 0.0, b * E], dtype=torch.float64, device="cuda")
        mx = torch.tensor([float(E)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(cnt)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        Etot, ibytes, bytes_tot = cnt[0].item(), cnt[1].item(), cnt[2].item()
        fk, ak = CASE_KERNELS[sub]
        tr_f = [traffic_of(traffic, k, E) for k in fk]
        tr_a = [traffic_of(traffic, k, E) for k in ak]
        recs.append({"case": "config" + sub, "workload": note, "scaling": scaling, "n_gpus": world, "elements_total": int(Etot), "elements_per_gpu_max": int(mx.item()),
                     "ms_per_step": ms, "fwd_ms": f_ms, "adj_ms": a_ms, "wait_for_exchange_ms": x_ms, "Melem_per_s": Etot / (ms * 1e-3) / 1e6,
                     "roofline": {"alg_bytes_per_elem_per_direction": b, "fwd_GBps": b * E / (f_ms * 1e-3) / 1e9, "adj_GBps": b * E / (a_ms * 1e-3) / 1e9,
                                  "fwd_frac": b * E / (f_ms * 1e-3) / 1e9 / peak, "adj_frac": b * E / (a_ms * 1e-3) / 1e9 / peak,
                                  "step_frac": 2 * bytes_tot / world / (ms * 1e-3) / 1e9 / peak, "peak": peak, "unit": "GB/s",
                                  "note": "fwd / adj: this rank's kernels (max over ranks of the time); step: all ranks' algorithmic bytes / N / step time",
                                  "kernels": {"fwd": fk, "adj": ak},
                                  "traffic": {"fwd": (sum(tr_f) if all(t is not None for t in tr_f) else None), "adj": (sum(tr_a) if all(t is not None for t in tr_a) else None),
                                              "alg_bytes_per_launch": b * E}},
                     "plan_bytes_per_elem": L.adfem_mesh_info(mesh.handle, _lib.INFO_PLAN_BYTES) / E, "setup_s_untimed": round(setup, 1),
                     "exchange": step.exchange, "exchange_overlaps_kernels": bool(step.overlap), "interface_bytes_per_step_total": int(ibytes), "exchange_alone_ms": xa_ms,
                     "steps": steps, "warmup": warmup})
        del step
        torch.cuda.empty_cache()
        t0 = time.perf_counter()
    del part, mesh
    torch.cuda.empty_cache()
    return recs if rank == 0 else []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="2", choices=["2", "2m", "3", "3m", "4", "4l", "4m", "4o", "5"], help="BASELINE config of the headline line (default 2, the metric's config)")
    ap.add_argument("--extra-configs", default="2m,3,3m,4,4o,5", help="comma list of the other configs timed into extra.configs ('none' to skip)")
    ap.add_argument("--extra-steps", type=int, default=10)
    ap.add_argument("--scale", type=float, default=1.0, help="shrinks the edge counts of configs 3-5 (smoke runs)")
    ap.add_argument("--numbering", default="random", choices=["random", "generator"], help="config 4: node / element numbering of the synthetic unstructured mesh")
    ap.add_argument("--library-exchange", type=int, default=1, help="multi-GPU interface exchange: 1 = inside libadfem_cuda (adfem_dist_*), 0 = torch.distributed")
    ap.add_argument("--size", type=int, default=4096, help="cells per side of the per-GPU mesh (config 2: 4096)")
    ap.add_argument("--cpu-size", type=int, default=2048, help="cell rows of the CPU baseline sample of the N=1 line (Mesh(size, cpu_size))")
    ap.add_argument("--cpu-budget-s", type=float, default=170.0, help="--impl reference: wall-clock budget that bounds the sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--structured", type=int, default=1, help="0: force the general tile kernels on the structured mesh")
    ap.add_argument("--general-steps", type=int, default=20, help="extra timed steps of the general (unstructured-mesh) tile kernels, N=1 only")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="adfem_set_option on the headline mesh (and on the extra configs)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    if args.config == "4":
        args.config = "4l"

    import numpy as np
    import torch
    import torch.distributed as dist

    import adfem_jl_b200 as A
    from adfem_jl_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libadfem_cuda has no CPU fallback")
    torch.cuda.set_device(local)
    numa, numa_why = bind_to_gpu_numa_node(local) if world > 1 else (None, "single rank")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    peak, peak_src = read_peaks()
    traffic = read_traffic()
    sampler = ClockSampler(local)
    sampler.start()
    sampler.ready.wait(10)
    opts = [(kv.split("=")[0], int(kv.split("=")[1])) for kv in args.opt]

    # ---- headline case
    case = args.config
    t_setup = time.perf_counter()
    mesh, part, op, cpg, note, scaling = build_case(case, rank, world, args.scale if case != "2" else 1.0, size=args.size, numbering=args.numbering)
    mesh.set_option("structured", args.structured)
    for k, v in opts:
        mesh.set_option(k, v)
    structured = case == "2" and bool(args.structured) and L.adfem_mesh_info(mesh.handle, _lib.INFO_STRUCTURED) == 1
    E, G = mesh.nelem, mesh.ngauss
    coef_h = dK_h = None
    if case == "2":
        nnz_scalar = int(L.adfem_csr_nnz(mesh.handle, 1))
        if nnz_scalar < 0:
            raise RuntimeError(_lib.last_error())
        xy = A.gauss_nodes_soa(mesh)                              # (2, G): kappa(x, y) = 1 + sin(2 pi x) cos(2 pi y) / 2 is evaluated on the device (setup time)
        gx, gy = torch.from_numpy(xy[0]).cuda(), torch.from_numpy(xy[1]).cuda()
        coef_h = 1 + 0.5 * torch.sin(2 * np.pi * gx) * torch.cos(2 * np.pi * gy)
        del xy, gx, gy
        dK_h = np.random.default_rng(rank).uniform(-1, 1, nnz_scalar)
    step = DeviceStep(mesh, part, op, cpg, coef_h, dK_h, seed=rank, library=bool(args.library_exchange))
    nnz = step.nnz
    step()                                     # builds the mesh-static plans (untimed, reused by every later call)
    step.join()
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    K = args.steps
    ms_per_step, fwd_ms, adj_ms, xch_ms = time_steps(step, K, args.warmup, world, sampler)
    cnt = torch.tensor([float(E)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(cnt)
    Etot = cnt.item()
    xa_ms = step.exchange_alone_ms()
    value = Etot / (ms_per_step * 1e-3) / 1e6
    h = mesh.handle

    # ---- the general (any-mesh) tile kernels on the same mesh, for reference next to the structured fast path
    general = None
    if structured and world == 1 and args.general_steps > 0:
        mesh.set_option("structured", 0)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        _, gf, ga, _ = time_steps(step, args.general_steps, 0, 1)
        general = {"fwd_ms": gf, "adj_ms": ga, "ms_per_step": gf + ga, "value": E / ((gf + ga) * 1e-3) / 1e6, "unit": UNIT,
                   "kernels": CASE_KERNELS["2g"][0] + CASE_KERNELS["2g"][1],
                   "plan_bytes_per_elem": L.adfem_mesh_info(h, _lib.INFO_PLAN_BYTES) / E}
        mesh.set_option("structured", 1)
        step()                                 # vals of the structured path again (compared with the host path below)

    # ---- end-to-end through the host-buffer C-ABI calls (pinned host memory, copies inside the timed region)
    e2e = None
    if args.e2e_steps > 0:
        hk, hv = step.coef.cpu().pin_memory(), torch.empty(nnz, dtype=torch.float64).pin_memory()
        hd, hg = step.dK.cpu().pin_memory(), torch.empty(G * cpg, dtype=torch.float64).pin_memory()
        ph = [C.c_void_p(t.data_ptr()) for t in (hk, hv, hd, hg)]

        def e2e_step():
            _lib.check(L.adfem_assemble_csr_host(h, op, ph[0], ph[1]))
            _lib.check(L.adfem_assemble_csr_adjoint_host(h, op, ph[2], ph[3]))
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
        bi, bo = 8 * (G * cpg + nnz), 8 * (nnz + G * cpg)
        e2e = {"value": Etot / te.item() / 1e6, "unit": UNIT, "h2d_bytes_per_step": bi, "d2h_bytes_per_step": bo,
               "ms_per_step": te.item() * 1e3, "api": "adfem_assemble_csr_host + adfem_assemble_csr_adjoint_host (pinned host buffers)",
               "host_link_GBps_all_ranks": world * (bi + bo) / te.item() / 1e9, "link_probe": link_probe(world),
               "note": "bound by the host <-> device path: compare host_link_GBps_all_ranks (what the step moved) with link_probe (plain copies of all "
                       "ranks at once); at N>1 the ranks share one host memory system / PCIe root (one NUMA node on this box).  Interface rows of the "
                       "multi-rank step stay partial sums on the host copy (the exchange runs on device buffers)"}
        if part is None:
            assert torch.equal(hv.cuda(), step.vals), "host-buffer path and device path disagree"
        del hk, hv, hd, hg

    # ---- the other BASELINE configs at full size (extra.configs)
    extras = []
    headline_nc, headline_exchange = step.nc, step.exchange
    plan_bytes = L.adfem_mesh_info(h, _lib.INFO_PLAN_BYTES) / E
    nnode = mesh.nnode
    b_general = alg_bytes_per_elem(mesh, nnz, cpg)
    ibytes = float(part.interface_bytes * step.nc * step.nc) if part is not None else 0.0
    xc = [c for c in args.extra_configs.split(",") if c and c != "none" and c != case and not (c == "4" and case in ("4l", "4m")) and not (c in ("2m", "3m", "4o") and world > 1)]
    if xc:
        del step, part, mesh
        torch.cuda.empty_cache()
        for c in xc:
            try:
                recs = case_records(c, rank, world, args.extra_steps, 3, args.scale, peak, traffic, bool(args.library_exchange), args.numbering, opts)
            except Exception as ex:          # an extra config must not cost the headline line
                recs = [{"case": "config" + c, "error": "%s: %s" % (type(ex).__name__, str(ex)[:300])}] if rank == 0 else []
                torch.cuda.empty_cache()
            extras.extend(recs)
    sampler.stop_flag = True
    sampler.join()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if structured:
        # the structured-mesh kernels take connectivity and coordinates as index arithmetic (they are implicit inputs of
        # Mesh(m,n,h)), so their compulsory streams are the coefficients and the values only: 8*g + 8*nnz/E per element
        bf = ba = 8 * 3 + 8 * nnz / E
    elif case in ("2m", "3m"):
        bf = ba = b_general - 12           # mapped grid: the connectivity is index arithmetic, coordinates + coefficients + values move
    else:
        bf = ba = b_general
    key = case if (case != "2" or structured) else "2g"
    fnames, anames = CASE_KERNELS[key]
    kern = {"fwd": {"name": "+".join(fnames), "ms": fwd_ms, "alg_bytes": bf * E, "GBps": bf * E / (fwd_ms * 1e-3) / 1e9},
            "adj": {"name": "+".join(anames), "ms": adj_ms, "alg_bytes": ba * E, "GBps": ba * E / (adj_ms * 1e-3) / 1e9}}
    dom = "fwd" if fwd_ms >= adj_ms else "adj"
    tr = [traffic_of(traffic, k, E) for k in (fnames if dom == "fwd" else anames)]
    roofline = {"bound": "hbm", "kernel": kern[dom]["name"], "achieved": kern[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["GBps"] / peak, "frac_of_nominal_8TBps": kern[dom]["GBps"] / 8000.0,
                "traffic": (sum(tr) if all(t is not None for t in tr) else None), "peak_source": peak_src,
                "alg_bytes_per_elem": {"fwd": bf, "adj": ba}, "step_frac": (bf + ba) * E / (ms_per_step * 1e-3) / 1e9 / peak,
                "kernels": kern}
    if structured:
        roofline["general_mesh_accounting"] = {
            "note": "SURVEY 8(d)'s figure for a general mesh (12 B connectivity + 8 B coordinates + 24 B coefficients + 28 B values per element "
                    "and direction); the structured kernels do not move the first 20 B, so this ratio can exceed 1 and is NOT the roofline fraction",
            "bytes_per_elem": b_general, "step_ratio": 2 * b_general * E / (ms_per_step * 1e-3) / 1e9 / peak}
        if general:
            general["roofline"] = {"alg_bytes_per_elem": b_general, "fwd_GBps": b_general * E / (general["fwd_ms"] * 1e-3) / 1e9,
                                   "adj_GBps": b_general * E / (general["adj_ms"] * 1e-3) / 1e9,
                                   "step_frac": 2 * b_general * E / (general["ms_per_step"] * 1e-3) / 1e9 / peak,
                                   "traffic": {"fwd": traffic_of(traffic, CASE_KERNELS["2g"][0][0], E), "adj": traffic_of(traffic, CASE_KERNELS["2g"][1][0], E)}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = host_threads()
        ne_cpu, times, single = cpu_reference_run(args.size, args.cpu_size, 3, threads, single_reps=1)
        best = min(times)
        cpu = {"value": ne_cpu / best / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"oracle port of FemLaplaceScalar_forward+_backward (COO with duplicates) on Mesh({args.size},{args.cpu_size},1/{args.size}) = "
                         f"{ne_cpu} triangles (a row slab of the config-2 mesh), best of 3, {threads} host threads over element blocks",
               "single_thread_value": ne_cpu / min(single) / 1e6,
               "single_thread_note": "the reference op as it is (no threading), same sample, one pass"}
    launches = {"2": 2, "2g": 2, "2m": 2, "3": 2, "3m": 2, "4": 2, "4l": 2, "4m": 2, "4o": 2, "5": 3}[key] + (4 if world > 1 else 0)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": note, "elements_per_gpu": E, "nodes_per_gpu": nnode, "nnz_per_gpu": nnz, "gauss_points_per_gpu": G,
                       "l2_policy": "inputs larger than L2 (coefficients %.2f GB, values %.2f GB per pass; no flush needed)" % (8 * G * cpg / 1e9, 8 * nnz / 1e9),
                       "setup_s_untimed": round(t_setup, 1), "numa_node_rank0": numa, "numa_note": numa_why,
                       "untimed_steps_before_timing": max(args.warmup, 50),
                       "path": "structured triangulation kernels (tri_grid.cuh): no mesh-static index data" if structured else "default kernels of this config",
                       "plan_bytes_per_elem": 0.0 if structured else plan_bytes,
                       "parallelism": ("element blocks; interface rows: " + str(headline_exchange) + " on a high-priority side stream (reduce(vals) overlaps the adjoint "
                                       "kernel, replicate(dK) the forward kernel); %d interface bytes per step on rank 0" % int(ibytes)) if world > 1 else "single GPU"},
            "roofline": roofline, "general_path": general, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches * K, "clocks": sampler.result(),
            "extra": {"configs": extras, "wait_for_exchange_ms": xch_ms, "exchange_alone_ms": xa_ms}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
