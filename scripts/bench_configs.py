#!/usr/bin/env python
"""Single-GPU timings of BASELINE.json configs 3-5 (and the source term of config 2) at sizes whose host-side symbolic phase
finishes in about a minute.  These are NOT bench.py lines (bench.py measures config 2); they record where the general tile
kernels stand on the other operator / element families.  One JSON line per case on stdout.

  python scripts/bench_configs.py [--steps 20] [--cases 3,3f,3q,4l,4m,5,5s,src,gp] [--opt coef_presum=1]
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

import adfem_jl_b200 as A
from adfem_jl_b200 import _lib, meshgen


def timed(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[steps]) / steps


OPTS = {}


def run_csr(name, mesh, op, c_per_gauss, steps, note):
    L = _lib.lib()
    for k, v in OPTS.items():
        mesh.set_option(k, v)
    ncomp = mesh.dim if op == 2 else 1
    t0 = time.perf_counter()
    rowptr, _ = mesh.csr_pattern(ncomp)
    nnz, G, E = int(rowptr[-1]), mesh.ngauss, mesh.nelem
    gen = torch.Generator(device="cuda").manual_seed(0)
    coef = torch.rand(G * c_per_gauss, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    dK = torch.rand(nnz, dtype=torch.float64, device="cuda", generator=gen) - 0.5
    vals = torch.empty(nnz, dtype=torch.float64, device="cuda")
    grad = torch.empty(G * c_per_gauss, dtype=torch.float64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    h = mesh.handle
    pk, pv, pd, pg = (C.c_void_p(t.data_ptr()) for t in (coef, vals, dK, grad))
    fwd = lambda: _lib.check(L.adfem_assemble_csr(h, op, pk, pv, st))
    adj = lambda: _lib.check(L.adfem_assemble_csr_adjoint(h, op, pd, pg, st))
    fwd(); adj()
    torch.cuda.synchronize()
    setup = time.perf_counter() - t0
    tf, ta = timed(fwd, steps), timed(adj, steps)
    d, dim, g = mesh.elem_ndof, mesh.dim, mesh.gauss_per_elem
    # conn + coords + coefficients + values, per element and direction
    b = 4 * d + 8 * dim * mesh.nnode / E + 8 * c_per_gauss * g + 8 * nnz / E
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    line = {"case": name, "note": note, "elements": E, "ndof": mesh.ndof, "nnz": nnz, "fwd_ms": tf, "adj_ms": ta,
            "Melem_per_s": E / ((tf + ta) * 1e-3) / 1e6, "alg_bytes_per_elem_per_direction": b,
            "fwd_GBps": b * E / (tf * 1e-3) / 1e9, "adj_GBps": b * E / (ta * 1e-3) / 1e9, "peak_GBps": peak,
            "step_frac": 2 * b * E / ((tf + ta) * 1e-3) / 1e9 / peak, "setup_s": round(setup, 1),
            "plan_bytes_per_elem": L.adfem_mesh_info(h, _lib.INFO_PLAN_BYTES) / E,
            "structured_path": bool(L.adfem_mesh_info(h, _lib.INFO_STRUCTURED)) and op != 2, "options": dict(OPTS),
            "tiles": [int(L.adfem_mesh_info(h, _lib.INFO_TILES_FWD)), int(L.adfem_mesh_info(h, _lib.INFO_TILES_ADJ))]}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--cases", default="3,4l,4m,5,src")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every mesh edge count by this factor (smoke runs)")
    ap.add_argument("--smem-budget", type=int, default=0)
    ap.add_argument("--tile-threads", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="adfem_set_option on every mesh, e.g. --opt coef_presum=1")
    args = ap.parse_args()
    for kv in args.opt:
        k, v = kv.split("=")
        OPTS[k] = int(v)
    cases = args.cases.split(",")
    if args.smem_budget:
        OPTS["smem_budget"] = args.smem_budget
    if args.tile_threads:
        OPTS["tile_threads"] = args.tile_threads
    s = args.scale
    torch.cuda.set_device(0)
    if "3" in cases:
        m = A.Mesh(int(4096 * s), int(2048 * s), 1.0 / int(4096 * s))
        run_csr("config3_elasticity_P1_tri", m, 2, 9, args.steps, "Mesh(4096,2048,h) P1, per-Gauss-point 3x3 H, CSR fwd + H-adjoint (general tile kernels)")
        del m
    if "3f" in cases:
        # config 3 with the constitutive pre-step fused: E, nu per Gauss point in, H never materialised
        L = _lib.lib()
        m = A.Mesh(int(4096 * s), int(2048 * s), 1.0 / int(4096 * s))
        for k, v in OPTS.items():
            m.set_option(k, v)
        rowptr, _ = m.csr_pattern(2)
        nnz, G = int(rowptr[-1]), m.ngauss
        Emod = torch.rand(G, dtype=torch.float64, device="cuda") + 0.5
        nu = torch.rand(G, dtype=torch.float64, device="cuda") * 0.4
        vals = torch.empty(nnz, dtype=torch.float64, device="cuda")
        dK = torch.rand(nnz, dtype=torch.float64, device="cuda") - 0.5
        gE, gnu = torch.empty_like(Emod), torch.empty_like(nu)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())
        fwd = lambda: _lib.check(L.adfem_assemble_csr_plane(m.handle, 1, p(Emod), p(nu), p(vals), st))
        adj = lambda: _lib.check(L.adfem_assemble_csr_plane_adjoint(m.handle, 1, p(Emod), p(nu), p(dK), p(gE), p(gnu), st))
        tf, ta = timed(fwd, args.steps), timed(adj, args.steps)
        b = 12 + 8 * 2 * m.nnode / m.nelem + 8 * 2 * m.gauss_per_elem + 8 * nnz / m.nelem      # connectivity + coordinates + (E, nu) + values
        print(json.dumps({"case": "config3_fused_plane_stress", "elements": m.nelem, "nnz": nnz, "fwd_ms": tf, "adj_ms": ta,
                          "Melem_per_s": m.nelem / ((tf + ta) * 1e-3) / 1e6, "alg_bytes_per_elem_per_direction": b,
                          "fwd_GBps": b * m.nelem / (tf * 1e-3) / 1e9, "adj_GBps": b * m.nelem / (ta * 1e-3) / 1e9, "options": dict(OPTS)}), flush=True)
        del m
    if "3q" in cases:
        # config 3, second half: the literal structured compute_fem_stiffness_matrix1 (UnivariateFemStiffness) with SpatialVaryingTangentElastic type 1
        L = _lib.lib()
        mq = int(2896 * s)
        G4 = 4 * mq * mq
        mu = torch.rand(G4, dtype=torch.float64, device="cuda") + 0.5
        hmat = torch.empty((G4, 2, 2), dtype=torch.float64, device="cuda")
        vv = torch.empty(16 * G4, dtype=torch.float64, device="cuda")
        gvv = torch.rand(16 * G4, dtype=torch.float64, device="cuda") - 0.5
        gh, gmu = torch.empty_like(hmat), torch.empty_like(mu)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: C.c_void_p(t.data_ptr())
        h = 1.0 / mq

        def fwd():
            _lib.check(L.adfem_svt(p(mu), C.c_longlong(mq), C.c_longlong(mq), 1, p(hmat), st))
            _lib.check(L.adfem_quad_stiffness1(p(hmat), 1, mq, mq, C.c_double(h), None, None, p(vv), st))

        def adj():
            _lib.check(L.adfem_quad_stiffness1_grad(p(gvv), 1, mq, mq, C.c_double(h), p(gh), st))
            _lib.check(L.adfem_svt_grad(p(gh), C.c_longlong(mq), C.c_longlong(mq), 1, p(gmu), st))

        tf, ta = timed(fwd, args.steps), timed(adj, args.steps)
        b = 8 * (1 + 4 + 16) * 4          # per cell: mu in, hmat out + in, vv out (COO-compatible 64 slots per cell)
        print(json.dumps({"case": "config3_quad_stiffness1_svt", "cells": mq * mq, "fwd_ms": tf, "adj_ms": ta, "bytes_per_cell_per_direction": b,
                          "fwd_GBps": b * mq * mq / (tf * 1e-3) / 1e9, "adj_GBps": b * mq * mq / (ta * 1e-3) / 1e9}), flush=True)
        # the same pair fused: the 4mn x 2 x 2 tensor is never written
        ffwd = lambda: _lib.check(L.adfem_quad_stiffness1_svt(p(mu), 1, mq, mq, C.c_double(h), None, None, p(vv), st))
        fadj = lambda: _lib.check(L.adfem_quad_stiffness1_svt_grad(p(gvv), 1, mq, mq, C.c_double(h), p(gmu), st))
        tf, ta = timed(ffwd, args.steps), timed(fadj, args.steps)
        b = 8 * (1 + 16) * 4
        print(json.dumps({"case": "config3_quad_stiffness1_svt_fused", "cells": mq * mq, "fwd_ms": tf, "adj_ms": ta, "bytes_per_cell_per_direction": b,
                          "fwd_GBps": b * mq * mq / (tf * 1e-3) / 1e9, "adj_GBps": b * mq * mq / (ta * 1e-3) / 1e9}), flush=True)
    if "4l" in cases or "4m" in cases:
        n = int(1000 * s)
        c, e = meshgen.jitter_unstructured(n, n, 1.0 / n, seed=2)
        m = A.Mesh(c, e, degree=2)
        if "4l" in cases:
            run_csr("config4_laplace_P2_unstructured", m, 0, 1, args.steps, "jittered, randomly renumbered triangulation, P2 (d=6, g=6), %dx%d cells" % (n, n))
        if "4m" in cases:
            run_csr("config4_mass_P2_unstructured", m, 1, 1, args.steps, "same mesh, mass matrix")
        del m
    if "5" in cases:
        n = int(64 * s)
        c, e = meshgen.tet_grid(n, n, n, 1.0 / n)
        m = A.Mesh3(c, e)
        run_csr("config5_elasticity_P1_tet", m, 2, 36, args.steps, "Mesh3(%d,%d,%d,h) P1 tets, per-Gauss-point 6x6 Voigt H (3-D extension N2)" % (n, n, n))
        del m
    if "5s" in cases:
        # the reference's own 3-D op (FemLaplaceScalarT) on the same tetrahedral grid
        n = int(64 * s)
        c, e = meshgen.tet_grid(n, n, n, 1.0 / n)
        m = A.Mesh3(c, e)
        run_csr("config5_laplace_P1_tet", m, 0, 1, args.steps, "Mesh3(%d,%d,%d,h) P1 tets, scalar Laplace (FemLaplaceScalarT)" % (n, n, n))
        del m
    if "src" in cases:
        L = _lib.lib()
        n = int(4096 * s)
        m = A.Mesh(n, n, 1.0 / n)
        G = m.ngauss
        f = torch.rand(G, dtype=torch.float64, device="cuda")
        rhs = torch.empty(m.ndof, dtype=torch.float64, device="cuda")
        gr = torch.rand(m.ndof, dtype=torch.float64, device="cuda")
        gf = torch.empty(G, dtype=torch.float64, device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        fwd = lambda: _lib.check(L.adfem_source(m.handle, C.c_void_p(f.data_ptr()), C.c_void_p(rhs.data_ptr()), st))
        adj = lambda: _lib.check(L.adfem_source_adjoint(m.handle, C.c_void_p(gr.data_ptr()), C.c_void_p(gf.data_ptr()), st))
        tf, ta = timed(fwd, args.steps), timed(adj, args.steps)
        b = 12 + 8 * 2 * m.nnode / m.nelem + 24 + 8 * m.ndof / m.nelem
        print(json.dumps({"case": "config2_source_term_P1_tri", "elements": m.nelem, "fwd_ms": tf, "adj_ms": ta, "alg_bytes_per_elem_per_direction": b,
                          "fwd_GBps": b * m.nelem / (tf * 1e-3) / 1e9, "adj_GBps": b * m.nelem / (ta * 1e-3) / 1e9}), flush=True)
    if "gp" in cases:
        # Gauss-point operators on the config-4-like unstructured P2 mesh and on Mesh(2048,2048) P1 (general kernels)
        L = _lib.lib()
        L.adfem_gauss_op_len.restype = C.c_longlong
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        n = int(1000 * s)
        c, e = meshgen.jitter_unstructured(n, n, 1.0 / n, seed=2)
        for tag, m in (("P2_unstructured", A.Mesh(c, e, degree=2)), ("P1_grid", A.Mesh(int(2048 * s), int(2048 * s), 1.0 / int(2048 * s)))):
            for k, v in OPTS.items():
                m.set_option(k, v)
            for kind, name in enumerate(("fem_to_gauss", "dof_to_gauss", "grad", "strain", "strain_energy")):
                nin, nout = L.adfem_gauss_op_len(m.handle, kind, 0), L.adfem_gauss_op_len(m.handle, kind, 1)
                x = torch.rand(nin, dtype=torch.float64, device="cuda")
                y = torch.empty(nout, dtype=torch.float64, device="cuda")
                w = torch.rand(nout, dtype=torch.float64, device="cuda")
                gx = torch.empty(nin, dtype=torch.float64, device="cuda")
                fwd = lambda: _lib.check(L.adfem_gauss_op(m.handle, kind, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), st))
                adj = lambda: _lib.check(L.adfem_gauss_op_adjoint(m.handle, kind, C.c_void_p(w.data_ptr()), C.c_void_p(gx.data_ptr()), st))
                tf, ta = timed(fwd, args.steps), timed(adj, args.steps)
                b = 4 * m.elem_ndof + 8 * m.dim * m.nnode / m.nelem + 8 * (nin + nout) / m.nelem      # connectivity + coordinates + vectors
                print(json.dumps({"case": "gauss_op_%s_%s" % (name, tag), "elements": m.nelem, "fwd_ms": tf, "adj_ms": ta, "alg_bytes_per_elem_per_direction": b,
                                  "fwd_GBps": b * m.nelem / (tf * 1e-3) / 1e9, "adj_GBps": b * m.nelem / (ta * 1e-3) / 1e9}), flush=True)
            nu = torch.rand(m.ngauss, dtype=torch.float64, device="cuda") + 0.5
            u = torch.rand(m.ndof, dtype=torch.float64, device="cuda")
            go = torch.rand(m.ndof, dtype=torch.float64, device="cuda")
            out, gu, gnu = torch.empty_like(u), torch.empty_like(u), torch.empty_like(nu)
            p = lambda t: C.c_void_p(t.data_ptr())
            fwd = lambda: _lib.check(L.adfem_laplace_term(m.handle, p(nu), p(u), p(out), st))
            adj = lambda: _lib.check(L.adfem_laplace_term_adjoint(m.handle, p(nu), p(u), p(go), p(gnu), p(gu), st))
            tf, ta = timed(fwd, args.steps), timed(adj, args.steps)
            b = 4 * m.elem_ndof + 8 * m.dim * m.nnode / m.nelem + 8 * m.gauss_per_elem + 16 * m.ndof / m.nelem
            print(json.dumps({"case": "laplace_term_%s" % tag, "elements": m.nelem, "fwd_ms": tf, "adj_ms": ta, "alg_bytes_per_elem_fwd": b,
                              "fwd_GBps": b * m.nelem / (tf * 1e-3) / 1e9}), flush=True)
            del m


if __name__ == "__main__":
    main()
