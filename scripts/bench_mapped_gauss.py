#!/usr/bin/env python
"""Gauss-point scatter operators and the Laplace term on a mapped + jittered structured grid (Mesh(n,n,1/n) connectivity, P1): the index-free
one-thread-per-node kernels with positions from the coordinate array (option structured = 1, the default) against the general adjacency-walking
kernels (structured = 0).  One JSON line per operator on stdout.   python scripts/bench_mapped_gauss.py [--n 2048] [--steps 20]"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import adfem_jl_b200 as A
from adfem_jl_b200 import _lib, meshgen
from bench_configs import timed


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    n = args.n
    c, e = meshgen.tri_grid(n, n, 1.0 / n)
    rng = np.random.default_rng(3)
    c = np.stack([c[:, 0] + 0.02 * np.sin(3.0 * c[:, 1]), c[:, 1] + 0.02 * np.cos(2.0 * c[:, 0])], 1) + rng.uniform(-0.2 / n, 0.2 / n, c.shape)
    m = A.Mesh(c, e)
    L = _lib.lib()
    assert L.adfem_mesh_info(m.handle, _lib.INFO_STRUCTURED) == 3
    L.adfem_gauss_op_len.restype = C.c_longlong
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    res = {}
    for on in (1, 0):
        m.set_option("structured", on)
        for kind, name in enumerate(("fem_to_gauss", "dof_to_gauss", "grad", "strain", "strain_energy")):
            nin, nout = L.adfem_gauss_op_len(m.handle, kind, 0), L.adfem_gauss_op_len(m.handle, kind, 1)
            torch.manual_seed(kind)           # the same input for both kernels
            if name == "strain_energy":       # forward is the scatter (Gauss points -> dofs)
                x, y = torch.rand(nin, dtype=torch.float64, device="cuda"), torch.empty(nout, dtype=torch.float64, device="cuda")
                fn = lambda: _lib.check(L.adfem_gauss_op(m.handle, kind, p(x), p(y), st))
            else:                             # adjoint is the scatter
                x, y = torch.rand(nout, dtype=torch.float64, device="cuda"), torch.empty(nin, dtype=torch.float64, device="cuda")
                fn = lambda: _lib.check(L.adfem_gauss_op_adjoint(m.handle, kind, p(x), p(y), st))
            res.setdefault(name + "_scatter", {})[on] = (timed(fn, args.steps), 8 * m.dim * m.nnode + 8 * (nin + nout), y.clone())
        nu = torch.rand(m.ngauss, dtype=torch.float64, device="cuda") + 0.5
        u = torch.rand(m.ndof, dtype=torch.float64, device="cuda")
        out = torch.empty_like(u)
        fn = lambda: _lib.check(L.adfem_laplace_term(m.handle, p(nu), p(u), p(out), st))
        res.setdefault("laplace_term", {})[on] = (timed(fn, args.steps), 8 * m.dim * m.nnode + 8 * m.ngauss + 16 * m.ndof, None)
    for name, r in res.items():
        (t1, b, y1), (t0, _, y0) = r[1], r[0]
        same = None if y1 is None else float((y1 - y0).abs().max() / y0.abs().max())       # FMA contraction may differ between the two kernels
        print(json.dumps({"case": name + "_P1_mapped_grid", "elements": m.nelem, "structured_ms": t1, "general_ms": t0, "speedup": t0 / t1,
                          "alg_bytes_structured": b, "structured_GBps": b / (t1 * 1e-3) / 1e9, "general_plus_connectivity_GBps": (b + 12 * m.nelem) / (t0 * 1e-3) / 1e9,
                          "max_abs_diff_over_max_abs": same}), flush=True)


if __name__ == "__main__":
    main()
