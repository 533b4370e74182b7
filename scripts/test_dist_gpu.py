#!/usr/bin/env python
"""Multi-GPU check of the in-library interface exchange (csrc/dist.cu: pack kernel -> ncclSend/ncclRecv group -> deterministic unpack) against
the torch.distributed reference path of adfem.jl_b200/dist.py (whose lists are checked against the oracle by the gloo tests on CPU).

  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/test_dist_gpu.py
Prints one line per case from rank 0 and exits non-zero on any mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

from adfem_jl_b200 import dist as adist
from adfem_jl_b200 import meshgen


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [("tri_P1", meshgen.jitter_unstructured(60, 45, 0.02, seed=4, permute=False), 1),
             ("tri_P2_scattered", meshgen.jitter_unstructured(24, 24, 0.04, seed=5, permute=True), 2),
             ("tet_P1", meshgen.tet_grid(8, 8, 8, 0.125), 1)]
    comm = adist.make_nccl_comm(rank, world)
    ok = True
    for name, (coords, elems), degree in cases:
        part, _ = adist.partition_elements(coords, elems, rank, world, degree=degree)
        nnz = int(part.rowptr[-1])
        for nc in ([1, part.mesh.dim] if degree == 1 else [1]):
            gen = torch.Generator(device="cuda").manual_seed(100 + rank)
            vals = torch.rand(nc * nc * nnz, dtype=torch.float64, device="cuda", generator=gen)
            dv = torch.rand(nc * nc * nnz, dtype=torch.float64, device="cuda", generator=gen)
            dg = torch.rand(len(part.ghost_idx) * nc * nc, dtype=torch.float64, device="cuda", generator=gen)
            part._dist = None
            v1 = vals.clone(); part.reduce_interface(v1, ncomp=nc); g1 = part.ghost_vals.clone()
            d1 = dv.clone(); part.replicate_interface(d1, dg, ncomp=nc)
            part.use_library(comm)
            v2 = vals.clone(); part.reduce_interface(v2, ncomp=nc); g2 = part.ghost_vals.clone()
            v3 = vals.clone(); part.reduce_interface(v3, ncomp=nc)
            d2 = dv.clone(); part.replicate_interface(d2, dg, ncomp=nc)
            torch.cuda.synchronize()
            good = torch.allclose(v1, v2, rtol=1e-14, atol=0) and torch.equal(g1, g2) and torch.equal(d1, d2) and torch.equal(v2, v3)
            flag = torch.tensor([int(good), int(torch.equal(v1, v2))], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            sent = torch.tensor([sum(part.send_counts) * nc * nc], device="cuda")
            dist.all_reduce(sent)
            if rank == 0:
                print("%s ncomp=%d world=%d: entries exchanged %d, library == torch path %s (bit-identical %s), run-to-run identical" %
                      (name, nc, world, int(sent.item()), bool(flag[0].item()), bool(flag[1].item())), flush=True)
            ok = ok and bool(flag[0].item())
        del part
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
