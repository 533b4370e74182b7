"""CPU tests of bench.py's command-line contract: the reference arm prints one complete JSON line on rank 0 only, and the product arm
refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_line():
    out = _run(["--impl", "reference", "--size", "48", "--steps", "2", "--warmup", "1", "--gpus", "1"])
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Melem/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["steps"] == 2 and d["warmup"] == 1 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "Mesh(48,48,1/48)" in cb["sample"]
    assert d["config"]["same_mesh_as_gpu_arm"] is True
    assert d["e2e"] == {"value": d["value"], "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    out = _run(["--impl", "reference", "--size", "16", "--steps", "1", "--warmup", "0", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_bounds_its_sample():
    """a large step count shrinks the sample to a row slab of the same mesh (stated in the line) instead of running for hours"""
    out = _run(["--impl", "reference", "--size", "256", "--steps", "3", "--warmup", "0", "--cpu-budget-s", "0.05"])
    assert out.returncode == 0, out.stderr
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert d["config"]["same_mesh_as_gpu_arm"] is False and "row slab" in d["cpu_baseline"]["sample"]


def test_bench_cases_host_logic():
    """bench.py's case builder on host-only meshes: element counts of the BASELINE configs scale as stated and SURVEY 8(d)'s
    algorithmic bytes per element come out at 72 / 348 / 264 / ~1363 B."""
    sys.path.insert(0, ROOT)
    import bench
    want = {"2": (72, 1), "3": (348, 9), "4l": (264, 1), "5": (1363, 36)}
    for case, (b, cpg) in want.items():
        mesh, part, op, c, note, scaling = bench.build_case(case, 0, 1, scale=0.01 if case != "5" else 0.06, size=40, host_only=True)
        assert part is None and c == cpg and ("config " + case[0]) in note
        nc = mesh.dim if op == 2 else 1
        rowptr, _ = mesh.csr_pattern(1)
        got = bench.alg_bytes_per_elem(mesh, nc * nc * int(rowptr[-1]), cpg)
        assert abs(got - b) / b < 0.12, (case, got)     # small meshes: boundary effects in nnz/E and nnode/E
        assert scaling == ("strong" if case in ("4l", "5") else "weak")
    # every other case name of the default run builds too (mapped grids, mass matrix, Morton-ordered elements)
    for case, nelem in (("2m", 2 * 40 * 40), ("3m", None), ("4m", None), ("4o", None), ("4", None)):
        mesh, part, op, c, note, scaling = bench.build_case(case, 0, 1, scale=0.01, size=40 if case == "2m" else None, host_only=True)
        assert part is None and mesh.nelem > 0 and (nelem is None or mesh.nelem == nelem)


def test_product_arm_needs_a_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = _run(["--steps", "1", "--warmup", "1", "--size", "8"])
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_multi_gpu_config_script_dry_run():
    """scripts/bench_dist_configs.py --dry-run at world size 2 (gloo, host-only meshes): partitions of configs 3, 4 and 5 are built and both
    interface exchanges run (block operator included); no kernel is launched and no rate is reported."""
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "scripts", "bench_dist_configs.py"), "--dry-run", "--scale", "0.02", "--steps", "1", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert [d["case"] for d in lines] == ["config3", "config4l", "config5"]
    for d in lines:
        assert d["n_gpus"] == 2 and d["dry_run"] is True and d["Melem_per_s"] is None and d["interface_bytes_per_step_total"] > 0
    # without --dry-run and without a GPU the script refuses to run
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "bench_dist_configs.py"), "--cases", "3", "--scale", "0.01"], capture_output=True, text=True, timeout=300)
    import torch
    if not torch.cuda.is_available():
        assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)
