# Round 2, call r: P2 adjoint with 3 x 72 KB CTAs per SM (default now) at 2 M and at 16 M elements; parity of the P2 CSR kernels
TAG=${1:-r2r}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_baseline_size_parity.py -m gpu -q -x --timeout 900 -k "csr or config4 or tile_overlap" > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python scripts/bench_configs.py --cases 4l,4m --steps 20 > gpurun_out/cfg4_$TAG.jsonl 2> gpurun_out/cfg4_$TAG.err
echo "cfg4 (2 M) rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg4_$TAG.jsonl
timeout 600 python scripts/bench_configs.py --cases 4l --steps 20 --opt tile_threads=320 > gpurun_out/cfg4_t320_$TAG.jsonl 2> gpurun_out/cfg4_t320_$TAG.err
echo "cfg4 (2 M, 320 threads) rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg4_t320_$TAG.jsonl
timeout 1200 python bench.py --config 4 --extra-configs 4o --no-cpu-baseline --e2e-steps 0 --steps 20 > gpurun_out/bench_cfg4_$TAG.json 2> gpurun_out/bench_cfg4_$TAG.err
echo "bench cfg4 rc=$?"; python scripts/bench_line.py cfg4 < gpurun_out/bench_cfg4_$TAG.json; tail -3 gpurun_out/bench_cfg4_$TAG.err
timeout 1200 python bench.py --config 4 --extra-configs none --no-cpu-baseline --e2e-steps 0 --steps 20 --opt smem_budget_adj=204800 > gpurun_out/bench_cfg4_adj200_$TAG.json 2> gpurun_out/bench_cfg4_adj200_$TAG.err
echo "bench cfg4 adj 200 KB rc=$?"; python scripts/bench_line.py cfg4-adj200 < gpurun_out/bench_cfg4_adj200_$TAG.json; tail -3 gpurun_out/bench_cfg4_adj200_$TAG.err
