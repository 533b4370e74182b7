# Round 2, call n: P2 tile forward with four gather items in flight per thread (phase B), against the numbers of r2final1 / r2a
TAG=${1:-r2n}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_baseline_size_parity.py -m gpu -q -x --timeout 900 -k "csr or config4" > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 600 python scripts/bench_configs.py --cases 4l,4m --steps 20 > gpurun_out/cfg4_$TAG.jsonl 2> gpurun_out/cfg4_$TAG.err
echo "cfg4 (2 M, random numbering) rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg4_$TAG.jsonl; tail -2 gpurun_out/cfg4_$TAG.err
timeout 900 python bench.py --config 4 --extra-configs 4 --no-cpu-baseline --e2e-steps 0 --steps 20 > gpurun_out/bench_cfg4_$TAG.json 2> gpurun_out/bench_cfg4_$TAG.err
echo "bench cfg4 rc=$?"; python scripts/bench_line.py cfg4 < gpurun_out/bench_cfg4_$TAG.json; tail -3 gpurun_out/bench_cfg4_$TAG.err
