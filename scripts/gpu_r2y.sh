# Round 2, call y: MAPPED structured elasticity kernels — GPU parity, then config 3 connectivity on mapped node positions at full size
TAG=${1:-r2y}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_widen_gauss_ops.py tests/test_baseline_size_parity.py -m gpu -q -x --timeout 400 -k "structured_elasticity_kernels or config3 or fused_plane" > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 400 python bench.py --config 3m --extra-configs 3 --no-cpu-baseline --e2e-steps 0 --steps 20 > gpurun_out/bench_cfg3m_$TAG.json 2> gpurun_out/bench_cfg3m_$TAG.err
echo "bench cfg3m rc=$?"; python scripts/bench_line.py cfg3m < gpurun_out/bench_cfg3m_$TAG.json; tail -3 gpurun_out/bench_cfg3m_$TAG.err
