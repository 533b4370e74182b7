# Round 2, call i: Dirichlet kernels (segmented reduce, DirichletBd), config-5 variants (adjoint register caps, node kernel split by parity)
TAG=${1:-r2i}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_aux_ops.py tests/test_baseline_size_parity.py tests/test_widen_gauss_ops.py -m gpu -q -x --timeout 900 -k "dirichlet or config5 or tet" > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
for opt in "tet_adj_blocks=3" "tet_adj_blocks=4" "tet_adj_blocks=5" "tet_split=1"; do
  timeout 600 python scripts/bench_configs.py --cases 5 --steps 10 --scale 2 --opt $opt > gpurun_out/cfg5_x2_${opt}_$TAG.jsonl 2> gpurun_out/cfg5_x2_${opt}_$TAG.err
  echo "cfg5 x2 $opt rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg5_x2_${opt}_$TAG.jsonl; tail -2 gpurun_out/cfg5_x2_${opt}_$TAG.err
done
timeout 600 python scripts/bench_configs.py --cases 5 --steps 10 > gpurun_out/cfg5_$TAG.jsonl 2> gpurun_out/cfg5_$TAG.err
echo "cfg5 rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg5_$TAG.jsonl
timeout 900 python bench.py --config 5 --extra-configs none --no-cpu-baseline --e2e-steps 0 --steps 10 > gpurun_out/bench_cfg5_$TAG.json 2> gpurun_out/bench_cfg5_$TAG.err
echo "bench cfg5 rc=$?"; python scripts/bench_line.py cfg5 < gpurun_out/bench_cfg5_$TAG.json; tail -3 gpurun_out/bench_cfg5_$TAG.err
