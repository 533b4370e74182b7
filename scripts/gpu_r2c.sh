# config 5 forward, second generation (tet_node.cuh): parity tests, timing at two sizes, ncu of the two new kernels
TAG=${1:-r2c}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_widen_gauss_ops.py tests/test_exact_goldens.py -m gpu -q -k "tet or exact" --timeout 600 > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
timeout 600 python scripts/bench_configs.py --cases 5 --steps 10 > gpurun_out/cfg5_$TAG.jsonl 2> gpurun_out/cfg5_$TAG.err
echo "cfg5 rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg5_$TAG.jsonl
timeout 600 python scripts/bench_configs.py --cases 5 --steps 10 --opt tet_node=0 > gpurun_out/cfg5_old_$TAG.jsonl 2> gpurun_out/cfg5_old_$TAG.err
echo "cfg5 (old fwd) rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg5_old_$TAG.jsonl
timeout 600 python scripts/bench_configs.py --cases 5 --steps 5 --scale 2 > gpurun_out/cfg5_x2_$TAG.jsonl 2> gpurun_out/cfg5_x2_$TAG.err
echo "cfg5 x2 rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg5_x2_$TAG.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tet_presum_x|k_tet_node_fwd" -s 6 -c 2 -f -o gpurun_out/prof_cfg5_$TAG \
  python scripts/bench_configs.py --cases 5 --steps 1 > gpurun_out/prof_cfg5_$TAG.log 2>&1
echo "ncu cfg5 rc=$?"
