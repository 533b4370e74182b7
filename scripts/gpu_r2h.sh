# Round 2, call h: BASELINE-size parity tests, config 5 after the adjoint rasterisation (three sizes), ncu of the config-5 and config-4 kernels
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_baseline_size_parity.py tests/test_widen_gauss_ops.py -m gpu -q -x --timeout 900 -k "config or tet" > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
for sc in 1 2; do
  timeout 600 python scripts/bench_configs.py --cases 5 --steps 10 --scale $sc > gpurun_out/cfg5_x${sc}_$TAG.jsonl 2> gpurun_out/cfg5_x${sc}_$TAG.err
  echo "cfg5 x$sc rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg5_x${sc}_$TAG.jsonl
done
timeout 900 python bench.py --config 5 --extra-configs none --no-cpu-baseline --e2e-steps 0 --steps 10 > gpurun_out/bench_cfg5_$TAG.json 2> gpurun_out/bench_cfg5_$TAG.err
echo "bench cfg5 rc=$?"; python scripts/bench_line.py cfg5 < gpurun_out/bench_cfg5_$TAG.json; tail -3 gpurun_out/bench_cfg5_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tet_presum_xg|k_tet_node_fwd|k_tet_grid_elast_adj" -s 9 -c 3 -f -o gpurun_out/prof_cfg5_$TAG \
  python scripts/bench_configs.py --cases 5 --steps 1 --scale 2 > gpurun_out/prof_cfg5_$TAG.log 2>&1
echo "ncu cfg5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tile_fwd|k_tile_adj" -s 4 -c 2 -f -o gpurun_out/prof_cfg4_$TAG \
  python scripts/bench_configs.py --cases 4l --steps 1 > gpurun_out/prof_cfg4_$TAG.log 2>&1
echo "ncu cfg4 rc=$?"
ls -la gpurun_out/*.ncu-rep
