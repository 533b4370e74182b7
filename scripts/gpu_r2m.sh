# Round 2, call m: config-5 variants — adjoint with 16-byte shared-memory accesses, node kernel with the heavy incidence list split over two warps
TAG=${1:-r2m}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_widen_gauss_ops.py tests/test_baseline_size_parity.py tests/test_structured.py -m gpu -q -x --timeout 900 -k "tet or config5 or structured_parity and mapped" > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
for opt in "tet_node=1" "tet_node=2"; do
  timeout 600 python scripts/bench_configs.py --cases 5 --steps 10 --scale 2 --opt $opt > gpurun_out/cfg5_x2_${opt}_$TAG.jsonl 2> gpurun_out/cfg5_x2_${opt}_$TAG.err
  echo "cfg5 x2 $opt rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg5_x2_${opt}_$TAG.jsonl; tail -2 gpurun_out/cfg5_x2_${opt}_$TAG.err
done
timeout 900 python bench.py --config 5 --extra-configs none --no-cpu-baseline --e2e-steps 0 --steps 10 --opt tet_node=2 > gpurun_out/bench_cfg5_node2_$TAG.json 2> gpurun_out/bench_cfg5_node2_$TAG.err
echo "bench cfg5 tet_node=2 rc=$?"; python scripts/bench_line.py cfg5 < gpurun_out/bench_cfg5_node2_$TAG.json; tail -3 gpurun_out/bench_cfg5_node2_$TAG.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tet_node_fwd|k_tet_grid_elast_adj" -s 6 -c 2 -f -o gpurun_out/prof_cfg5_$TAG \
  python scripts/bench_configs.py --cases 5 --steps 1 --scale 2 --opt tet_node=2 > gpurun_out/prof_cfg5_$TAG.log 2>&1
echo "ncu cfg5 rc=$?"
