# Round 2, call p: scalar tile forward with one barrier per tile (option tile_overlap): parity, then config 2 (general kernels) and config 4 A/B
TAG=${1:-r2p}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -k "tile_overlap or csr_scalar" > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
for ov in 1 0; do
  timeout 1200 python bench.py --no-cpu-baseline --e2e-steps 0 --extra-configs 4,4o --steps 30 --opt tile_overlap=$ov > gpurun_out/bench_ov${ov}_$TAG.json 2> gpurun_out/bench_ov${ov}_$TAG.err
  echo "bench tile_overlap=$ov rc=$?"; python scripts/bench_line.py ov=$ov < gpurun_out/bench_ov${ov}_$TAG.json; tail -3 gpurun_out/bench_ov${ov}_$TAG.err
done
