# Round 2, call g: full GPU test suite, config-5 forward variants (z-chunk pipeline on two streams, 16-byte pre-sum loads, adjoint store phase),
# the new bench.py line with extra.configs at BASELINE sizes, config 4 with generator numbering
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
for ch in 1 4 8 16; do
  timeout 600 python scripts/bench_configs.py --cases 5 --steps 10 --scale 2 --opt tet_chunks=$ch > gpurun_out/cfg5_x2_ch${ch}_$TAG.jsonl 2> gpurun_out/cfg5_x2_ch${ch}_$TAG.err
  echo "cfg5 x2 chunks=$ch rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg5_x2_ch${ch}_$TAG.jsonl
done
timeout 600 python scripts/bench_configs.py --cases 5 --steps 10 > gpurun_out/cfg5_$TAG.jsonl 2> gpurun_out/cfg5_$TAG.err
echo "cfg5 rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg5_$TAG.jsonl
timeout 1500 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; python scripts/bench_line.py < gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 900 python bench.py --config 4l --numbering generator --extra-configs none --no-cpu-baseline --e2e-steps 0 --steps 20 > gpurun_out/bench_cfg4gen_$TAG.json 2> gpurun_out/bench_cfg4gen_$TAG.err
echo "bench cfg4 generator numbering rc=$?"; python scripts/bench_line.py < gpurun_out/bench_cfg4gen_$TAG.json; tail -3 gpurun_out/bench_cfg4gen_$TAG.err
