# Round 2, validation pass: smoke, the whole GPU suite, the bench line (all configs), the reference arm, the ncu launch list of a short bench run
TAG=${1:-r2final}
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 2400 python -m pytest tests -m gpu -q --timeout 1200 > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout 1700 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; python scripts/bench_line.py < gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2> gpurun_out/bench_reference_$TAG.err
echo "reference arm rc=$?"; cut -c1-400 gpurun_out/bench_reference_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 3 --warmup 3 --general-steps 2 --extra-steps 2 --e2e-steps 0 --no-cpu-baseline --extra-configs 2m,3,5 --scale 0.5 > gpurun_out/launches_bench_$TAG.log 2>&1
echo "ncu launch list rc=$?"; wc -l gpurun_out/launches_$TAG.csv
