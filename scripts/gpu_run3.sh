# full GPU validation + bench + ncu launch list + full profiles of the hot kernels. Usage: bash scripts/gpu_run3.sh <tag>
TAG=${1:-r1e}
mkdir -p gpurun_out
(time python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; cat gpurun_out/bench_ref_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_grid -s 6 -c 2 -f -o gpurun_out/prof_grid_$TAG python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --general-steps 0 > gpurun_out/prof_grid_$TAG.log 2>&1
echo "ncu full (structured kernels) rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 6 -c 2 -f -o gpurun_out/prof_tile_$TAG python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --structured 0 > gpurun_out/prof_tile_$TAG.log 2>&1
echo "ncu full (general tile kernels) rc=$?"
