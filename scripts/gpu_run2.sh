# tests + bench + profile in one call. Usage: bash scripts/gpu_run2.sh <tag>
TAG=${1:-r1b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -15 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
for T in 128 320 512; do
timeout 600 python bench.py --steps 30 --warmup 3 --e2e-steps 0 --no-cpu-baseline --tile-threads $T 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernels']; print('threads $T', 'fwd ms',k['fwd']['ms'],'adj ms',k['adj']['ms'],'step',d['ms_per_step'])
    else: print(l.strip())
"
done
ARGS="--size 2048 --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 6 -c 2 -f -o gpurun_out/prof_$TAG python bench.py $ARGS > gpurun_out/prof_$TAG.log 2>&1
echo "ncu full rc=$?"
