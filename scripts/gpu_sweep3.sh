# tests first, then a sweep of the software-pipelined kernels (pipeline 3) against the CTA-per-tile default on config 2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_sw3.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_sw3.log
run() { timeout 300 python bench.py --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernels']; print('$*', 'fwd ms %.3f adj ms %.3f step %.3f planB/elem %.1f' % (k['fwd']['ms'],k['adj']['ms'],d['ms_per_step'],d['config']['plan_bytes_per_elem']))
    elif 'rror' in l: print(l.strip())
"; }
(run --pipeline 0
for B in 24000 32000 40000 52000; do for T in 256 320 384 512; do run --pipeline 3 --smem-budget $B --tile-threads $T; done; done) 2>&1 | tee gpurun_out/sweep3.txt
