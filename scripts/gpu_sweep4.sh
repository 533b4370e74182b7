# tests first, then a sweep of the v2 (head/body, software-pipelined) tile kernels on config 2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_sw4.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_sw4.log
run() { timeout 300 python bench.py --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernels']; print('$*', 'fwd ms %.3f adj ms %.3f step %.3f planB/elem %.1f' % (k['fwd']['ms'],k['adj']['ms'],d['ms_per_step'],d['config']['plan_bytes_per_elem']))
    elif 'rror' in l: print(l.strip())
"; }
(run
run --pipeline 0
run --coef-prefetch 0
for B in 49152 57344 73728 110000; do for T in 256 320 384; do run --smem-budget $B --tile-threads $T; done; done) 2>&1 | tee gpurun_out/sweep4.txt
