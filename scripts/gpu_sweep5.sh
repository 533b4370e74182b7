# tests first, then a sweep of CTA size / tile size for the v2 tile kernels on config 2
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_sw5.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_sw5.log
run() { timeout 300 python bench.py --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernels']; print('$*', 'fwd ms %.3f adj ms %.3f step %.3f planB/elem %.1f' % (k['fwd']['ms'],k['adj']['ms'],d['ms_per_step'],d['config']['plan_bytes_per_elem']))
    elif 'rror' in l: print(l.strip())
"; }
(run
for T in 192 256 320 384; do for E in 1 2; do run --tile-threads $T --elems-per-tile $((T*E)) --rows-per-tile $((T*4/5)); done; done
run --tile-threads 256 --elems-per-tile 256 --rows-per-tile 160
run --tile-threads 256 --elems-per-tile 512 --rows-per-tile 128
run --tile-threads 320 --elems-per-tile 320 --rows-per-tile 200
run --tile-threads 320 --elems-per-tile 640 --rows-per-tile 272
) 2>&1 | tee gpurun_out/sweep5.txt
