# parameter sweep of the tile kernels on config 2
mkdir -p gpurun_out
for B in 30000 38000 46000 53248 72000; do for T in 128 192 256 320; do
timeout 300 python bench.py --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline --tile-threads $T --smem-budget $B 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); k=d['roofline']['kernels']; print('budget $B threads $T', 'fwd ms %.3f adj ms %.3f step %.3f planB/elem %.1f' % (k['fwd']['ms'],k['adj']['ms'],d['ms_per_step'],d['config']['plan_bytes_per_elem']))
    elif 'rror' in l: print(l.strip())
"
done; done 2>&1 | tee gpurun_out/sweep1.txt
