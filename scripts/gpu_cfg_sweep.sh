# tile-size sweep of the general kernels on configs 3-5
mkdir -p gpurun_out
for B in 73728 112000 200000; do for T in 256 512; do
  timeout 300 python scripts/bench_configs.py --cases 3,4l,5 --steps 10 --smem-budget $B --tile-threads $T 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('budget $B threads $T', d['case'], 'fwd %.3f adj %.3f ms frac %.3f tiles %s planB %.0f' % (d['fwd_ms'], d['adj_ms'], d['step_frac'], d['tiles'], d['plan_bytes_per_elem']))
    elif 'rror' in l: print(l.strip())
"
done; done 2>&1 | tee gpurun_out/cfg_sweep.txt
