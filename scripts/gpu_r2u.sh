# Round 2, call u: P1 tile forward with one barrier per tile at larger shared-memory budgets (the tile size of the two-barrier kernel restored)
TAG=${1:-r2u}
mkdir -p gpurun_out
for sb in 98304 114688; do
  timeout 900 python bench.py --no-cpu-baseline --e2e-steps 0 --extra-configs none --steps 30 --opt tile_overlap=1 --opt smem_budget=$sb --opt smem_budget_adj=73728 > gpurun_out/bench_p1ov_${sb}_$TAG.json 2> gpurun_out/bench_p1ov_${sb}_$TAG.err
  echo "bench P1 overlap smem=$sb rc=$?"; python scripts/bench_line.py p1ov-$sb < gpurun_out/bench_p1ov_${sb}_$TAG.json; tail -2 gpurun_out/bench_p1ov_${sb}_$TAG.err
done
timeout 900 python bench.py --no-cpu-baseline --e2e-steps 0 --extra-configs none --steps 30 --opt tile_overlap=0 --opt smem_budget=98304 --opt smem_budget_adj=73728 > gpurun_out/bench_p1_98304_$TAG.json 2> gpurun_out/bench_p1_98304_$TAG.err
echo "bench P1 two barriers smem=98304 rc=$?"; python scripts/bench_line.py p1-98304 < gpurun_out/bench_p1_98304_$TAG.json
