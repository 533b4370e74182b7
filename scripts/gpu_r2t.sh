TAG=${1:-r2t}
mkdir -p gpurun_out
for sb in 106496 151552 204800; do
  timeout 1200 python bench.py --config 4o --extra-configs none --no-cpu-baseline --e2e-steps 0 --steps 20 --opt smem_budget_adj=$sb > gpurun_out/bench_cfg4o_adj${sb}_$TAG.json 2> gpurun_out/bench_cfg4o_adj${sb}_$TAG.err
  echo "bench cfg4o adj $sb rc=$?"; python scripts/bench_line.py cfg4o-adj$sb < gpurun_out/bench_cfg4o_adj${sb}_$TAG.json; tail -2 gpurun_out/bench_cfg4o_adj${sb}_$TAG.err
done
