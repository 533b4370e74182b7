# Round 2, call v: config 4 at 2 M elements in Morton element order (what one of 8 ranks holds): one-barrier forward on / off
TAG=${1:-r2v}
mkdir -p gpurun_out
for ov in -1 0; do
  timeout 600 python bench.py --config 4o --scale 0.35361 --extra-configs none --no-cpu-baseline --e2e-steps 0 --steps 30 --opt tile_overlap=$ov > gpurun_out/bench_4o_2M_ov${ov}_$TAG.json 2> gpurun_out/bench_4o_2M_ov${ov}_$TAG.err
  echo "cfg4o 2M tile_overlap=$ov rc=$?"; python scripts/bench_line.py ov$ov < gpurun_out/bench_4o_2M_ov${ov}_$TAG.json; tail -2 gpurun_out/bench_4o_2M_ov${ov}_$TAG.err
done
