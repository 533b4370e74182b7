# Round 2, call q: launch-shape sweep of the P2 tile kernels (config 4 at 2 M elements, random numbering) after the one-barrier forward
TAG=${1:-r2q}
mkdir -p gpurun_out
: > gpurun_out/sweep4_$TAG.txt
for th in 128 192 256 320 384; do
  for sb in 73728 106496 151552 204800; do
    timeout 300 python scripts/bench_configs.py --cases 4l --steps 10 --opt tile_threads=$th --opt smem_budget=$sb > gpurun_out/tmp_$TAG.jsonl 2> gpurun_out/tmp_$TAG.err
    echo "threads=$th smem=$sb $(python scripts/cfg_line.py < gpurun_out/tmp_$TAG.jsonl) $(tail -1 gpurun_out/tmp_$TAG.err | cut -c1-120)" | tee -a gpurun_out/sweep4_$TAG.txt
  done
done
