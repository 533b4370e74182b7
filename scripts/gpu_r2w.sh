# Round 2, call w (2 GPUs): config 4 with the one-barrier forward on / off under torchrun
TAG=${1:-r2w}
mkdir -p gpurun_out
for ov in 0 -1; do
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2957$((ov+2)) bench.py --gpus 2 --steps 20 --warmup 3 \
    --extra-configs 4 --e2e-steps 0 --opt tile_overlap=$ov > gpurun_out/bench_n2_ov${ov}_$TAG.json 2> gpurun_out/bench_n2_ov${ov}_$TAG.err
  echo "N=2 tile_overlap=$ov rc=$?"; python scripts/bench_line.py ov$ov < gpurun_out/bench_n2_ov${ov}_$TAG.json; tail -2 gpurun_out/bench_n2_ov${ov}_$TAG.err | cut -c1-200
done
