# Round 2, 2-GPU call: the in-library interface exchange (csrc/dist.cu) against the torch.distributed path, then configs 5 (weak, strong) and 4 on 2 GPUs
TAG=${1:-r2f}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 scripts/test_dist_gpu.py > gpurun_out/distcheck_$TAG.log 2>&1
echo "dist check rc=$?"; tail -8 gpurun_out/distcheck_$TAG.log
for lib in 1 0; do
  timeout 900 $TR --master-port 2954$lib scripts/bench_dist_configs.py --cases 5g,5 --steps 10 --library $lib > gpurun_out/dist5_lib${lib}_$TAG.jsonl 2> gpurun_out/dist5_lib${lib}_$TAG.err
  echo "dist cfg5 library=$lib rc=$?"; cut -c1-900 gpurun_out/dist5_lib${lib}_$TAG.jsonl; tail -3 gpurun_out/dist5_lib${lib}_$TAG.err
done
timeout 900 $TR --master-port 29551 scripts/bench_dist_configs.py --cases 4l --steps 10 > gpurun_out/dist4_$TAG.jsonl 2> gpurun_out/dist4_$TAG.err
echo "dist cfg4 rc=$?"; cut -c1-900 gpurun_out/dist4_$TAG.jsonl; tail -3 gpurun_out/dist4_$TAG.err
