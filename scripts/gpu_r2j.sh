# Round 2, multi-GPU call: bench.py exactly as the driver launches it (torchrun, N ranks), headline config 2 + extra.configs (3 weak, 4 and 5 strong)
# usage: gpurun --gpus N -- 'bash scripts/gpu_r2j.sh N TAG [extra-configs]'
N=${1:-2}; TAG=${2:-r2j}; XC=${3:-3,4,5}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n${N}_$TAG.txt 2>&1; nproc; free -g | head -2
timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 50 --warmup 5 \
  --extra-configs $XC > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
echo "bench N=$N rc=$?"; python scripts/bench_line.py N=$N < gpurun_out/bench_n${N}_$TAG.json; tail -4 gpurun_out/bench_n${N}_$TAG.err
