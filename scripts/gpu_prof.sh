# ncu captures of the hot kernels (one GPU). Usage: bash scripts/gpu_prof.sh <tag> [size]
TAG=${1:-r1}; SIZE=${2:-2048}
mkdir -p gpurun_out
ARGS="--size $SIZE --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 6 -c 2 -f -o gpurun_out/prof_$TAG python bench.py $ARGS > gpurun_out/prof_$TAG.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/prof_$TAG.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/launches_$TAG.log
tail -3 gpurun_out/prof_$TAG.log; tail -3 gpurun_out/launches_$TAG.log
