# ncu full capture of the CSR tile kernels with explicit launch options. Usage: bash scripts/gpu_prof2.sh <tag> <bench args...>
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile -s 6 -c 2 -f -o gpurun_out/prof_$TAG python bench.py --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline "$@" > gpurun_out/prof_$TAG.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/prof_$TAG.log | cut -c1-300
