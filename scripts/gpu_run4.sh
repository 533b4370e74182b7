# GPU tests, then A/B bench runs. Usage: bash scripts/gpu_run4.sh <tag> "<args1>" "<args2>" ...
TAG=$1; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
bash scripts/gpu_ab.sh "$@"
