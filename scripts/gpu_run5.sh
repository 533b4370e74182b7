# structured-path tests, then A/B bench runs. Usage: bash scripts/gpu_run5.sh <tag> "<args1>" ...
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_structured.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
bash scripts/gpu_ab.sh "$@"
