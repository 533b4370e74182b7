TAG=${1:-r2o}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_structured.py tests/test_gpu_aux_ops.py tests/test_gpu_parity.py -m gpu -q --timeout 900 > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
timeout 600 python scripts/bench_configs.py --cases src --steps 20 > gpurun_out/src_$TAG.jsonl 2> gpurun_out/src_$TAG.err
echo "src rc=$?"; cut -c1-400 gpurun_out/src_$TAG.jsonl
