mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
(time python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
echo "bench rc=$?" >> gpurun_out/bench1.err
tail -c 3000 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -5; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
