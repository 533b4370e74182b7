# Round 2, last call (4 GPU-minutes left): the mapped-grid Gauss-point operators first, then the whole GPU suite, then smoke
TAG=${1:-r2z}
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_widen_gauss_ops.py -m gpu -q -k "structured_scatter_operators or gauss_ops_2d or laplace_term" > gpurun_out/pytest_mapped_$TAG.log 2>&1
echo "mapped gauss ops rc=$?"; tail -3 gpurun_out/pytest_mapped_$TAG.log
timeout 170 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout 60 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
