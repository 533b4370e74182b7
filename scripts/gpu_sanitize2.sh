# Round 2: compute-sanitizer over the kernels written or changed in round 2 (small meshes): structured elasticity (triangles, tetrahedra: pre-sum,
# node kernel, z-chunk pipeline on two streams, adjoint), Gauss-point operators, Dirichlet segmented reductions, DirichletBd
TAG=${1:-r2s}
mkdir -p gpurun_out
SEL='test_structured_tet_elasticity_forward or test_structured_elasticity_kernels or test_impose_dirichlet or test_dirichlet_bd or test_gauss_ops_2d or test_laplace_term or test_fused_plane_stiffness'
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_widen_gauss_ops.py tests/test_gpu_aux_ops.py -m gpu -q --timeout 1600 -k "$SEL" > gpurun_out/sanitize_memcheck_$TAG.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck_$TAG.log | tail -3
SELR='test_structured_tet_elasticity_forward and (5-6-3 or 3-2-1 or 4-4-1) or test_structured_elasticity_kernels or test_impose_dirichlet_coo or test_dirichlet_bd'
timeout 2400 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 7 python -m pytest tests/test_widen_gauss_ops.py tests/test_gpu_aux_ops.py -m gpu -q --timeout 2300 -k "$SELR" > gpurun_out/sanitize_racecheck_$TAG.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck_$TAG.log | sort | uniq -c | tail -8
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_widen_gauss_ops.py -m gpu -q --timeout 800 -k "test_structured_tet_elasticity_forward and (5-6-3 or 3-2-1)" > gpurun_out/sanitize_synccheck_$TAG.log 2>&1
echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_synccheck_$TAG.log | tail -3
