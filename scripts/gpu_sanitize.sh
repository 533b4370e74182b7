# compute-sanitizer passes over the parity tests of the CSR kernels (small meshes): memcheck + racecheck (shared-memory hazards)
mkdir -p gpurun_out
SEL='test_csr_scalar_ops or test_csr_stiffness or test_structured_parity or test_structured_source_term or test_structured_host_buffer_pipeline'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_structured.py -m gpu -q -x --timeout 1400 -k "$SEL" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck.log | tail -3
timeout 2400 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_structured.py -m gpu -q -x --timeout 2300 -k "test_csr_scalar_ops and tri_unstruct or test_csr_stiffness and tri_struct or test_structured_parity and 63 or test_structured_source_term and 63" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.log | sort | uniq -c | tail -8
