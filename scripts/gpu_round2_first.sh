# First GPU call of round 2: validates everything that was written in the GPU-less tail of round 1 and takes the A/B numbers those changes
# were written for.  Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_round2_first.sh r2a'.  Everything lands in gpurun_out/.
TAG=${1:-r2a}
mkdir -p gpurun_out
(time python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
# core parity first (rows a-e), then the widening rows; --timeout per test, no -x so that one failing new test does not hide the others
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest (verified suite) rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
# the tests written without GPU time: every test runs (no -x), short tracebacks
ADFEM_RUN_UNVERIFIED=1 timeout 1500 python -m pytest tests/test_widen_gauss_ops.py -m gpu -q --timeout 600 --tb=short > gpurun_out/pytest_unverified_$TAG.log 2>&1
echo "pytest (unverified rows) rc=$?"; tail -40 gpurun_out/pytest_unverified_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; python scripts/bench_line.py $TAG < gpurun_out/bench_$TAG.json
# configs 3 / 5 with and without the Gauss-summed coefficients (option coef_presum), then the Gauss-point operators
for P in 0 1; do
  timeout 600 python scripts/bench_configs.py --cases 3,5 --steps 10 --opt coef_presum=$P > gpurun_out/cfg35_presum${P}_$TAG.jsonl 2> gpurun_out/cfg35_presum${P}_$TAG.err
  echo "configs 3,5 coef_presum=$P rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg35_presum${P}_$TAG.jsonl
done
timeout 600 python scripts/bench_configs.py --cases 3 --steps 10 --opt structured_elasticity=1 > gpurun_out/cfg3_gridelast_$TAG.jsonl 2> gpurun_out/cfg3_gridelast_$TAG.err
echo "config 3 structured elasticity kernels rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg3_gridelast_$TAG.jsonl
for R in 0 1; do
  timeout 600 python scripts/bench_configs.py --cases 4l,4m --steps 10 --opt row_gather=$R > gpurun_out/cfg4_rowgather${R}_$TAG.jsonl 2> gpurun_out/cfg4_rowgather${R}_$TAG.err
  echo "config 4 row_gather=$R rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg4_rowgather${R}_$TAG.jsonl
done
timeout 600 python scripts/bench_configs.py --cases 4l --steps 10 --opt row_gather=1 --opt adjoint_tiled=0 > gpurun_out/cfg4_rowgather_adjgather_$TAG.jsonl 2> gpurun_out/cfg4_rowgather_adjgather_$TAG.err
echo "config 4 row_gather=1 + direct-gather adjoint rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg4_rowgather_adjgather_$TAG.jsonl
timeout 600 python scripts/bench_configs.py --cases 3,5 --steps 10 --opt row_gather=1 > gpurun_out/cfg35_rowgather_$TAG.jsonl 2> gpurun_out/cfg35_rowgather_$TAG.err
echo "configs 3,5 row_gather=1 (plan-free forward) rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg35_rowgather_$TAG.jsonl
for S in 0 1; do
  timeout 600 python scripts/bench_configs.py --cases 5s --steps 10 --opt structured_elasticity=$S > gpurun_out/cfg5s_struct${S}_$TAG.jsonl 2> gpurun_out/cfg5s_struct${S}_$TAG.err
  echo "3-D scalar Laplace on Mesh3, structured_elasticity=$S rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg5s_struct${S}_$TAG.jsonl
done
timeout 600 python scripts/bench_configs.py --cases 5 --steps 10 --opt structured_elasticity=1 > gpurun_out/cfg5_tetgrid_$TAG.jsonl 2> gpurun_out/cfg5_tetgrid_$TAG.err
echo "config 5 structured tetrahedral forward rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg5_tetgrid_$TAG.jsonl
timeout 600 python scripts/bench_configs.py --cases 3f --steps 10 --opt structured_elasticity=1 > gpurun_out/cfg3f_gridelast_$TAG.jsonl 2> gpurun_out/cfg3f_gridelast_$TAG.err
echo "config 3 fused moduli, structured kernels rc=$?"; python scripts/cfg_line.py < gpurun_out/cfg3f_gridelast_$TAG.jsonl
timeout 600 python scripts/bench_configs.py --cases 3f,3q,gp --steps 10 > gpurun_out/gp_$TAG.jsonl 2> gpurun_out/gp_$TAG.err
echo "gauss-point ops rc=$?"; cut -c1-260 gpurun_out/gp_$TAG.jsonl
timeout 600 python scripts/bench_configs.py --cases gp --steps 10 --opt structured=0 > gpurun_out/gp_general_$TAG.jsonl 2> gpurun_out/gp_general_$TAG.err
echo "gauss-point ops, general kernels on the structured mesh rc=$?"; grep P1_grid gpurun_out/gp_general_$TAG.jsonl | cut -c1-260
# one full capture of the new kernels (scatter / Laplace term are the ones expected to need work)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gp_scatter|k_laplace_term|k_presum|k_expand" -c 8 -f -o gpurun_out/prof_gp_$TAG \
  python scripts/bench_configs.py --cases gp --steps 1 --scale 0.5 > gpurun_out/prof_gp_$TAG.log 2>&1
echo "ncu (gauss-point kernels) rc=$?"
# multi-GPU (run with gpurun --gpus 2 or 8): configs 3-5 with the interface exchange, weak scaling
if [ "${GPUS:-1}" -gt 1 ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $GPUS --master-addr 127.0.0.1 --master-port 29517 scripts/bench_dist_configs.py --cases 3,4l,5 --steps 10 \
    > gpurun_out/dist_cfg_${GPUS}gpu_$TAG.jsonl 2> gpurun_out/dist_cfg_${GPUS}gpu_$TAG.err
  echo "multi-GPU configs rc=$?"; cut -c1-400 gpurun_out/dist_cfg_${GPUS}gpu_$TAG.jsonl
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $GPUS --master-addr 127.0.0.1 --master-port 29518 scripts/bench_dist_configs.py --cases 3,5 --steps 10 \
    --opt structured_elasticity=1 > gpurun_out/dist_cfg_struct_${GPUS}gpu_$TAG.jsonl 2> gpurun_out/dist_cfg_struct_${GPUS}gpu_$TAG.err
  echo "multi-GPU configs 3 and 5, structured elasticity kernels rc=$?"; cut -c1-400 gpurun_out/dist_cfg_struct_${GPUS}gpu_$TAG.jsonl
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $GPUS --master-addr 127.0.0.1 --master-port 29519 scripts/bench_dist_configs.py --cases 5 --steps 10 \
    --opt structured_elasticity=1 --no-overlap > gpurun_out/dist_cfg5_noverlap_${GPUS}gpu_$TAG.jsonl 2> gpurun_out/dist_cfg5_noverlap_${GPUS}gpu_$TAG.err
  echo "config 5 without exchange overlap rc=$?"; cut -c1-400 gpurun_out/dist_cfg5_noverlap_${GPUS}gpu_$TAG.jsonl
fi
# one full capture of the structured elasticity kernels (config 3 at half size)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_grid_elast|k_tet_grid" -s 4 -c 2 -f -o gpurun_out/prof_gridelast_$TAG \
  python scripts/bench_configs.py --cases 3 --steps 2 --scale 0.5 --opt structured_elasticity=1 > gpurun_out/prof_gridelast_$TAG.log 2>&1
echo "ncu (structured elasticity kernels) rc=$?"
