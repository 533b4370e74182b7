# Round 2, call b2: ncu --set full of the config-5 structured kernels, the config-4 P2 tile kernels and the config-3 structured kernels
# (gpurun_out must stay below 64 MiB in total or nothing is copied back: few launches per report).
TAG=${1:-r2b}
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_presum_coef|k_tet_grid" -s 9 -c 3 -f -o gpurun_out/prof_cfg5_$TAG \
  python scripts/bench_configs.py --cases 5 --steps 1 > gpurun_out/prof_cfg5_$TAG.log 2>&1
echo "ncu cfg5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tile_fwd|k_tile_adj" -s 5 -c 2 -f -o gpurun_out/prof_cfg4_$TAG \
  python scripts/bench_configs.py --cases 4l --steps 1 > gpurun_out/prof_cfg4_$TAG.log 2>&1
echo "ncu cfg4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_grid_elast" -s 5 -c 2 -f -o gpurun_out/prof_cfg3_$TAG \
  python scripts/bench_configs.py --cases 3 --steps 1 > gpurun_out/prof_cfg3_$TAG.log 2>&1
echo "ncu cfg3 rc=$?"
ls -la gpurun_out/*.ncu-rep
