# Round 2, call x: ncu --set full of the MAPPED structured kernels (config 2 connectivity on mapped node positions, 33.5 M triangles)
TAG=${1:-r2x}
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_grid_fwd|k_grid_adj" -s 12 -c 2 -f -o gpurun_out/prof_cfg2m_$TAG \
  python bench.py --config 2m --extra-configs none --no-cpu-baseline --e2e-steps 0 --steps 3 --warmup 3 > gpurun_out/prof_cfg2m_$TAG.log 2>&1
echo "ncu cfg2m rc=$?"; tail -2 gpurun_out/prof_cfg2m_$TAG.log | cut -c1-300
ls -la gpurun_out/prof_cfg2m_$TAG.ncu-rep
