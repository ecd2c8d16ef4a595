#!/bin/bash
# new output-column test, FP64 FMA peak, ncu --set full of the pair kernels of both regimes at 96^3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_plume3d.py -m gpu -q -k h5part > gpurun_out/r02_h5part.log 2>&1; tail -5 gpurun_out/r02_h5part.log
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/dfma_peak profiles/dfma_peak.cu && /tmp/dfma_peak > gpurun_out/r02_dfma_peak.json; cat gpurun_out/r02_dfma_peak.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_mix_scan|k_mix_gather|k_dens_near" -s 8 -c 8 -f -o gpurun_out/r02_mix_n96 \
  python bench.py --edge 96 --steps 2 --warmup 1 --no-e2e --no-cpu --regime mixed > gpurun_out/r02_ncu_mix.log 2>&1
tail -3 gpurun_out/r02_ncu_mix.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tile_gather|k_tile_scan2" -s 6 -c 6 -f -o gpurun_out/r02_uni_n96 \
  python bench.py --edge 96 --steps 2 --warmup 1 --no-e2e --no-cpu --regime uniform > gpurun_out/r02_ncu_uni.log 2>&1
tail -3 gpurun_out/r02_ncu_uni.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep | tail -3
