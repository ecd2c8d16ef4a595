#!/bin/bash
# round 2, run 14 (final profiles, bounded to the GPU-minutes left): bcond / growth tests, smoke(), ncu launch list of the bench
# command at 216^3 (both regimes), ncu --set full of the pair kernels of both regimes at 96^3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_bcond.py tests/test_gpu_grow.py -m gpu -q > gpurun_out/r02n_tests.log 2>&1; tail -2 gpurun_out/r02n_tests.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02n_smoke.log 2>&1; tail -1 gpurun_out/r02n_smoke.log
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02n_launches_n216.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-plume > gpurun_out/r02n_launches_bench.log 2>&1
tail -c 200 gpurun_out/r02n_launches_bench.log; echo
timeout 70 ncu --set full --clock-control none --import-source on -k regex:"k_mix_scan|k_mix_gather" -s 3 -c 3 -f -o gpurun_out/r02n_mix_n96 \
  python bench.py --edge 96 --steps 1 --warmup 1 --no-e2e --no-cpu --no-plume --regime mixed > gpurun_out/r02n_ncu_mix.log 2>&1
timeout 60 ncu --set full --clock-control none --import-source on -k regex:"k_tile_gather" -s 3 -c 2 -f -o gpurun_out/r02n_uni_n96 \
  python bench.py --edge 96 --steps 1 --warmup 1 --no-e2e --no-cpu --no-plume --regime uniform > gpurun_out/r02n_ncu_uni.log 2>&1
ls -la gpurun_out/r02n_* | awk '{print $5, $9}'
