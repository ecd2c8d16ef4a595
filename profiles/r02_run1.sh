#!/bin/bash
# round 2, first GPU call: every GPU test (incl. the new ones and the drop-in program), then the default bench in both
# regimes as the code stood at the start of the round (baseline for the mixed-h work)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_run1_gpu.txt
PSPH_RUN_DROPIN=1 timeout 1500 python -m pytest tests -m gpu -x -q -rs --durations=15 > gpurun_out/r02_run1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_run1_pytest.log
tail -30 gpurun_out/r02_run1_pytest.log
timeout 600 python bench.py --edge 96 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_run1_bench_n96.json 2> gpurun_out/r02_run1_bench_n96.err
tail -c 3000 gpurun_out/r02_run1_bench_n96.json
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_run1_bench_n216.json 2> gpurun_out/r02_run1_bench_n216.err
tail -c 3000 gpurun_out/r02_run1_bench_n216.json; tail -5 gpurun_out/r02_run1_bench_n216.err
