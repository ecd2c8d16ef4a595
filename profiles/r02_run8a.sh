#!/bin/bash
# round 2, run 8a (1 GPU): the reference's own plume case like for like (bench.py "plume" object), small cube line around it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --edge 96 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02h_bench_plume_n96.json 2> gpurun_out/r02h_bench_plume_n96.err
tail -c 600 gpurun_out/r02h_bench_plume_n96.err
python - <<'PY'
import json
for f in ("gpurun_out/r02h_bench_plume_n96.json",):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j["value"], j["ms_per_step"]); print(json.dumps(j.get("plume"), indent=1))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 python -m pytest tests/test_gpu_grow.py tests/test_gpu_erupt.py tests/test_gpu_bcond.py -m gpu -q > gpurun_out/r02h_tests.log 2>&1; tail -3 gpurun_out/r02h_tests.log
