#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/rcp_check profiles/rcp_check.cu && /tmp/rcp_check > gpurun_out/r02_rcp_check.json; cat gpurun_out/r02_rcp_check.json
timeout 600 python -m pytest tests/test_gpu_gather_grp.py tests/test_gpu_plume3d_sml.py -m gpu -q -x > gpurun_out/r02_mix_tests.log 2>&1; tail -5 gpurun_out/r02_mix_tests.log
timeout 600 python bench.py --edge 96 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02_bench_n96.json 2> gpurun_out/r02_bench_n96.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_n96.json",):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        for k,v in j["regimes"].items():
            print(f, k, v["value"], v["ms_per_step"], v["roofline"]["stage_ms_per_step"], v["roofline"]["kernel_ms_per_launch"], v.get("adaptive_sml_ms_per_call"))
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-2000:])
PY
