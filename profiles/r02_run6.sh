#!/bin/bash
# round 2, run 6: pair body with the third-order reciprocal + regrouped heat term; warps of the mixed-h mom gather
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/rcp_check profiles/rcp_check.cu && /tmp/rcp_check > gpurun_out/r02f_rcp_check.json; cat gpurun_out/r02f_rcp_check.json
timeout 900 python -m pytest tests/test_gpu_gather_grp.py tests/test_gpu_plume3d_sml.py tests/test_gpu_plume3d.py tests/test_gpu_cube_lockstep.py tests/test_gpu_shocktube.py -m gpu -q -x > gpurun_out/r02f_tests.log 2>&1; tail -5 gpurun_out/r02f_tests.log
show() {
python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    j = json.loads(open(f).read().strip().splitlines()[-1])
    for k, v in j["regimes"].items():
        print(f, k, round(v["value"] / 1e6, 2), "M/s", v["ms_per_step"], v["roofline"]["kernel_ms_per_launch"], v.get("adaptive_sml_ms_per_call"))
except Exception as e:
    print(f, "unreadable", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
}
for mw in 16 20 24; do
  PSPH_MWARPS=$mw timeout 600 python bench.py --edge 96 --steps 5 --warmup 3 --no-cpu --no-e2e --regime mixed > gpurun_out/r02f_n96_mw$mw.json 2> gpurun_out/r02f_n96_mw$mw.err
  show gpurun_out/r02f_n96_mw$mw.json
done
for tw in 161 201 241; do
  PSPH_TWARPS=$tw timeout 600 python bench.py --edge 96 --steps 5 --warmup 3 --no-cpu --no-e2e --regime uniform > gpurun_out/r02f_n96_tw$tw.json 2> gpurun_out/r02f_n96_tw$tw.err
  show gpurun_out/r02f_n96_tw$tw.json
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02f_n216.json 2> gpurun_out/r02f_n216.err
show gpurun_out/r02f_n216.json
