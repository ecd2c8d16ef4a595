#!/bin/bash
# round 2, run 15 (last GPU seconds of the round): the default bench line at full size, the drop-in program's tests, and an A/B of
# the dynamic row cursor in k_mix_gather (PSPH_MIX_DYN=1 build: libpsph_dyn.so) with its parity tests
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 150 python bench.py > gpurun_out/r02o_bench_default.json 2> gpurun_out/r02o_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r02o_bench_default.json").read().strip().splitlines()[-1])
    print("value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"] and j["e2e"]["value"], "cpu", j.get("cpu_baseline", {}).get("value"))
    for k, v in j["regimes"].items():
        print(k, round(v["value"] / 1e6, 2), "M/s", v["ms_per_step"])
    p = j.get("plume") or {}
    print("plume", p.get("value"), p.get("ms_per_step"), p.get("speedup_vs_all_host_cores"), p.get("error"))
except Exception as e:
    print("unreadable", e); print(open("gpurun_out/r02o_bench_default.err").read()[-800:])
PY
timeout 90 python -m pytest tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r02o_dropin.log 2>&1; tail -2 gpurun_out/r02o_dropin.log
export PSPH_LIB=$PWD/plume_sph_b200/libpsph_dyn.so
timeout 50 python -m pytest tests/test_gpu_gather_grp.py tests/test_gpu_plume3d_sml.py -m gpu -q -x > gpurun_out/r02o_dyn_tests.log 2>&1; tail -1 gpurun_out/r02o_dyn_tests.log
for v in dyn std; do
  [ $v = std ] && unset PSPH_LIB
  timeout 40 python bench.py --edge 96 --steps 5 --warmup 3 --no-cpu --no-e2e --no-plume --regime mixed > gpurun_out/r02o_n96_$v.json 2> gpurun_out/r02o_n96_$v.err
  python - $v <<'PY'
import json, sys
try:
    j = json.loads(open("gpurun_out/r02o_n96_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    v = j["regimes"]["mixed_h"]; print(sys.argv[1], v["ms_per_step"], v["roofline"]["kernel_ms_per_launch"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
