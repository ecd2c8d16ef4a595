#!/bin/bash
# GPU tests (all), drop-in stage diff of step 3, then the bench in both regimes at 96^3 and 216^3
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -rs --durations=10 > gpurun_out/r02_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
grep -n "passed\|failed\|^FAILED\|AssertionError" gpurun_out/r02_pytest.log | cut -c1-300 | head -30
timeout 300 python profiles/dropin_diff.py 3 > gpurun_out/r02_dropin_diff3.txt 2>&1; cut -c1-400 gpurun_out/r02_dropin_diff3.txt
timeout 600 python bench.py --edge 96 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_n96.json 2> gpurun_out/r02_bench_n96.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_n96.json",):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        for k,v in j["regimes"].items():
            print(f, k, v["value"], v["ms_per_step"], v["roofline"]["stage_ms_per_step"], v["roofline"]["kernel_ms_per_launch"], v.get("adaptive_sml_ms_per_call"))
        print("e2e", j["e2e"]["value"], "bad", j["invariant_violations"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-2000:])
PY
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_n216.json 2> gpurun_out/r02_bench_n216.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_n216.json",):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        for k,v in j["regimes"].items():
            print(f, k, v["value"], v["ms_per_step"], v["roofline"]["stage_ms_per_step"], v["roofline"]["kernel_ms_per_launch"], v.get("adaptive_sml_ms_per_call"))
        print("e2e", j["e2e"]["value"], "bad", j["invariant_violations"])
    except Exception as e:
        print(f, "unreadable", e); print(open(f.replace(".json",".err")).read()[-2000:])
PY
