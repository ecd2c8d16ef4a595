#!/bin/bash
# round 2, run 9: row state of the mixed-h mom gather in uniform registers (A) against the plain build (B), same runs;
# pair-path diagnostics of both regimes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
show() {
python - "$1" <<'PY'
import json, sys
f = sys.argv[1]
try:
    j = json.loads(open(f).read().strip().splitlines()[-1])
    for k, v in j["regimes"].items():
        print(f, k, round(v["value"] / 1e6, 2), "M/s", v["ms_per_step"], v["roofline"]["kernel_ms_per_launch"], v.get("pair_path_rank0"))
except Exception as e:
    print(f, "unreadable", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
}
for rep in 1 2; do
timeout 600 python bench.py --edge 96 --steps 5 --warmup 3 --no-cpu --no-e2e --no-plume > gpurun_out/r02i_n96_uni_$rep.json 2> gpurun_out/r02i_n96_uni_$rep.err
show gpurun_out/r02i_n96_uni_$rep.json
PSPH_LIB=$PWD/plume_sph_b200/libpsph_nouni.so timeout 600 python bench.py --edge 96 --steps 5 --warmup 3 --no-cpu --no-e2e --no-plume > gpurun_out/r02i_n96_plain_$rep.json 2> gpurun_out/r02i_n96_plain_$rep.err
show gpurun_out/r02i_n96_plain_$rep.json
done
timeout 600 python -m pytest tests/test_gpu_gather_grp.py tests/test_gpu_plume3d_sml.py -m gpu -q -x > gpurun_out/r02i_tests.log 2>&1; tail -3 gpurun_out/r02i_tests.log
