#!/bin/bash
# round 2, run 11: entries of the other gate family dropped before the pair rounds (warp-level packing) in k_mix_gather;
# over 8x the warps; parity tests, 96^3 and 216^3 (both regimes)
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
timeout 900 python -m pytest tests/test_gpu_gather_grp.py tests/test_gpu_plume3d_sml.py tests/test_gpu_plume3d.py tests/test_gpu_cube_lockstep.py tests/test_gpu_shocktube.py -m gpu -q -x > gpurun_out/r02k_tests.log 2>&1; tail -3 gpurun_out/r02k_tests.log
timeout 600 python bench.py --edge 96 --steps 5 --warmup 3 --no-cpu --no-e2e --no-plume > gpurun_out/r02k_n96.json 2> gpurun_out/r02k_n96.err
show gpurun_out/r02k_n96.json
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-plume > gpurun_out/r02k_n216.json 2> gpurun_out/r02k_n216.err
show gpurun_out/r02k_n216.json
