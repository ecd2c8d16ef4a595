#!/bin/bash
# round 2, run 12 (2 GPUs): slab tests incl. domain growth across slabs; default bench line at N = 2 (216^3 per GPU, both regimes,
# halo_exchange stage time, the reference's plume strong-scaled)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02l_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=6 > gpurun_out/r02l_pytest_multi.log 2>&1; tail -15 gpurun_out/r02l_pytest_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02l_bench_2gpu.json 2> gpurun_out/r02l_bench_2gpu.err
tail -c 800 gpurun_out/r02l_bench_2gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/r02l_bench_2gpu.json",):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j["value"], j["ms_per_step"], j.get("e2e", {}) and j["e2e"].get("value"))
        for k, v in j["regimes"].items():
            print(k, round(v["value"] / 1e6, 2), "M/s", v["ms_per_step"], v["roofline"]["stage_ms_per_step"])
        print(json.dumps(j.get("plume"), indent=1))
    except Exception as e:
        print(f, "unreadable", e)
PY
