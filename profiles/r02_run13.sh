#!/bin/bash
# round 2, run 13 (8 GPUs): slab tests with 4 slabs; default bench line (incl. the plume leg) at N = 4 and N = 8, 96^3 per GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02m_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=6 > gpurun_out/r02m_pytest_multi.log 2>&1; tail -12 gpurun_out/r02m_pytest_multi.log
for N in 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --edge 96 --steps 5 --warmup 3 > gpurun_out/r02m_bench_${N}gpu_n96.json 2> gpurun_out/r02m_bench_${N}gpu_n96.err
echo "N=$N rc=$?"; tail -c 400 gpurun_out/r02m_bench_${N}gpu_n96.err
done
python - <<'PY'
import json
for N in (4, 8):
    f = "gpurun_out/r02m_bench_%dgpu_n96.json" % N
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j["value"], j["ms_per_step"], j.get("e2e", {}) and j["e2e"].get("value"))
        for k, v in j["regimes"].items():
            print(k, round(v["value"] / 1e6, 2), "M/s", v["ms_per_step"], v["roofline"]["stage_ms_per_step"].get("halo_exchange"))
        p = j.get("plume") or {}
        print("plume", p.get("value"), p.get("ms_per_step"), p.get("error"))
    except Exception as e:
        print(f, "unreadable", e)
PY
