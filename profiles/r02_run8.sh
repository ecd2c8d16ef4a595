#!/bin/bash
# round 2, run 8 (2 GPUs): slab tests incl. domain growth across slabs; the reference's own plume case like for like
# (bench.py "plume" object) at N = 1 and N = 2, small cube line around it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02h_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --durations=6 > gpurun_out/r02h_pytest_multi.log 2>&1; tail -15 gpurun_out/r02h_pytest_multi.log
timeout 900 python bench.py --edge 96 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02h_bench_plume_n96.json 2> gpurun_out/r02h_bench_plume_n96.err
tail -c 600 gpurun_out/r02h_bench_plume_n96.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --edge 96 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02h_bench_plume_n96_2gpu.json 2> gpurun_out/r02h_bench_plume_n96_2gpu.err
tail -c 600 gpurun_out/r02h_bench_plume_n96_2gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/r02h_bench_plume_n96.json", "gpurun_out/r02h_bench_plume_n96_2gpu.json"):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, j["value"], j["ms_per_step"]); print(json.dumps(j.get("plume"), indent=1))
    except Exception as e:
        print(f, "unreadable", e)
PY
