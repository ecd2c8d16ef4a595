#!/bin/bash
# multi-GPU parity tests (threads of one process, one context per device), log kept
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_multi_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rs --durations=5 > gpurun_out/r02_pytest_multi.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_multi.log
tail -60 gpurun_out/r02_pytest_multi.log | cut -c1-300
