#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q -s > gpurun_out/r02_dropin_pytest.log 2>&1; tail -40 gpurun_out/r02_dropin_pytest.log | cut -c1-600
