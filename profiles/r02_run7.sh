#!/bin/bash
# round 2, run 7: 3-D GSPH (reconstructed Riemann states, Roe) and domain growth on the device + the whole GPU suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_grow.py tests/test_gpu_plume3d.py tests/test_gpu_plume3d_sml.py -m gpu -q -x --durations=5 > gpurun_out/r02g_new_tests.log 2>&1; tail -25 gpurun_out/r02g_new_tests.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02g_pytest_gpu.log 2>&1; tail -15 gpurun_out/r02g_pytest_gpu.log
