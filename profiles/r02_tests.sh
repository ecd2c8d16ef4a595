#!/bin/bash
# every GPU test (no -x: one call shows every failure), log kept under gpurun_out/
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q -rs --durations=15 "$@" > gpurun_out/r02_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
tail -40 gpurun_out/r02_pytest.log
