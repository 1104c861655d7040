#!/bin/bash
# One gpurun call that validates a tree on a B200: kernel self-check, GPU test suite, smoke, bench.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_check.sh'
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; grep -E "shape|TC_CHECK|rror|timeout" gpurun_out/tc_check.log | tail -6
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-400
