#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forest_gpu.py -x -q > gpurun_out/pytest_forest.log 2>&1; tail -15 gpurun_out/pytest_forest.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_forest.py --n 200000 --d 64 --trees 32 --cpu-sample 1 > gpurun_out/bench_forest_small.log 2>&1; tail -1 gpurun_out/bench_forest_small.log
