#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_ridge.py > gpurun_out/bench_ridge.log 2>&1; tail -1 gpurun_out/bench_ridge.log
timeout 900 python tools/bench_ovr.py --n 100000 --d 512 --k 200 --cpu-sample 1 > gpurun_out/bench_ovr_small.log 2>&1; tail -1 gpurun_out/bench_ovr_small.log
