#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/hostmem.txt; nproc >> gpurun_out/hostmem.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 1200 python tools/bench_forest.py --n 2000000 --d 64 --trees 128 --cpu-sample 1 > gpurun_out/bench_forest_2m.log 2>&1; tail -1 gpurun_out/bench_forest_2m.log
timeout 1200 python tools/bench_ovr.py --cpu-sample 1 > gpurun_out/bench_ovr_full.log 2>&1; tail -1 gpurun_out/bench_ovr_full.log
cat gpurun_out/hostmem.txt
