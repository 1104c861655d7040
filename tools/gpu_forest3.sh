#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forest_gpu.py -m gpu -q 2>&1 | tail -2
timeout 1500 python tools/bench_forest.py --trees 1024 --cpu-sample 1 > gpurun_out/bench_forest_1024.log 2>&1; tail -1 gpurun_out/bench_forest_1024.log | cut -c1-900
