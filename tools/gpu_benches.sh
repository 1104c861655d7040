#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python tools/bench_ridge.py > gpurun_out/bench_ridge_r01.log 2>&1; tail -1 gpurun_out/bench_ridge_r01.log | cut -c1-900
timeout 1500 python tools/bench_forest.py --trees 512 --cpu-sample 1 > gpurun_out/bench_forest_512.log 2>&1; tail -1 gpurun_out/bench_forest_512.log | cut -c1-900
