#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:forest_build_kernel -c 1 -o gpurun_out/prof_forest python tools/bench_forest.py --n 200000 --trees 8 --cpu-sample 0 > gpurun_out/ncu_forest.log 2>&1
tail -2 gpurun_out/ncu_forest.log | cut -c1-300
ls -la gpurun_out/prof_forest.ncu-rep
