#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== forest tests"; timeout 600 python -m pytest tests/test_forest_gpu.py -x -q > gpurun_out/pytest_forest.log 2>&1; tail -2 gpurun_out/pytest_forest.log
echo "== forest config 4"; timeout 600 python tools/bench_forest.py --trees 1024 --cpu-sample 0 > gpurun_out/bench_forest.log 2>&1; tail -1 gpurun_out/bench_forest.log | cut -c1-700
echo "== forest config 4 (phase profile)"; SKDIST_B200_FOREST_PROF=1 timeout 600 python tools/bench_forest.py --trees 1024 --cpu-sample 0 > gpurun_out/bench_forest_prof.log 2>&1; grep "forest prof" gpurun_out/bench_forest_prof.log | head -20
