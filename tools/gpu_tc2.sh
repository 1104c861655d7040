#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; grep -E "worst|shape|TC_CHECK|Error|error|timeout" gpurun_out/tc_check.log | tail -12
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/bench_tc2.log 2>&1; tail -1 gpurun_out/bench_tc2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_eval_kernel -s 1 -c 1 -o gpurun_out/prof_tc2 python bench.py --steps 1 --warmup 0 --cpu-sample 0 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out/prof_tc2.ncu-rep
