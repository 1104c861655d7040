#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; grep -E "worst|shape" gpurun_out/tc_check.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_tc.csv python bench.py --kernel 2 --steps 1 --warmup 0 --cpu-sample 0 > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_eval_kernel -s 1 -c 1 -o gpurun_out/prof_tc python bench.py --kernel 2 --steps 1 --warmup 0 --cpu-sample 0 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
