#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
# 1. plain bench (the number), full default settings incl. cpu baseline
timeout 900 python bench.py > gpurun_out/bench_r01_final.log 2>&1; tail -1 gpurun_out/bench_r01_final.log | cut -c1-300
# 2. reference arm
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r01_reference.log 2>&1; tail -1 gpurun_out/bench_r01_reference.log | cut -c1-400
# 3. launch list of the same command (one step)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_tc_v4.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/ncu_list.log 2>&1; tail -2 gpurun_out/ncu_list.log | cut -c1-200
# 4. full capture of the evaluation kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_eval_kernel -s 5 -c 1 -o gpurun_out/prof_tc_v4 python bench.py --steps 1 --warmup 0 --cpu-sample 0 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
