#!/bin/bash
# ncu evidence for profiles/: launch list of one bench step + one full capture of the evaluation kernel.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_eval_kernel -s 5 -c 1 -o gpurun_out/prof_tc python bench.py --steps 1 --warmup 0 --cpu-sample 0 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
