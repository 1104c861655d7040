#!/bin/bash
# first GPU visit: tests, smoke, parity diagnostics, reduced + full bench, ncu launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 300 python tools/diag_parity.py > gpurun_out/diag.log 2>&1
timeout 600 python bench.py --candidates 64 --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/bench_small.log 2>&1
timeout 900 python bench.py --steps 1 --warmup 1 --cpu-sample 1 > gpurun_out/bench_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --n 200000 --candidates 32 --steps 1 --warmup 0 --cpu-sample 0 > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench_small.log; tail -2 gpurun_out/bench_full.log
