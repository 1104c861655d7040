#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_ridge.py > gpurun_out/bench_ridge.log 2>&1; tail -2 gpurun_out/bench_ridge.log
timeout 600 python bench.py --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/bench_tc4.log 2>&1; tail -1 gpurun_out/bench_tc4.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['avg_launch_ms'], j['roofline']['frac'], j['clocks'])"
