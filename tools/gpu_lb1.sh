#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 2 --cpu-sample 0 > gpurun_out/bench_lb.log 2>&1; tail -1 gpurun_out/bench_lb.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['avg_launch_ms'], j['roofline']['frac'], j['clocks'], j['config']['best_C'], j['config']['mean_test_score_best'])"
