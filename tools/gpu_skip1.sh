#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/sgd_check.py > gpurun_out/sgd_check.log 2>&1; cat gpurun_out/sgd_check.log | tail -26
timeout 300 python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; grep -E "shape|TC_CHECK|rror|timeout" gpurun_out/tc_check.log | tail -6
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/diag_parity.py 2 > gpurun_out/diag_tc.log 2>&1; grep -c "status" gpurun_out/diag_tc.log
timeout 600 python bench.py --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/bench_tc5.log 2>&1; tail -1 gpurun_out/bench_tc5.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['avg_launch_ms'], j['roofline']['frac'], j['clocks'], j['config']['best_C'], j['config']['mean_test_score_best'])"
