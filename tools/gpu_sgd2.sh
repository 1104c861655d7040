#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/sgd_check.py > gpurun_out/sgd_check.log 2>&1; grep -c "0.00e+00 intercept diff 0.00e+00" gpurun_out/sgd_check.log; grep -v "0.00e+00 intercept diff 0.00e+00" gpurun_out/sgd_check.log | head -5
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sgd" 2>&1 | tail -3
SKDIST_B200_TRACE=2 timeout 900 python tools/bench_ovr.py --cpu-sample 1 > gpurun_out/bench_ovr_spec2.log 2>&1; grep "sgd epoch" gpurun_out/bench_ovr_spec2.log | awk "NR<=3 || NR%8==0"; tail -1 gpurun_out/bench_ovr_spec2.log | cut -c1-800
