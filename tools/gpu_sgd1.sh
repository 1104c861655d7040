#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/sgd_check.py > gpurun_out/sgd_check.log 2>&1; grep -c "0.00e+00 intercept diff 0.00e+00" gpurun_out/sgd_check.log; grep -v "0.00e+00 intercept diff 0.00e+00" gpurun_out/sgd_check.log | head -5
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sgd" 2>&1 | tail -3
timeout 900 python tools/bench_ovr.py --cpu-sample 1 > gpurun_out/bench_ovr_spec.log 2>&1; tail -1 gpurun_out/bench_ovr_spec.log | cut -c1-700
SKDIST_B200_SGD_SPEC=0 timeout 900 python tools/bench_ovr.py --cpu-sample 0 > gpurun_out/bench_ovr_nospec.log 2>&1; tail -1 gpurun_out/bench_ovr_nospec.log | cut -c1-300
