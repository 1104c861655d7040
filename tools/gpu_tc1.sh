#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python tools/tc_check.py 1 > gpurun_out/tc_check1.log 2>&1; echo "rc=$?" >> gpurun_out/tc_check1.log
tail -15 gpurun_out/tc_check1.log
if grep -q "TC_CHECK PASS" gpurun_out/tc_check1.log; then
  timeout 400 python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; echo "rc=$?" >> gpurun_out/tc_check.log
  tail -25 gpurun_out/tc_check.log
  timeout 300 python tools/diag_parity.py 2 > gpurun_out/diag_tc.log 2>&1
  tail -5 gpurun_out/diag_tc.log
  timeout 600 python bench.py --kernel 2 --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/bench_tc.log 2>&1
  tail -2 gpurun_out/bench_tc.log
fi
