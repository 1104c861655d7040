#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; grep -E "shape|TC_CHECK|rror|timeout" gpurun_out/tc_check.log | tail -6
STEPS=2 SKDIST_B200_TRACE=2 timeout 300 python tools/step_trace.py 512 > gpurun_out/step_epi.log 2>&1; grep -E "round +(2|3|40|80|100) |step" gpurun_out/step_epi.log | tail -8
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for dbg in 3 4; do
  STEPS=1 SKDIST_B200_TRACE=2 SKDIST_B200_FORCE_ROUNDS=8 SKDIST_B200_TC_DEBUG=$dbg timeout 300 python tools/step_trace.py 512 > gpurun_out/step_dbg$dbg.log 2>&1
  echo "debug=$dbg"; grep round gpurun_out/step_dbg$dbg.log | sed -n '4,5p'
done
