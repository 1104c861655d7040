#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export STEPS=1 SKDIST_B200_TRACE=2 SKDIST_B200_FORCE_ROUNDS=12
for dbg in 4; do
  SKDIST_B200_TC_DEBUG=$dbg timeout 300 python tools/step_trace.py 512 > gpurun_out/step_dbg$dbg.log 2>&1
  echo "debug=$dbg"; grep round gpurun_out/step_dbg$dbg.log | sed -n '4,7p'
done
