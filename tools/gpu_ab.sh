#!/bin/bash
# A/B: old epilogue (8 warps) vs new (16 warps + uniform sign path), same box, alternating
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader -lms 500 > gpurun_out/ab_smi.log 2>&1 &
SMI=$!
for rep in 1 2; do
  for v in old new; do
    if [ $v = old ]; then export SKDIST_B200_LIBPATH=$PWD/skdist_b200/lib/libskdist_b200_old.so; else unset SKDIST_B200_LIBPATH; fi
    echo "== $v rep $rep $(date +%s.%N)" | tee -a gpurun_out/ab_smi.log
    STEPS=3 SKDIST_B200_TRACE=1 timeout 300 python tools/step_trace.py 512 2>&1 | grep "^step"
  done
done
kill $SMI
