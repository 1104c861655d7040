#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py > gpurun_out/tc_check.log 2>&1; grep -E "shape|TC_CHECK|Error|error|timeout" gpurun_out/tc_check.log | tail -6
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for split in aligned balanced; do
  SKDIST_B200_TC_SPLIT=$split timeout 600 python bench.py --steps 2 --warmup 1 --cpu-sample 0 > gpurun_out/bench_tc3_$split.log 2>&1; echo $split; tail -1 gpurun_out/bench_tc3_$split.log | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['avg_launch_ms'], j['roofline']['frac'], j['clocks'])"
done
SKDIST_B200_TC_SPLIT=aligned timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_eval_kernel -s 1 -c 1 -o gpurun_out/prof_tc3 python bench.py --steps 1 --warmup 0 --cpu-sample 0 > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out/prof_tc3.ncu-rep
