#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
pr() { tail -1 $1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['avg_launch_ms'], j['roofline']['frac'], j['clocks'])"; }
timeout 600 python bench.py --steps 3 --warmup 2 --cpu-sample 0 > gpurun_out/bench_epi.log 2>&1; pr gpurun_out/bench_epi.log
for dbg in 3 4; do
SKDIST_B200_FORCE_ROUNDS=100 SKDIST_B200_TC_DEBUG=$dbg timeout 600 python bench.py --steps 3 --warmup 2 --cpu-sample 0 > gpurun_out/bench_dbg$dbg.log 2>&1; pr gpurun_out/bench_dbg$dbg.log
done
SKDIST_B200_FORCE_ROUNDS=100 timeout 600 python bench.py --steps 3 --warmup 2 --cpu-sample 0 > gpurun_out/bench_dbg0.log 2>&1; pr gpurun_out/bench_dbg0.log
