#!/bin/bash
# Round-2 development check: forest builder (tests, phase profile, config-4 line), tensor-core SGD
# (tests, config-3 line), short ncu capture of the forest builder on a reduced problem.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round2.sh'
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== forest tests"; timeout 600 python -m pytest tests/test_forest_gpu.py -x -q > gpurun_out/pytest_forest.log 2>&1; tail -3 gpurun_out/pytest_forest.log
echo "== sgd tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "sgd" > gpurun_out/pytest_sgd.log 2>&1; tail -15 gpurun_out/pytest_sgd.log
echo "== forest config 4 (phase profile on)"; SKDIST_B200_FOREST_PROF=1 timeout 600 python tools/bench_forest.py --trees ${TREES:-1024} --cpu-sample 0 > gpurun_out/bench_forest_prof.log 2>&1; grep "forest prof" gpurun_out/bench_forest_prof.log | head -20; tail -1 gpurun_out/bench_forest_prof.log | cut -c1-600
echo "== forest config 4"; timeout 600 python tools/bench_forest.py --trees ${TREES:-1024} --cpu-sample 0 > gpurun_out/bench_forest.log 2>&1; tail -1 gpurun_out/bench_forest.log | cut -c1-600
echo "== ovr sgd config 3"; SKDIST_B200_TRACE=2 timeout 900 python tools/bench_ovr.py --cpu-sample 1 > gpurun_out/bench_ovr.log 2>&1; grep "sgd-tc" gpurun_out/bench_ovr.log | tail -8; tail -1 gpurun_out/bench_ovr.log | cut -c1-700
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu forest (reduced problem: 200k rows, one wave)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forest_fast_kernel -c 1 -o gpurun_out/prof_forest_fast python tools/bench_forest.py --n 200000 --trees 1036 --cpu-sample 0 > gpurun_out/ncu_forest.log 2>&1; tail -2 gpurun_out/ncu_forest.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
fi
