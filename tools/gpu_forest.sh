#!/bin/bash
# Forest builder on a B200: parity tests, config-4 bench line, ncu capture of the builder kernel.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_forest.sh'
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forest_gpu.py -x -q > gpurun_out/pytest_forest.log 2>&1; tail -5 gpurun_out/pytest_forest.log
SKDIST_B200_FOREST_NODECAP=3000 timeout 300 python -m pytest tests/test_forest_gpu.py -x -q -k "bit_identical_to_sklearn" > gpurun_out/pytest_forest_cap.log 2>&1; tail -2 gpurun_out/pytest_forest_cap.log
timeout 600 python tools/bench_forest.py --trees ${TREES:-1024} --cpu-sample ${CPU_SAMPLE:-32} > gpurun_out/bench_forest.log 2>&1; tail -1 gpurun_out/bench_forest.log | cut -c1-900
if [ "${NCU:-1}" = "1" ]; then
# reduced problem: a --set full replay of the config-4 kernel (6 s per pass, ~40 passes) outlives any timeout and wedges the box
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forest_fast_kernel -c 1 -o gpurun_out/prof_forest_fast python tools/bench_forest.py --n 200000 --trees 1036 --cpu-sample 0 > gpurun_out/ncu_forest.log 2>&1; tail -2 gpurun_out/ncu_forest.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
fi
