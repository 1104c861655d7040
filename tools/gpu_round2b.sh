#!/bin/bash
# Round-2 development check (b): forest builder after the occupancy / compact-record changes, tensor-core
# SGD after the metadata / prefetch change, parity report on every logistic fixture, headline bench.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== forest tests"; timeout 600 python -m pytest tests/test_forest_gpu.py -x -q > gpurun_out/pytest_forest.log 2>&1; tail -3 gpurun_out/pytest_forest.log
echo "== sgd + row-bit tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "sgd or row_bit" > gpurun_out/pytest_sgd.log 2>&1; tail -3 gpurun_out/pytest_sgd.log
echo "== forest config 4 (phase profile on)"; SKDIST_B200_FOREST_PROF=1 timeout 600 python tools/bench_forest.py --trees 1024 --cpu-sample 0 > gpurun_out/bench_forest_prof.log 2>&1; grep "forest prof" gpurun_out/bench_forest_prof.log | head -20; tail -1 gpurun_out/bench_forest_prof.log | cut -c1-700
echo "== forest config 4"; timeout 600 python tools/bench_forest.py --trees 1024 --cpu-sample 0 > gpurun_out/bench_forest.log 2>&1; tail -1 gpurun_out/bench_forest.log | cut -c1-700
echo "== ovr sgd config 3"; SKDIST_B200_TRACE=2 timeout 900 python tools/bench_ovr.py --cpu-sample 0 > gpurun_out/bench_ovr.log 2>&1; grep "sgd-tc" gpurun_out/bench_ovr.log | sed -n '2,4p;$p'; tail -1 gpurun_out/bench_ovr.log | cut -c1-500
echo "== parity report"; timeout 900 python tools/parity_report.py > gpurun_out/parity_report.log 2>&1; grep fixture gpurun_out/parity_report.log | cut -c1-900
echo "== parity report, 2 gradient passes"; SKDIST_B200_TC_GPASSES=2 timeout 900 python tools/parity_report.py > gpurun_out/parity_report_g2.log 2>&1; grep tcgen05 gpurun_out/parity_report_g2.log | cut -c1-900
echo "== midsize test"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "midsize" > gpurun_out/pytest_mid.log 2>&1; grep -E "kernel [12]:|passed|failed|Error" gpurun_out/pytest_mid.log | tail -5
echo "== bench A/B gradient passes"; timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/bench_g3.log 2>&1; tail -1 gpurun_out/bench_g3.log | cut -c1-400
SKDIST_B200_TC_GPASSES=2 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/bench_g2.log 2>&1; tail -1 gpurun_out/bench_g2.log | cut -c1-400
echo "== headline bench with the CPU leg"; timeout 1200 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
