#!/bin/bash
# Round-2 development check (d): parity tests, forest + SGD timings, ncu captures, bench lines of configs 3-5.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "== midsize + row-bit tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -k "midsize or row_bit" > gpurun_out/pytest_mid.log 2>&1; tail -3 gpurun_out/pytest_mid.log
echo "== forest tests"; timeout 600 python -m pytest tests/test_forest_gpu.py -x -q > gpurun_out/pytest_forest.log 2>&1; tail -2 gpurun_out/pytest_forest.log
echo "== forest config 4"; timeout 600 python tools/bench_forest.py --trees 1024 --cpu-sample 0 > gpurun_out/bench_forest.log 2>&1; tail -1 gpurun_out/bench_forest.log | cut -c1-700
echo "== ovr sgd config 3"; SKDIST_B200_TRACE=2 timeout 900 python tools/bench_ovr.py --cpu-sample 0 > gpurun_out/bench_ovr.log 2>&1; grep "sgd-tc" gpurun_out/bench_ovr.log | sed -n '2,4p;22,23p;$p'; tail -1 gpurun_out/bench_ovr.log | cut -c1-400
echo "== ncu forest (reduced problem: 200k rows, one wave)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forest_fast_kernel -c 1 -o gpurun_out/prof_forest_fast python tools/bench_forest.py --n 200000 --trees 1036 --cpu-sample 0 > gpurun_out/ncu_forest.log 2>&1; tail -1 gpurun_out/ncu_forest.log | cut -c1-200
echo "== ncu sgd (100k rows)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgd_scan_kernel -s 60 -c 1 -o gpurun_out/prof_sgd_scan python tools/bench_ovr.py --n 100000 --cpu-sample 0 > gpurun_out/ncu_sgd_scan.log 2>&1; tail -1 gpurun_out/ncu_sgd_scan.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgd_gemm_kernel -s 121 -c 1 -o gpurun_out/prof_sgd_gemm python tools/bench_ovr.py --n 100000 --cpu-sample 0 > gpurun_out/ncu_sgd_gemm.log 2>&1; tail -1 gpurun_out/ncu_sgd_gemm.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
echo "== bench config 3"; timeout 900 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/bench_c3.log 2>&1; tail -1 gpurun_out/bench_c3.log
echo "== bench config 5"; timeout 900 python bench.py --config 5 --steps 3 --warmup 2 > gpurun_out/bench_c5.log 2>&1; tail -1 gpurun_out/bench_c5.log
echo "== bench config 4"; timeout 1200 python bench.py --config 4 --steps 1 --warmup 1 > gpurun_out/bench_c4.log 2>&1; tail -1 gpurun_out/bench_c4.log
