#!/bin/bash
# The round-2 measurements behind profiles/r02_*: pick the sections with STEPS (default: all but ncu).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'STEPS="tests forest sgd" bash tools/gpu_round2.sh'
# ncu captures run on REDUCED problems only (a --set full replay of the 6 s forest kernel does not end within
# any sensible timeout) and every command sits under its own `timeout`.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
STEPS=${STEPS:-"tests forest sgd parity bench configs"}
has() { [[ " $STEPS " == *" $1 "* ]]; }
if has tests; then
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
fi
if has forest; then      # r02_bench_forest_fast_v*.json, r02_forest_fast_phase_profile_v*.txt
echo "== forest config 4"; timeout 600 python tools/bench_forest.py --trees 1024 --cpu-sample 0 > gpurun_out/bench_forest.log 2>&1; tail -1 gpurun_out/bench_forest.log | cut -c1-700
echo "== forest config 4 (phase profile)"; SKDIST_B200_FOREST_PROF=1 timeout 600 python tools/bench_forest.py --trees 1024 --cpu-sample 0 > gpurun_out/bench_forest_prof.log 2>&1; grep "forest prof" gpurun_out/bench_forest_prof.log | head -20
fi
if has sgd; then         # r02_bench_ovr_sgd_tc_v*.json
echo "== ovr sgd config 3"; SKDIST_B200_TRACE=2 timeout 900 python tools/bench_ovr.py --cpu-sample 1 > gpurun_out/bench_ovr.log 2>&1; grep "sgd-tc" gpurun_out/bench_ovr.log | tail -4; tail -1 gpurun_out/bench_ovr.log | cut -c1-700
fi
if has parity; then      # r02_parity_report*.jsonl, r02_bench_{three,two}_gradient_passes.json
echo "== parity report"; timeout 900 python tools/parity_report.py > gpurun_out/parity_report.log 2>&1; grep fixture gpurun_out/parity_report.log | cut -c1-900
echo "== parity report, 2 gradient passes"; SKDIST_B200_TC_GPASSES=2 timeout 900 python tools/parity_report.py > gpurun_out/parity_report_g2.log 2>&1; grep tcgen05 gpurun_out/parity_report_g2.log | cut -c1-900
timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/bench_g3.log 2>&1; tail -1 gpurun_out/bench_g3.log | cut -c1-400
SKDIST_B200_TC_GPASSES=2 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-sample 0 > gpurun_out/bench_g2.log 2>&1; tail -1 gpurun_out/bench_g2.log | cut -c1-400
fi
if has bench; then       # r02_bench_headline_v1.json
echo "== headline bench with the CPU leg and the parity block"; timeout 1200 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
fi
if has configs; then     # r02_bench_config{3,4,5}_v1.json
echo "== bench config 3"; timeout 900 python bench.py --config 3 --steps 2 --warmup 1 > gpurun_out/bench_c3.log 2>&1; tail -1 gpurun_out/bench_c3.log
echo "== bench config 5"; timeout 900 python bench.py --config 5 --steps 3 --warmup 2 > gpurun_out/bench_c5.log 2>&1; tail -1 gpurun_out/bench_c5.log
echo "== bench config 4"; timeout 1200 python bench.py --config 4 --steps 1 --warmup 1 > gpurun_out/bench_c4.log 2>&1; tail -1 gpurun_out/bench_c4.log
fi
if has ncu; then         # r02_ncu_*_summary.txt (tools/ncu_summary.py reads the .ncu-rep files back home)
echo "== ncu forest (200k rows, one wave)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:forest_fast_kernel -c 1 -o gpurun_out/prof_forest_fast python tools/bench_forest.py --n 200000 --trees 1036 --cpu-sample 0 > gpurun_out/ncu_forest.log 2>&1; tail -1 gpurun_out/ncu_forest.log | cut -c1-200
echo "== ncu sgd (100k rows)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgd_scan_kernel -s 60 -c 1 -o gpurun_out/prof_sgd_scan python tools/bench_ovr.py --n 100000 --cpu-sample 0 > gpurun_out/ncu_sgd_scan.log 2>&1; tail -1 gpurun_out/ncu_sgd_scan.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sgd_gemm_kernel -s 121 -c 1 -o gpurun_out/prof_sgd_gemm python tools/bench_ovr.py --n 100000 --cpu-sample 0 > gpurun_out/ncu_sgd_gemm.log 2>&1; tail -1 gpurun_out/ncu_sgd_gemm.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
fi
