#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
pr() { tail -1 $1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(round(j['trees_per_s_e2e'],2), round(j['seconds'],2), round(j['device_seconds'],2), j['nodes_mean'])"; }
timeout 900 python tools/bench_forest.py --trees 148 --cpu-sample 0 > gpurun_out/bf_a.log 2>&1; pr gpurun_out/bf_a.log
SKDIST_B200_FOREST_NODECAP=524288 timeout 900 python tools/bench_forest.py --trees 148 --cpu-sample 0 > gpurun_out/bf_b.log 2>&1; pr gpurun_out/bf_b.log
timeout 900 python tools/bench_forest.py --trees 16 --cpu-sample 0 > gpurun_out/bf_c.log 2>&1; pr gpurun_out/bf_c.log
