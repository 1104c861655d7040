#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1200 python bench.py --impl reference --steps 2 --warmup 3 ) > gpurun_out/bench_ref2.log 2>&1; tail -5 gpurun_out/bench_ref2.log | cut -c1-900
( time timeout 1200 python bench.py ) > gpurun_out/bench_default.log 2>&1; grep -E "^\{|real" gpurun_out/bench_default.log | cut -c1-2500
