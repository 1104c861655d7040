#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/step_trace.py > gpurun_out/step_trace.log 2>&1; grep -E "trace|step" gpurun_out/step_trace.log | tail -30
timeout 300 python tools/sgd_check.py > gpurun_out/sgd_check.log 2>&1; grep log_loss gpurun_out/sgd_check.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
