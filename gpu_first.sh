#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stft.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/probe_sizes.py > gpurun_out/probe_sizes.log 2>&1; echo "rc=$?"; cat gpurun_out/probe_sizes.log | tail -12
timeout 100 python tools/probe_stft.py 10000 > gpurun_out/probe_a.log 2>&1; tail -3 gpurun_out/probe_a.log | head -2
