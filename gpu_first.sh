#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -22 gpurun_out/pytest_gpu.log
timeout 300 python tools/probe_stft.py 2000 > gpurun_out/probe.log 2>&1; tail -4 gpurun_out/probe.log
