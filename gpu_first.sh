#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 60 python tools/probe_stft.py 2000 > gpurun_out/probe_phased.log 2>&1; echo "probe rc=$?"; tail -3 gpurun_out/probe_phased.log
PDS_STFT_KERNEL=ws timeout 60 python tools/probe_stft.py 2000 > gpurun_out/probe.log 2>&1; echo "probe rc=$?"; tail -3 gpurun_out/probe.log
