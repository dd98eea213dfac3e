#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_stft.py -x -q -m gpu > gpurun_out/pytest_ws.log 2>&1
echo "stft pytest rc=$?"; tail -3 gpurun_out/pytest_ws.log
PDS_STFT_KERNEL=phased timeout 60 python tools/probe_stft.py 2000 > gpurun_out/probe_phased.log 2>&1; echo "probe rc=$?"; tail -3 gpurun_out/probe_phased.log
PDS_STFT_KERNEL=ws timeout 60 python tools/probe_stft.py 2000 > gpurun_out/probe.log 2>&1; echo "probe rc=$?"; tail -3 gpurun_out/probe.log
