#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for k in tc scalar; do
  PDS_STFT_KERNEL=$k timeout 100 python tools/probe_stft.py 4000 > gpurun_out/probe_$k.log 2>&1; echo "probe $k rc=$?"; tail -3 gpurun_out/probe_$k.log | head -2
  PDS_STFT_KERNEL=$k timeout 300 python tools/probe_other.py > gpurun_out/probe_other_$k.log 2>&1; echo "other $k rc=$?"; head -2 gpurun_out/probe_other_$k.log
done
