#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_post_si.py -x -q -m gpu -k "si_" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for k in fft direct; do
  PDS_SI_KERNEL=$k timeout 300 python tools/probe_si.py 2 > gpurun_out/probe_si_$k.log 2>&1; echo "si $k rc=$?"; tail -3 gpurun_out/probe_si_$k.log
done
