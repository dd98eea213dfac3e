#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stft.py -x -q -m gpu -k "(variants and (readme or kaldi or magnitude)) or edge or batch" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for f in 1 2; do
  timeout 100 python tools/probe_stft.py 10000 > gpurun_out/probe_a$f.log 2>&1; echo "probe rc=$?"; tail -3 gpurun_out/probe_a$f.log | head -2
done
