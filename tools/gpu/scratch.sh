#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stft.py -x -q -k "tcgen05" > gpurun_out/pytest_umma.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_umma.log
timeout 200 python tools/probe_umma.py time > gpurun_out/umma_time.txt 2>&1; echo "time rc=$?"; tail -2 gpurun_out/umma_time.txt
timeout 200 python tools/probe_umma.py probe > gpurun_out/umma_probe.txt 2>&1; echo "probe rc=$?"; tail -4 gpurun_out/umma_probe.txt
