#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_post_si.py -x -q -k "si" > gpurun_out/pytest_si.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_si.log
timeout 300 python tools/probe_si_long.py 600 > gpurun_out/probe_si_long.txt 2>&1; cat gpurun_out/probe_si_long.txt
