#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cli_torch.py -x -q > gpurun_out/pytest_torch.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_torch.log
