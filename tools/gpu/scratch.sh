#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_post_si.py tests/test_gpu_fullsize.py tests/test_gpu_cli_torch.py -x -q -m gpu -k "cmvn or pipeline or post_chain or cli or standard" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python tools/probe_other.py 2>&1 | tail -3
