#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stft.py tests/test_gpu_cli_torch.py tests/test_gpu_post_si.py -x -q -m gpu -k "int16 or (variants and readme) or edge or batch or preemph or dither or cli or chunk or pipeline" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python tools/probe_pcm.py > gpurun_out/probe_pcm.log 2>&1; echo "rc=$?"; cat gpurun_out/probe_pcm.log
