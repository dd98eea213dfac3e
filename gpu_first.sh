#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_post_si.py tests/test_gpu_cli_torch.py -x -q -m gpu -k "si_ or si" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
PDS_SI_KERNEL=direct timeout 600 python -m pytest tests/test_gpu_post_si.py -x -q -m gpu -k "si_" > gpurun_out/pytest_gpu_direct.log 2>&1
echo "pytest direct rc=$?"; tail -2 gpurun_out/pytest_gpu_direct.log
PDS_SI_KERNEL=fft timeout 300 python tools/probe_si.py 2 > gpurun_out/probe_si_fft.log 2>&1; echo "si fft rc=$?"; tail -1 gpurun_out/probe_si_fft.log
