#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_post_si.py -x -q -m gpu -k "si_" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
PDS_SI_KERNEL=fft timeout 300 python tools/probe_si.py 2 > gpurun_out/probe_si_fft.log 2>&1; echo "si fft rc=$?"; tail -1 gpurun_out/probe_si_fft.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:si_fft -s 2 -c 1 -o gpurun_out/prof_si_fft -f python tools/probe_si.py 0 > gpurun_out/ncu_si.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/ncu_si.log
