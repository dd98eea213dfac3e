#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/probe_si.py 0 > gpurun_out/plain_si.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:si_fft -s 2 -c 1 -o gpurun_out/prof_si_fft -f python tools/probe_si.py 0 > gpurun_out/ncu_si.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/ncu_si.log; cat gpurun_out/plain_si.log
