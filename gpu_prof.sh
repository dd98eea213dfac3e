#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD2="env PDS_STFT_KERNEL=ws python tools/probe_stft.py 600"
timeout 120 $CMD2 > gpurun_out/plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_ws -s 3 -c 1 -o gpurun_out/prof_ws -f $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/ncu_full.log
