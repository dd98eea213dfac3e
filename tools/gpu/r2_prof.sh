#!/bin/bash
# round 2, call 1: baseline bench line + full ncu capture (with source) of the hot kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMDP="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --prewarm-s 0"
timeout 300 $CMDP > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stft_tc -s 3 -c 1 -o gpurun_out/prof_stft_tc_r02a -f $CMDP > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full.log; cat gpurun_out/plain.log | tail -c 600
