#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python tools/probe_umma.py check > gpurun_out/umma_check.txt 2>&1; echo "check rc=$?"; grep -E "kernel|48000|Error|error" gpurun_out/umma_check.txt | head -20
timeout 200 python tools/probe_umma.py time > gpurun_out/umma_time.txt 2>&1; echo "time rc=$?"; cat gpurun_out/umma_time.txt | tail -4
if [ -n "$PROF" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_umma -s 2 -c 1 -o gpurun_out/prof_umma -f python tools/probe_umma.py prof 2000 > gpurun_out/ncu_umma.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_umma.log
fi
