#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python tools/probe_umma.py time > gpurun_out/umma_time.txt 2>&1; echo "time rc=$?"; tail -2 gpurun_out/umma_time.txt
timeout 200 python tools/probe_umma.py probe > gpurun_out/umma_probe.txt 2>&1; echo "probe rc=$?"; tail -4 gpurun_out/umma_probe.txt
CMDP="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --prewarm-s 0"
export PDS_STFT_KERNEL=u
timeout 300 $CMDP > gpurun_out/plain_umma.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stft_umma -s 3 -c 1 -o gpurun_out/prof_stft_umma_r02 -f $CMDP > gpurun_out/ncu_full_umma.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full_umma.log; cat gpurun_out/plain_umma.log | tail -c 900
