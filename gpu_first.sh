#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv,noheader -lms 50 > gpurun_out/clocks.log &
SMI=$!
for k in tc scalar tc scalar; do
  PDS_STFT_KERNEL=$k timeout 100 python tools/probe_stft.py 10000 > gpurun_out/probe_$k.log 2>&1; echo "probe $k rc=$?"; tail -3 gpurun_out/probe_$k.log | head -2
done
kill $SMI
sort gpurun_out/clocks.log | uniq -c | sort -rn | head -8
