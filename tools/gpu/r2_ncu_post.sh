#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMDP="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --prewarm-s 0 --sustain-s 0"
timeout 300 $CMDP > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deltas25 -s 6 -c 2 -o gpurun_out/prof_deltas25 -f $CMDP > gpurun_out/ncu_post.log 2>&1
echo "capture rc=$?"; tail -2 gpurun_out/ncu_post.log
