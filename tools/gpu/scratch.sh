#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deltas25s -s 10 -c 1 -o gpurun_out/prof_post_apply -f python tools/probe_post.py > gpurun_out/ncu_post3.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_post3.log
