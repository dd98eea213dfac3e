#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "Race reported\|hazard" gpurun_out/racecheck.log; tail -6 gpurun_out/racecheck.log
