#!/bin/bash
# scratch session script (development): edit, then `gpurun -- 'bash tools/gpu/scratch.sh'`
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python tools/probe_post.py 2>&1 | tail -1
timeout 300 python tools/probe_umma.py time 2>&1 | tail -2
