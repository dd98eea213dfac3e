#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --sustain-s 0 > gpurun_out/c4.json 2>gpurun_out/c4.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c4.json').read().strip().splitlines()[-1])
for k,v in d['per_config'].items(): print(k, round(v['value'],1), round(v['ms_per_step'],3), round(v['roofline']['frac'],3))
PY
