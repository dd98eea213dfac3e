#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --sustain-s 0 > gpurun_out/c5.json 2>gpurun_out/c5.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c5.json').read().strip().splitlines()[-1])
print('c2', d['ms_per_step'], 'c5', d['per_config']['c5']['ms_per_step'], 'post ms', d['per_config']['c5']['ms_per_step']-d['ms_per_step'])
PY
timeout 200 python -m pytest tests/test_gpu_post_si.py tests/test_gpu_corpus.py -x -q 2>&1 | tail -2
