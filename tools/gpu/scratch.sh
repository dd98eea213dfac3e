#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
time (timeout 600 python bench.py > gpurun_out/bench_final.json 2>gpurun_out/bench_final.err); echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['int16_pcm_input']['value'], d['roofline']['fp32_frac'], d['clocks'], d['cpu_baseline']['value'])"
time (timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_final.json 2>/dev/null); tail -c 300 gpurun_out/bench_ref_final.json
