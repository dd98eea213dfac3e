#!/bin/bash
# final check of a round: GPU test-suite, smoke, bench line, reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
T0=$(date +%s); timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 )) s"
T0=$(date +%s); timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "ref rc=$? wall=$(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], d['roofline']['kernel'])
print('e2e', d['e2e']['value'], 'floor', d['e2e']['copy_floor'], 'pcm', d['e2e']['int16_pcm_input']['value'])
for k,v in d['per_config'].items(): print(k, round(v['value'],1), round(v['ms_per_step'],3), round(v['roofline']['frac'],3))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], 'clocks', d['clocks'])
r=json.loads(open('gpurun_out/bench_ref_final.json').read().strip().splitlines()[-1])
print('ref arm', r['value'], r['cpu_baseline']['kind'], r['ms_per_step'])
PY
