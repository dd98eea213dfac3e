#!/bin/bash
# N-GPU checks (N = $1): NCCL all-reduce test, bench under torchrun, the CLI under torchrun
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_corpus.py -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench$N.json 2> gpurun_out/bench$N.err
echo "bench$N rc=$?"; tail -3 gpurun_out/bench$N.err | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench$N.json').read().splitlines() if l.startswith('{')][-1])
print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'pcm', d['e2e']['int16_pcm_input'] and d['e2e']['int16_pcm_input']['value'])
for k,v in d['per_config'].items(): print(k, v['value'], v['ms_per_step'], v.get('collective_us'))
PY
timeout 900 python tools/cli_bench.py $N ${CLI_UTTS:-2000} 2>&1 | tail -1 | tee gpurun_out/cli_bench_$N.json
nvidia-smi topo -m | head -12 > gpurun_out/topo$N.txt
