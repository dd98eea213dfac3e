#!/bin/bash
# N-GPU sanity run of the benchmark contract (N = $1, default 4)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-4}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench$N.json 2> gpurun_out/bench$N.err
echo "bench$N rc=$?"; tail -3 gpurun_out/bench$N.err
python -c "
import json; d=json.loads(open('gpurun_out/bench$N.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -c 200
