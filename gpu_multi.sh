#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench2.json 2> gpurun_out/bench2.err
echo "bench2 rc=$?"; tail -c 1800 gpurun_out/bench2.json; tail -3 gpurun_out/bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-e2e --workload c5 --utts 3000 > gpurun_out/bench2_c5.json 2> gpurun_out/bench2_c5.err
echo "bench2 c5 rc=$?"; tail -c 600 gpurun_out/bench2_c5.json; tail -3 gpurun_out/bench2_c5.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench2_ref.json 2> gpurun_out/bench2_ref.err
echo "bench2 ref rc=$?"; tail -c 400 gpurun_out/bench2_ref.json
