#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench2.json 2> gpurun_out/bench2.err
echo "bench2 rc=$?"; tail -c 900 gpurun_out/bench2.json; tail -3 gpurun_out/bench2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-e2e --workload c5 --utts 3000 > gpurun_out/bench2_c5.json 2> gpurun_out/bench2_c5.err
echo "bench2 c5 rc=$?"; tail -c 300 gpurun_out/bench2_c5.json; tail -3 gpurun_out/bench2_c5.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench2_ref.json 2> gpurun_out/bench2_ref.err
echo "bench2 ref rc=$?"; tail -c 200 gpurun_out/bench2_ref.json
python - <<'PY'
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize()
    print(name, 5 * n / (time.perf_counter() - t0) / 1e9, "GB/s")
PY
nvidia-smi topo -m | head -8
