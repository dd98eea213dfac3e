#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_post_si.py tests/test_gpu_cli_torch.py -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --workload c5 > gpurun_out/bench_c5.json 2>gpurun_out/bench_c5.err; echo "c5 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_c5.json').read().strip().splitlines()[-1]); print('c5', d['value'], d['ms_per_step'])"; tail -3 gpurun_out/bench_c5.err
