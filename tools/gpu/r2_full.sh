#!/bin/bash
# full GPU test-suite + bench line (+ optional reference arm)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-6000
if [ -n "$RUN_REF" ]; then timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json; fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
