#!/bin/bash
# round-2 evidence: GPU test-suite, bench line, reference arm, launch list and one full capture of the hot kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log; grep -A14 "worst linear" gpurun_out/pytest_gpu.log > gpurun_out/linear_errors.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02.json 2>/dev/null; echo "ref rc=$?"
CMDP="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --prewarm-s 0 --sustain-s 0"
timeout 300 $CMDP > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv $CMDP > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
CMDQ="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --prewarm-s 0 --sustain-s 0 --no-configs"
timeout 300 $CMDQ > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stft_tc2 -s 3 -c 1 -o gpurun_out/prof_stft_tc2_r02 -f $CMDQ > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full.log
timeout 200 python tools/probe_sizes.py > gpurun_out/probe_sizes.txt 2>&1; tail -7 gpurun_out/probe_sizes.txt
timeout 100 python tools/probe_post.py > gpurun_out/probe_post.txt 2>&1; tail -1 gpurun_out/probe_post.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
