#!/bin/bash
# round-1 evidence: bench line, ncu launch list of the SAME command, one full capture of the top kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
CMD="python bench.py --steps 10 --warmup 3"
timeout 600 $CMD > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"
CMDP="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --prewarm-s 0"
timeout 300 $CMDP > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMDP > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 300 $CMDP > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stft_tc -s 3 -c 1 -o gpurun_out/prof_stft_tc_r01 -f $CMDP > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01.json 2>/dev/null; echo "ref rc=$?"
timeout 300 python tools/probe_other.py > gpurun_out/probe_other.log 2>&1; echo "other rc=$?"; cat gpurun_out/probe_other.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
