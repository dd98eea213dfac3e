#!/bin/bash
# round-1 evidence: bench line, ncu launch list of the SAME command, one full capture of the top kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CMD="python bench.py --steps 5 --warmup 3"
timeout 600 $CMD > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"
CMDP="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
timeout 300 $CMDP > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMDP > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 300 $CMDP > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stft_tc -s 3 -c 1 -o gpurun_out/prof_stft_tc_r01 -f $CMDP > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01.json 2>/dev/null; echo "ref rc=$?"
tail -c 600 gpurun_out/bench_ref_r01.json
