#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stft.py -x -q -m gpu -k "variants and (readme or kaldi or magnitude)" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for k in tc scalar; do
  PDS_STFT_KERNEL=$k timeout 100 python tools/probe_stft.py 4000 > gpurun_out/probe_$k.log 2>&1; echo "probe $k rc=$?"; tail -3 gpurun_out/probe_$k.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_tc -s 3 -c 1 -o gpurun_out/prof_tc -f python tools/probe_stft.py 2000 > gpurun_out/ncu_tc.log 2>&1
echo "ncu rc=$?"; tail -1 gpurun_out/ncu_tc.log
