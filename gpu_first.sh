#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stft.py -x -q -m gpu -k "(variants and (readme or kaldi or magnitude)) or edge or batch or int16 or preemph or dither or chunk or host" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for f in 2 1 2; do
  PDS_TC_FRAMES=$f timeout 100 python tools/probe_stft.py 10000 > gpurun_out/probe_nf$f.log 2>&1; echo "probe NF=$f rc=$?"; tail -3 gpurun_out/probe_nf$f.log | head -2
done
timeout 300 ncu --metrics l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:stft_tc -s 3 -c 1 python tools/probe_stft.py 2000 2>&1 | grep -E "hit_rate|duration" 
