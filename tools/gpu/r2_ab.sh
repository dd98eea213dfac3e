#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
if [ $# -gt 0 ]; then timeout 300 python tools/probe_variants.py 10000 "$@" > gpurun_out/ab.log 2>&1; grep -v Warning gpurun_out/ab.log | grep "variant\|Error\|error" | head -20; fi
if [ -n "$RUN_TESTS" ]; then timeout 300 python -m pytest tests/test_gpu_stft.py -q -k "$RUN_TESTS" 2>&1 | tail -8 | tee gpurun_out/ab_tests.log; fi
for V in $NCU_VARIANT; do
timeout 300 python tools/probe_variants.py 3000 $V > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_ -s 8 -c 1 -o gpurun_out/prof_ab_${V/:/_} -f python tools/probe_variants.py 3000 $V > gpurun_out/ncu_ab.log 2>&1
echo "ncu $V rc=$?"
done
