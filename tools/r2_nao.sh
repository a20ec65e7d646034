#!/bin/bash
# Nao-cup scenario on the GPU box: parity tests, timings, optionally the rest of the GPU suite.
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_nao.py -x -v -m gpu -s > gpurun_out/nao_tests.txt 2>&1
echo "rc=$?" >> gpurun_out/nao_tests.txt
grep -E "FAILED|ERROR|passed|failed|rc=|Error|assert|differ|reach" gpurun_out/nao_tests.txt | cut -c1-250 | tail -30
timeout 300 python tools/nao_time.py > gpurun_out/nao_time.txt 2>&1
cat gpurun_out/nao_time.txt | tail -20
if [ "$1" = "arm" ]; then
  timeout 600 python -m pytest tests/ -x -q -m gpu -k "arm or scenarios or pprm" > gpurun_out/arm_flat_tests.txt 2>&1; echo "arm flat rc=$?"; tail -3 gpurun_out/arm_flat_tests.txt
  timeout 300 python tools/arm_time.py > gpurun_out/arm_time.txt 2>&1
  MPTG_ARM_WARP_PER_EDGE=1 timeout 300 python tools/arm_time.py >> gpurun_out/arm_time.txt 2>&1
  cat gpurun_out/arm_time.txt
fi
if [ "$1" = "all" ]; then
  timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/tests_all.txt 2>&1; echo "all rc=$?"; tail -5 gpurun_out/tests_all.txt
fi
