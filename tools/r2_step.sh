#!/bin/bash
# Nao far-skip check + C5 full-size tests + ramp timings
mkdir -p gpurun_out
timeout 600 python -u -m pytest tests/test_nao.py -x -q -m gpu > gpurun_out/nao_tests.txt 2>&1; echo "nao rc=$?"; tail -2 gpurun_out/nao_tests.txt
timeout 300 python tools/nao_time.py > gpurun_out/nao_time.txt 2>&1; cat gpurun_out/nao_time.txt
timeout 900 python -u -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "c5_full_size" -s > gpurun_out/c5_tests.txt 2>&1; echo "c5 rc=$?"; grep -E "C5 edges|passed|failed|Error|assert" gpurun_out/c5_tests.txt | tail -5
timeout 300 python tools/ramp_times.py > gpurun_out/ramp_times.txt 2>&1; cat gpurun_out/ramp_times.txt
