#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests/test_nao.py -x -q -m gpu > gpurun_out/nao_tests.txt 2>&1; echo "nao rc=$?"; tail -3 gpurun_out/nao_tests.txt
timeout 900 python -u -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "arm or l1_double" > gpurun_out/arm_tests.txt 2>&1; echo "arm rc=$?"; tail -3 gpurun_out/arm_tests.txt
