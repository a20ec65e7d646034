#!/bin/bash
mkdir -p gpurun_out
for p in 1 2 4 8; do MPTG_KNN_HOST_PARTS=$p timeout 300 python tools/e2e_parts.py; done > gpurun_out/e2e_parts.txt 2>&1
cat gpurun_out/e2e_parts.txt
timeout 900 python -u -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "knn" > gpurun_out/knn_tests.txt 2>&1; echo "knn rc=$?"; tail -2 gpurun_out/knn_tests.txt
