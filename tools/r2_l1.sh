#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "l1 or knn" > gpurun_out/l1_tests.txt 2>&1; echo "l1 rc=$?"; tail -3 gpurun_out/l1_tests.txt
timeout 300 python tools/knn_l1_scan_time.py > gpurun_out/l1_time.txt 2>&1
MPTG_KNN_L1_F64=1 timeout 300 python tools/knn_l1_scan_time.py > gpurun_out/l1_time_f64.txt 2>&1
echo "== mixed"; cat gpurun_out/l1_time.txt; echo "== all double"; cat gpurun_out/l1_time_f64.txt
timeout 600 python -u -m pytest tests/ -x -q -m gpu -k "pprm or planner or demo" > gpurun_out/pprm_tests.txt 2>&1; echo "pprm rc=$?"; tail -3 gpurun_out/pprm_tests.txt
