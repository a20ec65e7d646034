#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pprm" -s > gpurun_out/irs_tests.txt 2>&1; echo "irs rc=$?"; grep -E "PPRM-IRS device|passed|failed|Error|assert" gpurun_out/irs_tests.txt | tail -8
timeout 600 python -u -m pytest tests/test_host_cpp.py -x -q -m gpu -k "wave_planners_on_gpu" -s > gpurun_out/host_tests.txt 2>&1; echo "host rc=$?"; grep -E "IRS|passed|failed|FAIL" gpurun_out/host_tests.txt | tail -8
