#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -u -m pytest tests/test_nao.py tests/test_gpu_parity.py tests/test_reference_parity.py -x -q -m gpu -k "nao or arm or flat or pprm or scenarios or device_matches" > gpurun_out/flat_tests.txt 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/flat_tests.txt
timeout 300 python tools/nao_time.py 2>&1 | grep link
timeout 300 python tools/arm_time.py 2>&1
timeout 300 python tools/pprm_arm_time.py 2>&1
