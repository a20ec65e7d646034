#!/bin/bash
# exact-link-count arm validators (state arrays in registers) against the run-time-count ones, then the arm parity tests
mkdir -p gpurun_out
for mx in 0 8 16 32; do echo "== MPTG_ARM_EXACT_MAX=$mx"; MPTG_ARM_EXACT_MAX=$mx timeout 300 python tools/arm_time.py 2>&1 | grep links; done | tee gpurun_out/arm_exact.txt
MPTG_ARM_EXACT_MAX=32 timeout 900 python -m pytest tests -x -q -m gpu -k "arm or link or flat" 2>&1 | tail -3
