#!/bin/bash
# compute-sanitizer memcheck over the GPU parity suite (smaller cases only: the full-size waves are skipped)
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 30 python -m pytest tests/test_gpu_parity.py tests/test_nao.py tests/test_reference_parity.py -x -q -m gpu -k "not c5_full_size and not fp32_probe" > gpurun_out/sanitize_all.txt 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|passed|failed|out of bounds|misaligned|Error" gpurun_out/sanitize_all.txt | head -20
