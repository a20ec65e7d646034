#!/bin/bash
# compute-sanitizer memcheck over the kernels added this round (flat edge check, Nao validator, spanner search)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_nao.py tests/test_gpu_parity.py -x -q -m gpu -k "flat_edge or reference_vectors or order_free or pprm_irs or linkarm_golden" > gpurun_out/sanitize.txt 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|passed|failed|out of bounds|misaligned" gpurun_out/sanitize.txt | head -20
