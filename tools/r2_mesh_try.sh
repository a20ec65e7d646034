#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mesh or contact" > gpurun_out/sanitize_mesh.txt 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_mesh.txt | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mesh_valid or mesh_link or contact" > gpurun_out/racecheck_mesh.txt 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/racecheck_mesh.txt | tail -3
