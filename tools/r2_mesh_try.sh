#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/mesh_probe.py 9 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_parity.py -x -q -m gpu -k "mesh or dmv or c5_full_size_edge or contact or se3" 2>&1 | tail -3
