#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for c in 128 512; do MPTG_STAR_TIMING=1 timeout 120 python tools/ramp_times.py x $c; done 2>&1 | tee gpurun_out/ramp_queued.txt
