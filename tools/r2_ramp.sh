#!/bin/bash
mkdir -p gpurun_out
for cap in 512 1024 2048 4096 8192; do timeout 200 python tools/ramp_times.py all $cap; done > gpurun_out/ramp_caps.txt 2>&1
cat gpurun_out/ramp_caps.txt | grep -v "^   wave"
