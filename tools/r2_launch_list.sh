#!/bin/bash
# ncu launch list of the bench command (main step only; cold-cache, serialised: compare shares, not absolutes)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-secondary > gpurun_out/r2_launches_bench.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2_launches_bench.log | cut -c1-300
python tools/summarize_launches.py gpurun_out/r2_launches.csv > gpurun_out/r2_launches_a.txt 2>&1; head -20 gpurun_out/r2_launches_a.txt
