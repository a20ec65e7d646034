#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the planner tests: the one-CTA compaction and rewiring-tail kernels
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "prrtstar_queued or prrtstar_replays or prrt_replays" > gpurun_out/racecheck.txt 2>&1
echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed|Error" gpurun_out/racecheck.txt | head -20
