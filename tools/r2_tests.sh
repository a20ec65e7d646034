#!/bin/bash
# usage: r2_tests.sh "<pytest -k expression>" [timeout seconds]
mkdir -p gpurun_out
timeout ${2:-900} python -u -m pytest tests/ -x -v -m gpu -k "$1" -s > gpurun_out/tests.txt 2>&1
echo "rc=$?" >> gpurun_out/tests.txt
grep -E "FAILED|ERROR|contact|first contact|as edges|passed|failed|rc=|Error|assert" gpurun_out/tests.txt | cut -c1-250 | tail -40
