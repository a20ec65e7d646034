#!/bin/bash
# One GPU call near a milestone: the whole GPU suite, smoke, the bench line (ours + reference arm).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/tests_all.txt 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests_all.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.txt
( time timeout 1200 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err ) 2> gpurun_out/bench_time.txt; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_1gpu.err; grep real gpurun_out/bench_time.txt
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"
