#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err ) 2> gpurun_out/bench_time.txt; echo "bench rc=$?"; tail -c 400 gpurun_out/r2_bench_1gpu.err; cat gpurun_out/bench_time.txt
( time timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err ) 2>> gpurun_out/bench_time.txt; echo "ref rc=$?"; tail -4 gpurun_out/bench_time.txt
