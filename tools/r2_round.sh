#!/bin/bash
# One GPU call near a milestone: the whole GPU suite, the bench line (ours + reference arm), the Nao captures.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/tests_all.txt 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/tests_all.txt
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_bench_1gpu.err
if [ "$1" = "ref" ]; then timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; fi
ncu --set full --clock-control none --import-source on -k regex:flatLinkKernel -c 1 -o gpurun_out/r2_nao_link_b -f python tools/nao_time.py once > gpurun_out/ncu_nao_log.txt 2>&1
tail -1 gpurun_out/ncu_nao_log.txt
timeout 120 demos/_build/planning_demos --demo nao_cup --time-ms 20000 --device-prrt --check > gpurun_out/nao_demo.txt 2>&1; echo "nao demo rc=$?"; grep -E "OK|FAILED" gpurun_out/nao_demo.txt
