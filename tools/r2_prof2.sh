#!/bin/bash
mkdir -p gpurun_out
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:pprmSpannerKernel -c 1 -o gpurun_out/r2_spanner -f python tools/irs_once.py > gpurun_out/ncu_spanner_log.txt 2>&1; tail -1 gpurun_out/ncu_spanner_log.txt
ncu --set full --clock-control none --import-source on -k regex:flatLinkKernel -s 1 -c 1 -o gpurun_out/r2_arm_flat -f python tools/arm_link_once.py 16 > gpurun_out/ncu_arm_log.txt 2>&1; tail -1 gpurun_out/ncu_arm_log.txt
