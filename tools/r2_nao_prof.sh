#!/bin/bash
# ncu captures of the Nao kernels (float): the flat edge check and the state validator
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:flatLinkKernel -c 1 -o gpurun_out/r2_nao_link -f python tools/nao_time.py once > gpurun_out/ncu_nao_log.txt 2>&1
tail -2 gpurun_out/ncu_nao_log.txt
ncu --set full --clock-control none --import-source on -k regex:validKernel -c 1 -o gpurun_out/r2_nao_valid -f python tools/nao_time.py once > gpurun_out/ncu_nao_log2.txt 2>&1
tail -2 gpurun_out/ncu_nao_log2.txt
