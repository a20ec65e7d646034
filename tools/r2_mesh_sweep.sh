#!/bin/bash
mkdir -p gpurun_out
{
echo "== default (16 warps)"; python tools/mesh_probe.py 7
for v in "$@"; do echo "== variant $v"; MPTG_LIB=mpt_b200/_lib/variants/$v/libmptg.so timeout 120 python tools/mesh_probe.py 7; done
} > gpurun_out/mesh_sweep2.txt 2>&1
cat gpurun_out/mesh_sweep2.txt
