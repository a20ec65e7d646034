#!/bin/bash
mkdir -p gpurun_out
{ echo "== default"; timeout 120 python tools/mesh_probe.py 9 2>&1 | tail -1
for v in "$@"; do echo "== $v"; MPTG_LIB=mpt_b200/_lib/variants/$v/libmptg.so timeout 120 python tools/mesh_probe.py 9 2>&1 | tail -1; done; } | tee -a gpurun_out/mesh_sweep2.txt
