#!/bin/bash
# One GPU call: mesh parity tests (incl. the contact cases), edge-wave time for CTA-shape variants, ncu capture.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mesh or c5_full_size_edge or planner or prrt or pprm" -s > gpurun_out/mesh_tests.txt 2>&1
grep -E "contact|first contact|as edges|passed|failed|Error|error" gpurun_out/mesh_tests.txt | tail -30
{
echo "== default"; python tools/mesh_probe.py 7
for v in "$@"; do echo "== variant $v"; MPTG_LIB=mpt_b200/_lib/variants/$v/libmptg.so python tools/mesh_probe.py 7; done
} > gpurun_out/mesh_sweep.txt 2>&1
cat gpurun_out/mesh_sweep.txt
ncu --set full --clock-control none --import-source on -k regex:meshFlatKernel -c 1 -o gpurun_out/r2_mesh_flat -f python tools/mesh_probe.py 1 > gpurun_out/ncu_mesh_log.txt 2>&1
tail -2 gpurun_out/ncu_mesh_log.txt
