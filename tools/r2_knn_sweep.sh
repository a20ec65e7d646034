#!/bin/bash
# One GPU call: kNN parity tests, C5 wave time for split-rule variants, probe counters, ncu capture of the tree kernel.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "knn" > gpurun_out/knn_tests.txt 2>&1
tail -5 gpurun_out/knn_tests.txt
{
for rs in 0.7 1.0 1.5; do echo "== rotSplit $rs"; MPTG_KNN_ROT_SPLIT=$rs python tools/knn_probe.py 16; done
echo "== default, k sweep"; python tools/knn_probe.py 1 4 16 48 100
if [ -f mpt_b200/_lib/variants/probe/libmptg.so ]; then echo "== probe counters"; MPTG_LIB=mpt_b200/_lib/variants/probe/libmptg.so python tools/knn_probe.py 16; fi
for v in "$@"; do echo "== variant $v"; MPTG_LIB=mpt_b200/_lib/variants/$v/libmptg.so python tools/knn_probe.py 16; done
} > gpurun_out/knn_sweep.txt 2>&1
cat gpurun_out/knn_sweep.txt
ncu --set full --clock-control none --import-source on -k regex:knnBvhKernel -c 1 -o gpurun_out/r2_knn_cap -f python tools/knn_probe.py 16 > gpurun_out/ncu_log.txt 2>&1
tail -3 gpurun_out/ncu_log.txt
