#!/bin/bash
# usage: r2_multi.sh N   -- sharded kNN tests and bench.py on N GPUs of one box
N=$1
mkdir -p gpurun_out
[ "$2" = "notest" ] || timeout 600 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu -s 2>&1 | tail -8
MPTG_COMM_TIMING=1 NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "bench rc=$?"; grep "mptg comm" gpurun_out/r2_bench_${N}gpu.err | head -3; tail -2 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','knn_ms','edge_ms','edges_per_s','n_gpus','scaling')}, d.get('replicated_tree'), d['e2e']['value'])
PY
