"""Kernel experiment: C5 kNN wave (1M SE(3) points, 64K queries) -- device time per wave (CUDA events on the context
stream, _dev entry point, median of 7 after 3 warm-ups) and traversal counters.  Variants of the library are
selected with MPTG_LIB (tools/build_variant.sh); -DMPTG_KNN_PROBE adds counters on stderr."""
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
import mpt_b200 as m
from mpt_b200 import workloads as W

ctx = m.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
sp = m.se3_space(50, 1)
n, Q = 1 << 20, 1 << 16
ks = [int(a) for a in sys.argv[1:]] or [16]
pts = W.se3_states(n, 20261017, -100.0, 100.0)
qs = W.se3_states(Q, 20261018, -100.0, 100.0)
nn = m.Nearest(ctx, sp, n, m.KNN_BVH)
nn.insert(pts)
nn.build_index()
dq = torch.from_numpy(qs).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for k in ks:
    di = torch.empty((Q, k), dtype=torch.int32, device=dev)
    dd = torch.empty((Q, k), dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    ts = []
    for rep in range(10):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            nn.nearest_dev(dq.data_ptr(), Q, k, -1.0, di.data_ptr(), dd.data_ptr())
            e1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        if rep >= 3:
            ts.append(e0.elapsed_time(e1))
    st = nn.last_stats()
    print(f"k={k}: {np.median(ts):.3f} ms/wave (min {min(ts):.3f}); evals/query {st['distance_evals']/Q:.0f}, nodes/query {st['nodes_visited']/Q:.1f}; kth mean {dd[:, k-1].mean().item():.3f}")
