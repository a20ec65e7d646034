"""Where do the tiled scan and the tree cross?  Device time (CUDA events, median of 5) of mptg_knn_query_dev for both
strategies over set sizes and wave sizes, for one space: python tools/knn_crossover.py l2_2d_f64|se3_f32|l2_3d_f32 [k]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import mpt_b200 as m
from tests.test_oracle import random_states

name = sys.argv[1] if len(sys.argv) > 1 else "l2_2d_f64"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
sp = {"l2_2d_f64": m.lp_space(2, 2, m.F64), "se3_f32": m.se3_space(50, 1), "l2_3d_f32": m.lp_space(3, 2, m.F32), "se3_f64": m.se3_space(50, 1, m.F64)}[name]
ctx = m.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
tdt = torch.float64 if sp.dtype == np.float64 else torch.float32
print(f"{name} k={k}: ms scan / ms tree")
for n in (1024, 2048, 4096, 8192, 16384, 65536):
    pts = random_states(sp, n, 1)
    row = []
    for Q in (256, 2048, 16384):
        q = torch.from_numpy(random_states(sp, Q, 2)).to(dev)
        di = torch.empty((Q, k), dtype=torch.int32, device=dev)
        dd = torch.empty((Q, k), dtype=tdt, device=dev)
        res = []
        for strat in (m.KNN_BRUTE, m.KNN_BVH):
            nn = m.Nearest(ctx, sp, n, strat)
            nn.insert(pts)
            ts = []
            for rep in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    e0.record(stream)
                    nn.nearest_dev(q.data_ptr(), Q, k, -1.0, di.data_ptr(), dd.data_ptr())
                    e1.record(stream)
                ctx.sync()
                if rep >= 3:
                    ts.append(e0.elapsed_time(e1))
            res.append(float(np.median(ts)))
            nn.close()
        row.append(f"Q={Q}: {res[0]:.3f} / {res[1]:.3f}")
    print(f"  N={n:6d}  " + "   ".join(row))
