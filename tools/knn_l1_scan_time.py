"""Time of the arm-space kNN scan (one unweighted L1 part) for a 64 K-query wave over 128 K points, float64 and float32.
Usage (GPU box): python tools/knn_l1_scan_time.py [Q]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mpt_b200 as m  # noqa: E402
from mpt_b200 import workloads as W  # noqa: E402

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ctx = m.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
for scalar, dt, tdt in ((m.F64, np.float64, torch.float64), (m.F32, np.float32, torch.float32)):
    for dim in (8, 16, 32):
        for k in (16, 37):
            pts = W.box_states(1 << 17, dim, 61, -np.pi, np.pi, dt)
            q = W.box_states(Q, dim, 62, -np.pi, np.pi, dt)
            nn = m.Nearest(ctx, m.lp_space(dim, 1, scalar), 1 << 17, m.KNN_BRUTE)
            nn.insert(pts)
            dq = torch.from_numpy(q).to(dev)
            di = torch.empty((Q, k), dtype=torch.int32, device=dev)
            dd = torch.empty((Q, k), dtype=tdt, device=dev)
            ts = []
            for it in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    e0.record(stream)
                    nn.nearest_dev(dq.data_ptr(), Q, k, -1.0, di.data_ptr(), dd.data_ptr())
                    e1.record(stream)
                ctx.sync()
                if it >= 2:
                    ts.append(e0.elapsed_time(e1))
            nn.close()
            ms = float(np.mean(ts))
            print(f"{'f64' if scalar == m.F64 else 'f32'} D={dim:2d} k={k:2d} Q={Q}: {ms:8.3f} ms  {Q / ms / 1e3:7.2f} M queries/s  "
                  f"{Q * (1 << 17) * dim / ms / 1e9:6.2f} T coordinate pairs/s", flush=True)
