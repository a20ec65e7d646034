"""SE(3) kNN wave time against the wave size (k = 16): where the persistent search stops being bound by throughput and
starts being bound by the latency of one query (the home pass of the sharded search, small planner waves)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mpt_b200 as m  # noqa: E402
from mpt_b200 import workloads as W  # noqa: E402

ctx = m.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
K = 16
for lg in (17, 20):
    nn = m.Nearest(ctx, m.se3_space(50.0, 1.0), 1 << lg, m.KNN_AUTO)
    nn.insert(W.se3_states(1 << lg, W.TREE_SEED))
    qall = W.se3_states(65536, W.QUERY_SEED)
    for Q in (256, 1024, 4096, 8192, 16384, 65536):
        dq = torch.from_numpy(qall[:Q].copy()).to(dev)
        di = torch.empty((Q, K), dtype=torch.int32, device=dev)
        dd = torch.empty((Q, K), dtype=torch.float32, device=dev)
        ts = []
        for it in range(9):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                nn.nearest_dev(dq.data_ptr(), Q, K, -1.0, di.data_ptr(), dd.data_ptr())
                e1.record(stream)
            ctx.sync()
            if it >= 3:
                ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        print(f"N=2^{lg} Q={Q:6d}: {ms:.3f} ms  {Q / ms / 1e3:.1f} M queries/s  ({ms * 1e3 / Q * 4736:.1f} us x resident warps / query)")
    nn.close()
