"""End-to-end time of mptg_knn_query on the C5 wave (host buffers in, host buffers out) for a given number of parts
(MPTG_KNN_HOST_PARTS in the environment; default = the library's own choice):  python tools/e2e_parts.py"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mpt_b200 as m  # noqa: E402
from mpt_b200 import workloads as W  # noqa: E402

N, Q, K = 1 << 20, 1 << 16, 16
ctx = m.Context(0)
sp = m.se3_space(50, 1)
nn = m.Nearest(ctx, sp, N)
nn.insert(W.se3_states(N, W.TREE_SEED))
nn.build_index()
q = W.se3_states(Q, W.QUERY_SEED)
for kind in ("pinned", "pageable"):
    if kind == "pinned":
        hq = torch.from_numpy(q).pin_memory()
        hi, hd, hc = (torch.empty((Q, K), dtype=torch.int32).pin_memory(), torch.empty((Q, K), dtype=torch.float32).pin_memory(),
                      torch.empty(Q, dtype=torch.int32).pin_memory())
        ptrs = (hq.data_ptr(), hi.data_ptr(), hd.data_ptr(), hc.data_ptr())
    else:
        pq, pi, pd, pc = q.copy(), np.empty((Q, K), np.uint32), np.empty((Q, K), np.float32), np.empty(Q, np.uint32)
        ptrs = (pq.ctypes.data, pi.ctypes.data, pd.ctypes.data, pc.ctypes.data)
    ts = []
    for it in range(12):
        ctx.sync()
        t0 = time.perf_counter()
        nn.nearest_host_into(ptrs[0], Q, K, -1.0, ptrs[1], ptrs[2], ptrs[3])
        ts.append((time.perf_counter() - t0) * 1e3)
    print(f"parts={os.environ.get('MPTG_KNN_HOST_PARTS', 'auto')} {kind}: {np.median(ts[3:]):.3f} ms (min {min(ts[3:]):.3f}) = {Q / np.median(ts[3:]) / 1e3:.1f} M queries/s")
ref = nn.nearest(q[:4096], K)
os.environ["X"] = "1"
print("checksum", int(ref[0].astype(np.uint64).sum()), float(ref[1].astype(np.float64).sum()))
