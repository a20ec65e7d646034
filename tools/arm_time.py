"""Link-arm edge waves (the bench's secondary workload, 65,536 edges, float64) timed on the device; MPTG_ARM_WARP_PER_EDGE=1
in the environment selects the earlier kernel (one warp per edge) instead of the flat edge check (geom.cu flatLink)."""
import os
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
for n_links in (8, 16, 32):
    lengths, radius, circles = W.link_arm_scene(n_links)
    arm = m.Scenario.link_arm(ctx, lengths, radius, circles, m.F64)
    for delta in (0.5, 0.1):
        a, b = W.arm_edges(65536, n_links, 41, delta)
        da, db = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        ok = torch.empty(a.shape[0], dtype=torch.uint8, device=dev)
        ts = []
        for it in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.sync()
            with torch.cuda.stream(stream):
                e0.record(stream)
                arm.link_dev(da.data_ptr(), db.data_ptr(), a.shape[0], ok.data_ptr())
                e1.record(stream)
            ctx.sync()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        print(f"warp_per_edge={os.environ.get('MPTG_ARM_WARP_PER_EDGE', '0')} links {n_links} delta {delta}: {ms:.3f} ms = {65536 / ms / 1e3:.1f} M edges/s, "
              f"probes {arm.last_stats()['prim_tests']}, valid {ok.float().mean().item():.3f}")
