"""Device PPRM-IRS on the PNG-size grid up to ~60 K nodes, then ONE wave between cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mpt_b200 as m  # noqa: E402
from mpt_b200 import workloads as W  # noqa: E402

ctx = m.Context(0)
occ = W.synthetic_grid()
grid = m.Scenario.grid(ctx, occ, m.F64)
free = np.argwhere(occ == 0)
start, goal = free[len(free) // 7][::-1].astype(np.float64), free[-len(free) // 9][::-1].astype(np.float64)
pp = m.DevicePPRM(grid, m.lp_space(2, 2, m.F64), [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1], seed=23, capacity=1 << 18, max_wave=4096, spanner_stretch=5.0)
pp.add_start(start)
pp.add_goal(goal)
while pp.size < 60_000:
    pp.wave(4096)
ctx.sync()
torch.cuda.cudart().cudaProfilerStart()
pp.wave(4096)
ctx.sync()
torch.cuda.cudart().cudaProfilerStop()
print("nodes", pp.size)
