"""Wall clock of every wave of a ramped device-resident PRRT / PRRT* run on the 3976 x 2603 map (time to first solution):
    python tools/ramp_times.py            per-wave times
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ramp_launches.csv python tools/ramp_times.py once"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mpt_b200 as m  # noqa: E402
from mpt_b200 import workloads as W  # noqa: E402

once = len(sys.argv) > 1 and sys.argv[1] == "once"
cap = int(sys.argv[2]) if len(sys.argv) > 2 else 8192  # largest wave before the first solution
ctx = m.Context(0)
occ = W.synthetic_grid()
grid = m.Scenario.grid(ctx, occ, m.F64)
free = np.argwhere(occ == 0)
start = free[len(free) // 7][::-1].astype(np.float64)
goal = free[-len(free) // 9][::-1].astype(np.float64)
for name, cls in (("PRRT", m.DevicePRRT), ("PRRT*", m.DevicePRRTStar)):
    for attempt in range(1 if once else 3):
        pl = cls(grid, m.lp_space(2, 2, m.F64), [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1], range=200.0, goal=goal, goal_radius=12.0,
                 goal_bias=0.01, seed=17, capacity=1 << 18, max_wave=8192)
        pl.add_start(start)
        ctx.sync()
        t0, w, rows = time.perf_counter(), 64, []
        while not pl.solved() and pl.size < 200_000:
            t1 = time.perf_counter()
            pl.wave(w)
            rows.append((w, pl.size, (time.perf_counter() - t1) * 1e3))
            w = min(2 * w, cap)
        total = (time.perf_counter() - t0) * 1e3
        pl.close()
    print(f"{name} (waves up to {cap}): first solution after {total:.3f} ms, {len(rows)} waves, {rows[-1][1]} nodes")
    for w, size, ms in (rows if cap == 8192 else []):
        print(f"   wave of {w:5d} samples -> {size:6d} nodes  {ms:.3f} ms")
