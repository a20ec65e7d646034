"""Wall-clock time of every wave of a device-resident planner from a cold start (first-use costs show up here)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import mpt_b200 as m
from mpt_b200 import workloads as W
algo = sys.argv[1] if len(sys.argv) > 1 else "prrt"
wave = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
ctx = m.Context(0)
occ = W.synthetic_grid()
sp = m.lp_space(2, 2, m.F64)
grid = m.Scenario.grid(ctx, occ, m.F64)
free = np.argwhere(occ == 0)
start, goal = free[len(free) // 7][::-1].astype(np.float64), free[-len(free) // 9][::-1].astype(np.float64)
cls = {"prrt": m.DevicePRRT, "prrtstar": m.DevicePRRTStar}[algo]
pl = cls(grid, sp, [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1], range=200.0, goal=goal, goal_radius=12.0, seed=17, capacity=1 << 20, max_wave=wave)
pl.add_start(start)
ctx.sync()
out = []
for i in range(40):
    t = time.perf_counter()
    pl.wave(wave)
    ctx.sync()
    out.append(f"{pl.size}:{(time.perf_counter() - t) * 1e3:.2f}")
    print(f"-- wave {i} ended at size {pl.size}: {(time.perf_counter() - t) * 1e3:.2f} ms", file=sys.stderr, flush=True)
print(algo, " ".join(out))
