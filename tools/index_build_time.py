import sys, time
sys.path.insert(0, '.')
import mpt_b200 as m
from mpt_b200 import workloads as W
ctx = m.Context(0)
sp = m.se3_space(50, 1)
for n in (1 << 16, 1 << 18, 1 << 20):
    pts = W.se3_states(n, 1)
    nn = m.Nearest(ctx, sp, n, m.KNN_BVH)
    nn.insert(pts)
    t = time.perf_counter(); nn.build_index(); ctx.sync(); dt = time.perf_counter() - t
    print(f"index build n={n}: {dt*1e3:.1f} ms")
    nn.close()
