"""Cost of rebuilding the kNN tree on a growing set (the planner's pattern): wall-clock of build_index() after each growth step."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import mpt_b200 as m
from tests.test_oracle import random_states
ctx = m.Context(0)
for name, sp in (("l2_2d_f64", m.lp_space(2, 2, m.F64)), ("se3_f32", m.se3_space(50, 1))):
    pts = random_states(sp, 400_000, 1)
    nn = m.Nearest(ctx, sp, 1 << 20, m.KNN_BVH)
    n, out = 0, []
    for step in (20_000, 30_000, 40_000, 60_000, 100_000, 150_000):
        nn.insert(pts[n:n + step]); n += step
        ctx.sync(); t = time.perf_counter(); nn.build_index(); ctx.sync(); a = time.perf_counter() - t
        ctx.sync(); t = time.perf_counter(); nn.build_index(); ctx.sync(); b = time.perf_counter() - t
        out.append(f"n={n}: {a*1e3:.2f} / {b*1e3:.2f} ms")
    print(name, "(first build at this size / repeated build):", "  ".join(out))
    nn.close()
