"""Device-resident PPRM on the 8- / 16-link arm to 150 K nodes (the bench's secondary figure), for kernel comparisons:
MPTG_ARM_WARP_PER_EDGE=1 selects the earlier edge kernel."""
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mpt_b200 as m  # noqa: E402
from mpt_b200 import workloads as W  # noqa: E402

ctx = m.Context(0)
for n_links in (8, 16):
    lengths, radius, circles = W.link_arm_scene(n_links)
    spn = m.lp_space(n_links, 1, m.F64)
    arm = m.Scenario.link_arm(ctx, lengths, radius, circles, m.F64)
    cand = W.box_states(512, n_links, 3, -np.pi, np.pi)
    ok = arm.valid(cand) != 0
    for attempt in range(2):
        pp = m.DevicePPRM(arm, spn, -np.pi, np.pi, seed=23, capacity=1 << 18, max_wave=4096)
        pp.add_start(cand[ok][0])
        pp.add_goal(cand[ok][1])
        pp.wave(4096)
        ctx.sync()
        t0, n0 = time.perf_counter(), pp.size
        while pp.size < 150_000:
            pp.wave(4096)
        dt = time.perf_counter() - t0
        print(f"warp_per_edge={os.environ.get('MPTG_ARM_WARP_PER_EDGE', '0')} links {n_links}: {(pp.size - n0) / dt / 1e3:.1f} K nodes/s ({dt * 1e3:.1f} ms)", flush=True)
        pp.close()
