"""One wave of the device-resident PPRM on the N-link arm at ~100 K nodes, bracketed by cudaProfilerStart/Stop:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/pprm_wave_profile.py [links]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mpt_b200 as m  # noqa: E402
from mpt_b200 import workloads as W  # noqa: E402

n_links = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = m.Context(0)
lengths, radius, circles = W.link_arm_scene(n_links)
arm = m.Scenario.link_arm(ctx, lengths, radius, circles, m.F64)
cand = W.box_states(512, n_links, 3, -np.pi, np.pi)
ok = arm.valid(cand) != 0
pp = m.DevicePPRM(arm, m.lp_space(n_links, 1, m.F64), -np.pi, np.pi, seed=23, capacity=1 << 18, max_wave=4096)
pp.add_start(cand[ok][0])
pp.add_goal(cand[ok][1])
while pp.size < 100_000:
    pp.wave(4096)
ctx.sync()
torch.cuda.cudart().cudaProfilerStart()
pp.wave(4096)
ctx.sync()
torch.cuda.cudart().cudaProfilerStop()
print("nodes", pp.size)
