"""One wave of link-arm edges (the bench's secondary workload) for an ncu capture of bisectLinkKernel<ArmValidator>:
    ncu --set full --import-source on --clock-control none -k regex:bisectLinkKernel -s 2 -c 1 -o gpurun_out/prof_arm python tools/arm_link_once.py [n_links]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mpt_b200 as m  # noqa: E402
from mpt_b200 import workloads as W  # noqa: E402

n_links = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = m.Context(0)
lengths, radius, circles = W.link_arm_scene(n_links)
arm = m.Scenario.link_arm(ctx, lengths, radius, circles, m.F64)
a, b = W.arm_edges(65536, n_links, 41, 0.5)
for _ in range(4):
    ok = arm.link(a, b)
print("valid fraction", float(np.mean(ok != 0)), arm.last_stats())
