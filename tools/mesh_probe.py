"""Time the C5 edge wave alone (mesh link) and print the traversal counters; for kernel experiments.
usage: python tools/mesh_probe.py [reps]"""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import bench as B
import mpt_b200 as m
from mpt_b200 import workloads as W

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
robot, env, vmin, vmax = W.alpha_puzzle_like(env_tris_target=4000, robot_tris_target=1000)
step = W.se3_step_size(vmin, vmax, B.SO3_W)
ea, eb = W.se3_edges(B.E_WAVE, W.EDGE_SEED, B.MESH_LO, B.MESH_HI, B.EDGE_TRANS, B.EDGE_ANGLE)
ctx = m.Context(0)
sp = m.se3_space(B.SO3_W, B.L2_W)
mesh = m.Scenario.mesh_pair(ctx, robot, env, sp, step)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
d_ea, d_eb = torch.from_numpy(ea).to(dev), torch.from_numpy(eb).to(dev)
d_ok = torch.empty(B.E_WAVE, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
ts = []
for r in range(reps + 2):
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        mesh.link_dev(d_ea.data_ptr(), d_eb.data_ptr(), B.E_WAVE, d_ok.data_ptr())
        e1.record(stream)
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
st = mesh.last_stats()
ok = d_ok.cpu().numpy()
print(f"edge wave: {np.median(ts[2:]):.3f} ms (min {min(ts[2:]):.3f}); valid {ok.mean():.4f}; {st}")
