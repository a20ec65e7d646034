"""Where does a device-resident planner wave spend its time?  Grow a tree to --nodes nodes on the 3976x2603 synthetic map,
then run two waves between cudaProfilerStart/Stop:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/planner_wave_profile.py --algo prrtstar
and summarise with tools/summarize_launches.py.  Without ncu it prints wall-clock per wave."""
import argparse
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

import mpt_b200 as m
from mpt_b200 import workloads as W

ap = argparse.ArgumentParser()
ap.add_argument("--algo", default="prrtstar", choices=["prrt", "prrtstar", "pprm"])
ap.add_argument("--nodes", type=int, default=150_000)
ap.add_argument("--wave", type=int, default=8192)
ap.add_argument("--arm", type=int, default=0, help="N-link arm scene (PPRM) instead of the grid")
ap.add_argument("--se3", action="store_true", help="SE(3) rigid body among meshes (float32) instead of the grid")
args = ap.parse_args()

ctx = m.Context(0)
occ = W.synthetic_grid()
sp = m.lp_space(2, 2, m.F64)
grid = m.Scenario.grid(ctx, occ, m.F64)
free = np.argwhere(occ == 0)
start, goal = free[len(free) // 7][::-1].astype(np.float64), free[-len(free) // 9][::-1].astype(np.float64)
lo, hi = [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1]
if args.arm:
    lengths, radius, circles = W.link_arm_scene(args.arm)
    sp = m.lp_space(args.arm, 1, m.F64)
    grid = m.Scenario.link_arm(ctx, lengths, radius, circles, m.F64)
    cand = W.box_states(512, args.arm, 3, -np.pi, np.pi)
    okc = grid.valid(cand) != 0
    start, goal, lo, hi = cand[okc][0], cand[okc][1], -np.pi, np.pi
rng_ = 200.0
if args.se3:
    sp = m.se3_space(50, 1)
    robot, env, vmin, vmax = W.alpha_puzzle_like()
    grid = m.Scenario.mesh_pair(ctx, robot, env, sp, W.se3_step_size(vmin, vmax))
    cand = W.se3_states(256, 5, -45.0, 45.0)
    okc = grid.valid(cand) != 0
    start, goal = cand[okc][0], cand[okc][1]
    lo, hi, rng_ = [0, 0, 0, 0, -45, -45, -45], [0, 0, 0, 0, 45, 45, 45], 40.0
if args.algo == "prrt":
    pl = m.DevicePRRT(grid, sp, lo, hi, range=rng_, goal=goal, goal_radius=12.0, seed=17, capacity=1 << 20, max_wave=args.wave)
    pl.add_start(start)
elif args.algo == "prrtstar":
    pl = m.DevicePRRTStar(grid, sp, lo, hi, range=rng_, goal=goal, goal_radius=12.0, seed=17, capacity=1 << 20, max_wave=args.wave)
    pl.add_start(start)
else:
    pl = m.DevicePPRM(grid, sp, lo, hi, goal=goal, goal_radius=1e-6 if args.arm else 12.0, seed=17, capacity=1 << 18 if args.arm else 1 << 20, max_wave=args.wave)
    pl.add_start(start)
    pl.add_goal(goal)
while pl.size < args.nodes:
    pl.wave(args.wave)
ctx.sync()
torch.cuda.profiler.start()
t0, n0 = time.perf_counter(), pl.size
for _ in range(2):
    pl.wave(args.wave)
ctx.sync()
dt = time.perf_counter() - t0
torch.cuda.profiler.stop()
print(f"{args.algo}: 2 waves of {args.wave} samples at {n0} nodes: {dt * 1e3 / 2:.3f} ms per wave, {(pl.size - n0) / 2:.0f} nodes added per wave")
