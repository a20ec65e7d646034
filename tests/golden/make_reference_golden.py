#!/usr/bin/env python3
"""Generate tests/golden/reference_golden.npz: outputs of the REFERENCE'S OWN code (its headers compiled from
/root/reference against the stand-in Eigen/Nigh headers, oracle/ref_driver.cpp) on seeded inputs.
Can only run where /root/reference exists; the vectors are committed so that the CPU suite and the GPU box
(which has no /root/reference) check the oracle and the CUDA kernels against the reference's decisions.
Run from the repo root:  python tests/golden/make_reference_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import mpt_b200 as m  # noqa: E402
from mpt_b200 import workloads as W  # noqa: E402
from tests import oracle_binding, reference_binding  # noqa: E402
from tests.test_oracle import random_states  # noqa: E402

HOLO_CIRCLES = [[170, 140, 80], [800, 70, 50], [900, 380, 70]]                      # demo/holonomic_2d_point_planning.cpp:71-73
HOLO_RECTS = [[375, 140, 520, 220], [200, 320, 390, 390], [600, 200, 680, 450]]    # :74-76
ARM5 = ([10.0, 12.0, 8.0, 6.0, 4.0], 0.5, [[20, -20, 8], [-20, -30, 5], [0, 25, 10], [30, 10, 10], [-30, 10, 8]])  # link_manipulator_planning.cpp:62-76


def main():
    ref = reference_binding.load()
    orc = oracle_binding.load()
    out = {}
    rng = np.random.default_rng(77)
    # interpolate
    a3, b3 = rng.random((256, 3)) * 20 - 10, rng.random((256, 3)) * 20 - 10
    t = rng.random(256)
    out["l2_a"], out["l2_b"], out["l2_t"], out["l2_out"] = a3, b3, t, ref.interpolate("l2_3", a3, b3, t)
    sa, sb = rng.random(256) * 2 * np.pi - np.pi, rng.random(256) * 2 * np.pi - np.pi
    out["so2_a"], out["so2_b"], out["so2_out"] = sa, sb, ref.interpolate("so2", sa, sb, t)
    se3 = m.se3_space(50, 1)
    ea, eb = random_states(se3, 256, 5), random_states(se3, 256, 6)
    out["se3_a"], out["se3_b"], out["se3_t"] = ea, eb, t.astype(np.float32)
    out["se3_out_f32"] = ref.interpolate("se3_f32", ea, eb, t.astype(np.float32))
    out["se3_out_f64"] = ref.interpolate("se3_f64", ea.astype(np.float64), eb.astype(np.float64), t)
    # grid (in-range samples: x + 0.5 < width, y + 0.5 < height; the reference reads out of bounds beyond)
    occ = W.synthetic_grid(600, 400, seed=8)
    ga, gb = W.grid_edges(4096, 599, 399, 12, 50.0)
    out["grid_occ"], out["grid_a"], out["grid_b"] = occ, ga, gb
    out["grid_valid"], out["grid_link"] = ref.grid(occ, ga, gb)
    # holonomic scene as shipped
    ha, hb = W.grid_edges(4096, 1024, 512, 14, 150.0)
    out["holo_a"], out["holo_b"] = ha, hb
    out["holo_valid"], out["holo_link"] = ref.holonomic(HOLO_CIRCLES, HOLO_RECTS, ha, hb)
    # link arms: the shipped 5-link demo and the 8/16/32-link scenes of the workload
    for n_links in (5, 8, 16, 32):
        lengths, radius, circles = ARM5 if n_links == 5 else W.link_arm_scene(n_links)
        aa, ab = W.arm_edges(1024, n_links, 20 + n_links, 0.5)
        out[f"arm{n_links}_a"], out[f"arm{n_links}_b"] = aa, ab
        out[f"arm{n_links}_valid"], out[f"arm{n_links}_link"] = ref.linkarm(lengths, radius, circles, aa, ab)
    # DiscreteMotionValidator (the reference's loop) around the oracle's mesh state validator
    robot, env, vmin, vmax = W.alpha_puzzle_like(env_tris_target=600, robot_tris_target=200)
    step = W.se3_step_size(vmin, vmax)
    mesh = orc.mesh_pair(robot, env, se3, step)
    da, db = W.se3_edges(384, 31, -45.0, 45.0, 25.0, 1.0)
    ok, cnt = ref.dmv_se3(da, db, step, lambda q: mesh.valid(q[None, :])[0])
    out["dmv_robot"], out["dmv_env"], out["dmv_step"] = robot, env, np.float64(step)
    out["dmv_a"], out["dmv_b"], out["dmv_ok"], out["dmv_states"] = da, db, ok, cnt
    # GoalState
    gq = rng.random((64, 3)) * 2 - 1
    gq[0] = [0.25, 0.25, 0.25]
    out["goal_q"] = gq
    out["goal_is"], out["goal_dist"] = ref.goal_l2_3([0.25, 0.25, 0.25], 1e-6, gq)
    # the reference's own PRRT planner (oracle/ref_planner.cpp) on a synthetic occupancy grid, replayed sample stream
    from tests.kats import PRRT_CASES, prrt_scene
    occ, lo, hi, start, goal = prrt_scene()
    for i, (rng_, n, goal_radius, goal_bias, seed) in enumerate(PRRT_CASES):
        u = orc.sample_uniforms(m.F64, seed, 0, n, 3)
        st, par, gn = ref.prrt_grid(occ, lo, hi, start, goal, goal_radius, goal_bias, rng_, u)
        out[f"prrt{i}_states"], out[f"prrt{i}_parents"], out[f"prrt{i}_goal"] = st, par, np.uint32(gn)
        print(f"prrt case {i}: {n} samples -> {st.shape[0]} nodes, goal node {gn}")
    path = Path(__file__).with_name("reference_golden.npz")
    np.savez_compressed(path, **out)
    print(path, "grid link", out["grid_link"].mean(), "holo link", out["holo_link"].mean(), "arm5 link", out["arm5_link"].mean(),
          "dmv ok", ok.mean(), "dmv states", int(cnt.sum()))


if __name__ == "__main__":
    main()
