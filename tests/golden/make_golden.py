#!/usr/bin/env python3
"""Generate tests/golden/golden.npz: small seeded inputs with the CPU oracle's outputs.

The reference (UNC-Robotics/mpt) is header-only C++ that cannot be built in this environment
(Eigen, Nigh, FCL, libpng absent -- SURVEY.md section 8c), so these vectors come from the oracle,
whose arithmetic is pinned by the reference's known-answer tests (tests/kats.py, oracle/kat_main.cpp).
They freeze the oracle's behaviour: the CPU tests re-check the oracle against them, the GPU tests
check the CUDA kernels against them.  Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

import mpt_b200 as m  # noqa: E402
from mpt_b200 import workloads as W  # noqa: E402
from tests import oracle_binding  # noqa: E402


def main():
    o = oracle_binding.load()
    out = {}
    sp = m.se3_space(50, 1)
    out["se3_pts"] = W.se3_states(2048, W.TREE_SEED)
    out["se3_q"] = W.se3_states(128, W.QUERY_SEED)
    out["se3_knn_idx"], out["se3_knn_dist"], _ = o.knn(sp, out["se3_pts"], out["se3_q"], 16)
    l1 = m.lp_space(8, 1)
    out["l1_pts"] = W.box_states(1024, 8, 3, -np.pi, np.pi, np.float32)
    out["l1_q"] = W.box_states(64, 8, 4, -np.pi, np.pi, np.float32)
    out["l1_knn_idx"], out["l1_knn_dist"], _ = o.knn(l1, out["l1_pts"], out["l1_q"], 5)
    out["interp_t"] = np.linspace(0.0, 1.0, 32).astype(np.float32)
    out["se3_interp"] = o.interpolate(sp, out["se3_q"][:32], out["se3_pts"][:32], out["interp_t"])

    occ = W.synthetic_grid(400, 300, seed=3, n_blobs=125)
    out["grid_occ"] = occ
    out["grid_a"], out["grid_b"] = W.grid_edges(512, 400, 300, 5, 25.0)
    out["grid_link"] = o.grid(occ).link(out["grid_a"], out["grid_b"])

    lengths, radius, circles = W.link_arm_scene(8)
    out["arm_lengths"], out["arm_radius"], out["arm_circles"] = lengths, np.float64(radius), circles
    out["arm_a"], out["arm_b"] = W.arm_edges(256, 8, 9)
    out["arm_link"] = o.link_arm(lengths, radius, circles).link(out["arm_a"], out["arm_b"])

    robot, env, vmin, vmax = W.alpha_puzzle_like(env_tris_target=600, robot_tris_target=200)
    step = W.se3_step_size(vmin, vmax)
    out["mesh_robot"], out["mesh_env"], out["mesh_step"] = robot, env, np.float64(step)
    mesh = o.mesh_pair(robot, env, sp, step)
    out["mesh_states"] = W.se3_states(512, 21, -45.0, 45.0)
    out["mesh_valid"] = mesh.valid(out["mesh_states"])
    out["mesh_a"], out["mesh_b"] = W.se3_edges(256, 23, -45.0, 45.0, 12.0, 0.5)
    out["mesh_link"], nc = mesh.link(out["mesh_a"], out["mesh_b"], with_near_contact=True)
    out["mesh_link_near_contact"] = nc
    path = Path(__file__).with_name("golden.npz")
    np.savez_compressed(path, **out)
    print(path, {k: (v.shape, str(v.dtype)) for k, v in out.items() if hasattr(v, "shape")})
    print("grid link valid fraction", out["grid_link"].mean(), "arm", out["arm_link"].mean(), "mesh valid",
          out["mesh_valid"].mean(), "mesh link", out["mesh_link"].mean(), "near contact edges", int(nc.sum()))


if __name__ == "__main__":
    main()
