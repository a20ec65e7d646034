"""Synthetic planning workloads: inputs for tests and bench.py.

Everything here is input generation on the host (numpy); no algorithm of the hot path lives here.
Seeds are fixed so the CPU oracle and the GPU see identical inputs.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

TREE_SEED = 20261017
QUERY_SEED = 20261018
EDGE_SEED = 20261019


def so3_uniform(rng: np.random.Generator, n: int, dtype=np.float32) -> np.ndarray:
    """The reference's SO(3) sampler formula (src/mpt/impl/uniform_sampler_so3.hpp:57-67):
    a~U[0,1), b,c~U[0,2pi) -> (w,x,y,z) = (sqrt(1-a) sin b, sqrt(1-a) cos b, sqrt(a) sin c, sqrt(a) cos c);
    returned in coeff order (x,y,z,w)."""
    a = rng.random(n)
    b = rng.random(n) * 2 * np.pi
    c = rng.random(n) * 2 * np.pi
    w = np.sqrt(1 - a) * np.sin(b)
    x = np.sqrt(1 - a) * np.cos(b)
    y = np.sqrt(a) * np.sin(c)
    z = np.sqrt(a) * np.cos(c)
    return np.stack([x, y, z, w], axis=1).astype(dtype)


def se3_states(n: int, seed: int, lo: float = -100.0, hi: float = 100.0, dtype=np.float32) -> np.ndarray:
    """n SE(3) states [qx qy qz qw tx ty tz], translations ~U[lo,hi)^3 (C5 of BASELINE.json)."""
    rng = np.random.default_rng(seed)
    q = so3_uniform(rng, n, dtype)
    t = (rng.random((n, 3)) * (hi - lo) + lo).astype(dtype)
    return np.ascontiguousarray(np.concatenate([q, t], axis=1))


def box_states(n: int, dim: int, seed: int, lo, hi, dtype=np.float64) -> np.ndarray:
    """UniformBoxSampler (src/mpt/uniform_box_sampler.hpp:60-68): per-coordinate uniform in [lo,hi)."""
    rng = np.random.default_rng(seed)
    lo = np.broadcast_to(np.asarray(lo, dtype=np.float64), (dim,))
    hi = np.broadcast_to(np.asarray(hi, dtype=np.float64), (dim,))
    return np.ascontiguousarray((rng.random((n, dim)) * (hi - lo) + lo).astype(dtype))


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz,
    ], axis=-1)


def se3_edges(n: int, seed: int, lo: float, hi: float, max_trans: float, max_angle: float, dtype=np.float32):
    """n SE(3) edges (from, to): `from` uniform, `to` = from perturbed by a translation of length
    <= max_trans and a rotation of angle <= max_angle (a steered sample)."""
    rng = np.random.default_rng(seed)
    frm = se3_states(n, seed + 1, lo, hi, np.float64)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d *= (rng.random((n, 1)) ** (1 / 3)) * max_trans
    axis = rng.normal(size=(n, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    ang = rng.random(n) * max_angle
    dq = np.concatenate([axis * np.sin(ang / 2)[:, None], np.cos(ang / 2)[:, None]], axis=1)
    to = frm.copy()
    to[:, :4] = quat_mul(dq, frm[:, :4])
    to[:, 4:] = np.clip(frm[:, 4:] + d, lo, hi)
    return np.ascontiguousarray(frm.astype(dtype)), np.ascontiguousarray(to.astype(dtype))


# ---------------------------------------------------------------------------------- meshes
def tube_mesh(path_pts: np.ndarray, radius: float, sides: int) -> np.ndarray:
    """Triangle soup [n,3,3] of a closed tube swept along a polyline (parallel-transport frames)."""
    p = np.asarray(path_pts, dtype=np.float64)
    tang = np.gradient(p, axis=0)
    tang /= np.linalg.norm(tang, axis=1, keepdims=True)
    up = np.array([0.0, 0.0, 1.0])
    if abs(tang[0] @ up) > 0.9:
        up = np.array([1.0, 0.0, 0.0])
    n0 = np.cross(tang[0], up)
    n0 /= np.linalg.norm(n0)
    rings = []
    nrm = n0
    for i in range(len(p)):
        nrm = nrm - (nrm @ tang[i]) * tang[i]
        nrm /= np.linalg.norm(nrm)
        b = np.cross(tang[i], nrm)
        ang = np.arange(sides) * 2 * np.pi / sides
        rings.append(p[i] + radius * (np.cos(ang)[:, None] * nrm + np.sin(ang)[:, None] * b))
    rings = np.array(rings)
    tris = []
    for i in range(len(p) - 1):
        for j in range(sides):
            a, b2 = rings[i, j], rings[i, (j + 1) % sides]
            c, d = rings[i + 1, j], rings[i + 1, (j + 1) % sides]
            tris.append([a, b2, c])
            tris.append([b2, d, c])
    for ring, centre, flip in ((rings[0], p[0], True), (rings[-1], p[-1], False)):
        for j in range(sides):
            t = [centre, ring[j], ring[(j + 1) % sides]]
            tris.append(t[::-1] if flip else t)
    return np.asarray(tris, dtype=np.float32)


def alpha_puzzle_like(seed: int = 7, env_tris_target: int = 4000, robot_tris_target: int = 1000):
    """Two bent-tube meshes in the spirit of OMPL's alpha puzzle (the real meshes are not in the
    reference repository): the environment is a large twisted loop, the robot a smaller one.
    Returns (robot_tris [nr,3,3] recentred on the vertex mean as the reference does at
    demo/se3_rigid_body_scenario.hpp:181-193, env_tris [ne,3,3], volume_min, volume_max)."""
    rng = np.random.default_rng(seed)

    def loop(scale, wobble, nseg, phase):
        s = np.linspace(0, 1.6 * np.pi, nseg)
        pts = np.stack([
            scale * np.cos(s),
            scale * np.sin(s) * (1 + 0.2 * np.sin(3 * s + phase)),
            wobble * scale * np.sin(2 * s + phase),
        ], axis=1)
        return pts

    sides_e = 16
    nseg_e = max(8, env_tris_target // (2 * sides_e))
    env = tube_mesh(loop(30.0, 0.35, nseg_e, 0.3), 4.0, sides_e)
    sides_r = 10
    nseg_r = max(6, robot_tris_target // (2 * sides_r))
    robot = tube_mesh(loop(14.0, 0.45, nseg_r, 1.1 + rng.random()), 2.0, sides_r)
    centre = robot.reshape(-1, 3).mean(axis=0)
    robot = (robot - centre).astype(np.float32)
    vmin = np.array([-60.0, -60.0, -40.0])
    vmax = np.array([60.0, 60.0, 40.0])
    return robot, env.astype(np.float32), vmin, vmax


def se3_step_size(vmin, vmax, so3_weight: float = 50.0, resolution: float = 0.01) -> float:
    """stepSize of the SE(3) demo: (|max-min| + so3_weight*pi/2) * resolution
    (demo/se3_rigid_body_scenario.hpp:333, resolution 0.01 at se3_rigid_body_planning.cpp:172)."""
    return float((np.linalg.norm(np.asarray(vmax) - np.asarray(vmin)) + so3_weight * np.pi / 2) * resolution)


# ---------------------------------------------------------------------------------- occupancy grid
PNG_FILTERS = ((126, 106, 61, 15), (61, 53, 6, 15), (255, 255, 255, 5))  # demo/png_2d_planning.cpp:69-72


def filter_png(rgb: np.ndarray, filters=PNG_FILTERS) -> np.ndarray:
    """FilterColor::isObstacle over an [h,w,3] uint8 image (demo/png_2d_scenario.hpp:50-69,246-265)."""
    img = rgb.astype(np.int32)
    occ = np.zeros(img.shape[:2], dtype=bool)
    for r, g, b, tol in filters:
        inside = (
            (img[..., 0] >= r - tol) & (img[..., 0] <= r + tol)
            & (img[..., 1] >= g - tol) & (img[..., 1] <= g + tol)
            & (img[..., 2] >= b - tol) & (img[..., 2] <= b + tol)
        )
        occ |= inside
    return occ.astype(np.uint8)


def synthetic_grid(width: int = 3976, height: int = 2603, seed: int = 11, n_blobs: int = 125) -> np.ndarray:
    """A map-like occupancy grid of the shipped PNG's size (3976x2603, about one third occupied),
    used where the reference image is not available (the GPU box has no /root/reference)."""
    rng = np.random.default_rng(seed)
    occ = np.zeros((height, width), dtype=np.uint8)
    yy, xx = np.mgrid[0:height, 0:width]
    sc = width / 3976.0  # blob sizes are quoted for the full-size map
    for _ in range(n_blobs):
        cx, cy = rng.random() * width, rng.random() * height
        if rng.random() < 0.5:
            r = (30 + rng.random() * 170) * sc
            x0, x1 = int(max(0, cx - r)), int(min(width, cx + r))
            y0, y1 = int(max(0, cy - r)), int(min(height, cy + r))
            sub = (xx[y0:y1, x0:x1] - cx) ** 2 + (yy[y0:y1, x0:x1] - cy) ** 2 <= r * r
            occ[y0:y1, x0:x1] |= sub.astype(np.uint8)
        else:
            w, h = (40 + rng.random() * 400) * sc, (20 + rng.random() * 200) * sc
            x0, x1 = int(max(0, cx - w / 2)), int(min(width, cx + w / 2))
            y0, y1 = int(max(0, cy - h / 2)), int(min(height, cy + h / 2))
            occ[y0:y1, x0:x1] = 1
    return occ


def load_reference_png(path: Path) -> np.ndarray:
    from PIL import Image

    img = np.asarray(Image.open(path).convert("RGB"))
    return filter_png(img)


def grid_edges(n: int, width: int, height: int, seed: int, max_len: float | None, dtype=np.float64):
    """Edges between uniform samples of [0,w]x[0,h] (the reference's bounds, png_2d_scenario.hpp:144-150);
    max_len=None keeps them as drawn (no range set, as shipped), otherwise `to` is pulled in to max_len."""
    a = box_states(n, 2, seed, 0.0, [width, height], np.float64)
    b = box_states(n, 2, seed + 1, 0.0, [width, height], np.float64)
    if max_len is not None:
        d = b - a
        ln = np.linalg.norm(d, axis=1, keepdims=True)
        s = np.minimum(1.0, max_len / np.maximum(ln, 1e-12))
        b = a + d * s
    return np.ascontiguousarray(a.astype(dtype)), np.ascontiguousarray(b.astype(dtype))


# ---------------------------------------------------------------------------------- link arm
def link_arm_scene(n_links: int, n_circles: int = 8, seed: int = 5):
    """N links of length 4 among circles of radius 3 placed away from the base (C4)."""
    rng = np.random.default_rng(seed)
    lengths = np.full(n_links, 4.0)
    reach = lengths.sum()
    circles = []
    while len(circles) < n_circles:
        r = reach * (0.35 + 0.6 * rng.random())
        a = rng.random() * 2 * np.pi
        circles.append((r * np.cos(a), r * np.sin(a), 3.0))
    return lengths, 0.5, np.asarray(circles)


def arm_edges(n: int, n_links: int, seed: int, max_delta: float = 0.5, dtype=np.float64):
    a = box_states(n, n_links, seed, -np.pi, np.pi, np.float64)
    rng = np.random.default_rng(seed + 1)
    b = np.clip(a + (rng.random((n, n_links)) * 2 - 1) * max_delta, -np.pi, np.pi)
    return np.ascontiguousarray(a.astype(dtype)), np.ascontiguousarray(b.astype(dtype))


# ---------------------------------------------------------------------------------- Nao with cup and ball
NAO_START = np.array([1.125998, -0.691876, 1.888312, 0.776246, 0.245398, 1.259372, 0.279146, -1.587732, -0.510780, -1.823800])
NAO_GOAL = np.array([0.258284303377494, -0.2699099199363406, -0.01113121187052224, 1.2053012757652763, 1.2716626717484503,
                     -0.9826967097045605, 0.07355836822937814, 0.25450053440459897, -0.9512909033938429, -0.5297424293532234])
NAO_LO = np.deg2rad([-119.5, -94.5, -119.5, 0.5, -104.5, -119.5, 0.5, -119.5, -89.5, -104.5])
NAO_HI = np.deg2rad([119.5, -0.5, 119.5, 89.5, 104.5, 119.5, 94.5, 119.5, -0.5, 104.5])


def nao_states(n: int, seed: int, sigma: float = 0.3, uniform_fraction: float = 0.25, dtype=np.float64) -> np.ndarray:
    """Joint configurations of the Nao-cup scenario (demo/nao_cup/src/naocup.hpp:254-301): a quarter uniform in the joint
    limits (about 2 % of those are clear: the cup must stay upright), the rest scattered round the reference's start and goal
    configurations (the region a planner works in)."""
    rng = np.random.default_rng(seed)
    nu = int(n * uniform_fraction)
    q = np.empty((n, 10))
    q[:nu] = NAO_LO + (NAO_HI - NAO_LO) * rng.random((nu, 10))
    pick = rng.random((n - nu, 1)) < 0.5
    q[nu:] = np.where(pick, NAO_START, NAO_GOAL) + rng.normal(0.0, sigma, (n - nu, 10))
    return np.ascontiguousarray(np.clip(q, NAO_LO, NAO_HI).astype(dtype))


def nao_edges(n: int, seed: int, sigma: float = 0.3, reach: float = 0.25, dtype=np.float64):
    """Edges from nao_states to a configuration `reach`-scattered round them (|b - a| of a few tenths of a radian: some
    tens of 1-degree bisection midpoints each)."""
    a = nao_states(n, seed, sigma, 0.0, np.float64)
    rng = np.random.default_rng(seed + 1)
    b = np.clip(a + rng.normal(0.0, reach / np.sqrt(10.0), (n, 10)) * rng.random((n, 1)) * 2.0, NAO_LO, NAO_HI)
    return np.ascontiguousarray(a.astype(dtype)), np.ascontiguousarray(b.astype(dtype))


def link_arm_passage_scene(n_links: int):
    """A link-arm problem that needs a roadmap (in link_arm_scene start and goal connect almost directly): rings of
    circles of radius 3 round the base with gaps of about one circle radius, start = the arm stretched through the gap at angle 0,
    goal = stretched through the gap at 120 degrees.  To get from one to the other the arm has to fold back inside the inner ring.
    8 links: the reference's own PPRM (16 host threads) needs several thousand nodes; 16 links: it does not finish in 20 s.
    -> (lengths, link radius, circles [n, 3], start, goal)"""
    rings = {8: ((13.0, 6),), 16: ((13.0, 6), (30.0, 12)), 32: ((13.0, 6), (30.0, 12), (60.0, 24), (100.0, 40))}[n_links]
    circles = []
    for radius, count in rings:
        for i in range(count):
            a = 2 * np.pi * (i + 0.5) / count
            circles.append((radius * np.cos(a), radius * np.sin(a), 3.0))
    start, goal = np.zeros(n_links), np.zeros(n_links)
    goal[0] = 2 * np.pi / 3
    return np.full(n_links, 4.0), 0.5, np.asarray(circles), start, goal
