"""Adversarial contact configurations for the mesh path (test inputs only; numpy).

Every case is built so that float32 arithmetic is EXACT (identity rotation, small dyadic coordinates): whether the
triangles touch is then known by construction, and a separation of k units in the last place is a real one.
The reference decides these with fcl::collide (demo/se3_rigid_body_scenario.hpp:282-296); FCL is not available here, so
they pin the kernels to the oracle's definition and the two oracle formulations to each other."""
from __future__ import annotations

import numpy as np

ULP1 = float(np.spacing(np.float32(1.0)))  # spacing of float32 in [1, 2)

# environment: the square [0,4]^2 in the plane z = 1 (two triangles) and a small far triangle that fixes the bounding
# box -- and with it the contact band, 1e-6 of its diagonal -- at a known size
ENV = np.array([
    [[0, 0, 1], [4, 0, 1], [4, 4, 1]],
    [[0, 0, 1], [4, 4, 1], [0, 4, 1]],
    [[30, 30, 30], [31, 30, 30], [30, 31, 30]],
], dtype=np.float32)
ENV_DIAG = float(np.linalg.norm(ENV.reshape(-1, 3).max(0) - ENV.reshape(-1, 3).min(0)))
BAND = 1e-6 * ENV_DIAG

IDENT = (0.0, 0.0, 0.0, 1.0)

# robots (local frame) and the translation that puts them in exact contact with the square
CASES = {
    # lowest vertex (local origin) touches the interior of a face
    "vertex_on_face": dict(robot=[[0, 0, 0], [1, 0, 1], [0, 1, 1]], t=(1.0, 1.5, 1.0), axis=2, sign=+1),
    # the edge (0,-1,-.5)-(0,1,.5) crosses the square's boundary edge y = 0 in the single point (2,0,1)
    "edge_on_edge": dict(robot=[[0, -1, -0.5], [0, 1, 0.5], [1, -1, 2]], t=(2.0, 0.0, 1.0), axis=1, sign=-1),
    # an edge lying in the face
    "edge_in_face": dict(robot=[[0, 0, 0], [1, 0, 0], [0, 0, 1]], t=(1.0, 1.0, 1.0), axis=2, sign=+1),
    # coplanar, overlapping
    "coplanar_overlap": dict(robot=[[0, 0, 0], [1, 0, 0], [0, 1, 0]], t=(1.0, 1.0, 1.0), axis=2, sign=+1),
    # coplanar, sharing one boundary point only: vertex (0,0) of the robot on the square's corner (4,4)
    "coplanar_corner": dict(robot=[[0, 0, 0], [1, 0, 0], [0, 1, 0]], t=(4.0, 4.0, 1.0), axis=0, sign=+1),
    # vertex on vertex
    "vertex_on_vertex": dict(robot=[[0, 0, 0], [1, 1, 1], [1, 0, 1]], t=(4.0, 4.0, 1.0), axis=2, sign=+1),
}
KS = (-4096, -1024, -16, -1, 0, 1, 16, 1024, 4096)


def states_for(case: dict, dtype=np.float32):
    """States moving the robot by k ulps (of the coordinate's magnitude) along `axis`, `sign` pointing away from the
    obstacle: k <= 0 touches or penetrates, k > 0 is separated by exactly k ulps."""
    out = []
    for k in KS:
        t = np.array(case["t"], dtype=np.float64)
        step = float(np.spacing(np.float32(max(abs(t[case["axis"]]), 1.0))))
        t[case["axis"]] += case["sign"] * k * step
        out.append([*IDENT, *t])
    return np.asarray(out, dtype=dtype)


def expected_contact(name: str) -> np.ndarray:
    """Ground truth by construction (closed sets: touching is contact): 1 = collision."""
    # coplanar_overlap: moving along the normal in either direction separates the parallel planes;
    # vertex_on_vertex: the robot extends away from the square, below the plane it passes beside the corner
    if name in ("coplanar_overlap", "vertex_on_vertex"):
        return np.array([1 if k == 0 else 0 for k in KS], dtype=np.uint8)
    return np.array([1 if k <= 0 else 0 for k in KS], dtype=np.uint8)


def in_band() -> np.ndarray:
    """Which of the KS displacements lie inside the contact band (|k| ulp < BAND)."""
    return np.array([abs(k) * ULP1 * 4 < BAND for k in KS])  # ulp at |t| <= 4 is at most 4 ULP1


def world_triangles(case: dict, state) -> np.ndarray:
    return np.asarray(case["robot"], dtype=np.float64) + np.asarray(state[4:7], dtype=np.float64)
