"""CPU tests: the oracle against the reference's known-answer tests and against itself, golden
fixtures, host-side logic, and the C-ABI library's exports.  No GPU needed."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

import mpt_b200 as m
from mpt_b200 import _lib as L
from mpt_b200 import workloads as W
from tests import kats, oracle_binding

ROOT = Path(__file__).resolve().parent.parent


def test_reference_kats_cpp_binary():
    """oracle/kat_main.cpp: every reference metric/interpolation KAT, compiled C++ against the oracle."""
    oracle_binding.build()
    r = subprocess.run([str(ROOT / "oracle" / "_build" / "kat")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert r.stdout.count("PASS") >= 34 and "FAIL" not in r.stdout


def test_reference_kats_via_binding(oracle):
    assert kats.run_distance_kats(oracle.distance) == []
    assert kats.run_interpolate_kats(oracle.interpolate) == []


def test_fpmath_accuracy():
    """acos/sin/cos of include/mptg/mptg_fpmath.h vs libm (sampled; stride 1 is exhaustive)."""
    oracle_binding.build()
    r = subprocess.run([str(ROOT / "oracle" / "_build" / "fpmath_check"), "31"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout


SPACES = {
    "se3_50_1": lambda: m.se3_space(50, 1),
    "se3_1_1": lambda: m.se3_space(1, 1),
    "se3_5_2": lambda: m.se3_space(5, 2),
    "l2_2": lambda: m.lp_space(2, 2),
    "l2_3": lambda: m.lp_space(3, 2),
    "l1_8": lambda: m.lp_space(8, 1),
    "linf_16": lambda: m.lp_space(16, 0),
    "so3": lambda: m.so3_space(),
    "se2_3_2": lambda: m.se2_space(3, 2),
    "so2_l1_5": lambda: m.so2_space(5, 1),
}


def random_states(space, n, seed):
    """Random states of any product space (rotations unit-norm, angles in [-pi,pi], coords in [-10,10])."""
    rng = np.random.default_rng(seed)
    cols = []
    for i in range(space.desc.n_parts):
        p = space.desc.part[i]
        if p.kind == L.PART_SO3:
            cols.append(W.so3_uniform(rng, n, np.float64))
        elif p.kind == L.PART_SO2:
            cols.append(rng.random((n, p.dim)) * 2 * np.pi - np.pi)
        else:
            cols.append(rng.random((n, p.dim)) * 20 - 10)
    return np.ascontiguousarray(np.concatenate(cols, axis=1).astype(space.dtype))


@pytest.mark.parametrize("name", sorted(SPACES))
def test_tree_matches_brute(oracle, name):
    sp = SPACES[name]()
    pts = random_states(sp, 3000, 1)
    pts[100] = pts[7]  # exact duplicates: ties must resolve by index
    pts[200] = pts[7]
    q = random_states(sp, 200, 2)
    q[0] = pts[7]
    for k, radius in ((1, -1.0), (16, -1.0), (40, -1.0), (16, 3.0)):
        i0, d0, c0 = oracle.knn(sp, pts, q, k, radius)
        i1, d1, c1 = oracle.tree(sp, pts).knn(q, k, radius)
        assert (i0 == i1).all() and (c0 == c1).all()
        assert np.array_equal(d0, d1)
    i0, d0, _ = oracle.knn(sp, pts, q[:1], 3)
    assert list(i0[0]) == [7, 100, 200] and (d0[0] == 0).all()


def test_knn_order_and_padding(oracle):
    sp = m.lp_space(2, 2)
    pts = np.array([[0, 0], [1, 0], [0, 1], [3, 0]], dtype=np.float32)
    idx, dist, cnt = oracle.knn(sp, pts, np.array([[0, 0]], dtype=np.float32), 6)
    assert list(idx[0]) == [0, 1, 2, 3, L.NO_INDEX, L.NO_INDEX]
    assert cnt[0] == 4 and np.isinf(dist[0, 4:]).all()
    idx, dist, cnt = oracle.knn(sp, pts, np.array([[0, 0]], dtype=np.float32), 6, radius=1.0)
    assert cnt[0] == 3 and list(idx[0][:3]) == [0, 1, 2]


# ------------------------------------------------------------------ samplers
def test_philox_known_answers(oracle):
    """Philox4x32-10 against the published Random123 known-answer vectors that fit the sample streams' counter
    layout (c3 = 0 only for the first), and against an independent restatement for the rest."""
    def philox(c, k):
        c, k = list(c), list(k)
        for _ in range(10):
            a, b = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
            c = [(b >> 32) ^ c[1] ^ k[0], b & 0xFFFFFFFF, (a >> 32) ^ c[3] ^ k[1], a & 0xFFFFFFFF]
            k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
        return c
    # Random123 kat_vectors: philox4x32 10
    assert philox([0] * 4, [0] * 2) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]
    import ctypes as C
    out = (C.c_uint32 * 4)()
    oracle.lib.orc_philox_block(C.c_uint64(0), C.c_uint64(0), C.c_uint32(0), out)
    assert list(out) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    rng = np.random.default_rng(5)
    for _ in range(50):
        seed, g, blk = int(rng.integers(0, 2**63)), int(rng.integers(0, 2**63)), int(rng.integers(0, 2**31))
        oracle.lib.orc_philox_block(C.c_uint64(seed), C.c_uint64(g), C.c_uint32(blk), out)
        assert list(out) == philox([g & 0xFFFFFFFF, g >> 32, blk, 0], [seed & 0xFFFFFFFF, seed >> 32])


def sampler_kats(sample_from_uniforms):
    """test/scenario_sampler_test.cpp:198-270 drives the uniform samplers with a generator that always returns its
    minimum (all uniforms 0): SO(3) -> coeffs (1,0,0,0), box -> its minimum corner, SE(3) -> both."""
    bad = []
    for scalar in (m.F32, m.F64):
        so3 = m.so3_space(scalar)
        if not np.array_equal(sample_from_uniforms(so3, 0, 0, np.zeros((1, 3)))[0], [1, 0, 0, 0]):
            bad.append(("so3", scalar))
        lp2 = m.lp_space(2, 2, scalar)
        if not np.array_equal(sample_from_uniforms(lp2, [1, 2], [3, 5], np.zeros((1, 2)))[0], [1, 2]):
            bad.append(("lp2", scalar))
        se3 = m.se3_space(50, 1, scalar)
        lo, hi = [0, 0, 0, 0, 1, 2, 3], [0, 0, 0, 0, 4, 5, 6]
        if not np.array_equal(sample_from_uniforms(se3, lo, hi, np.zeros((1, 6)))[0], [1, 0, 0, 0, 1, 2, 3]):
            bad.append(("se3", scalar))
        so2 = m.so2_space(3, 1, scalar)
        got = sample_from_uniforms(so2, 0, 0, np.zeros((1, 3)))[0]
        if not np.array_equal(got, np.full(3, -np.pi, dtype=got.dtype)):
            bad.append(("so2", scalar))
    return bad


def test_sampler_reference_kats(oracle):
    assert sampler_kats(lambda sp, lo, hi, u: oracle.sample_from_uniforms(sp, lo, hi, u, u.shape[1])) == []


def test_sampler_distributions(oracle):
    """Uniform on the box; unit quaternions uniform on S^3 (each coefficient mean 0, variance 1/4; the rotation
    angle theta = 2 acos|w| has density (1 - cos theta)/pi, i.e. P(theta < t) = (t - sin t)/pi)."""
    n = 200_000
    for scalar in (m.F32, m.F64):
        se3 = m.se3_space(50, 1, scalar)
        lo, hi = [0, 0, 0, 0, -10, 0, 5], [0, 0, 0, 0, 10, 1, 7]
        s = oracle.sample(se3, lo, hi, seed=12345, first=0, n=n).astype(np.float64)
        q, t = s[:, :4], s[:, 4:]
        assert np.abs(np.linalg.norm(q, axis=1) - 1).max() < (2e-6 if scalar == m.F32 else 1e-14)
        assert np.abs(q.mean(axis=0)).max() < 0.005 and np.abs(q.var(axis=0) - 0.25).max() < 0.005
        theta = 2 * np.arccos(np.minimum(1.0, np.abs(q[:, 3])))
        for tt in (0.5, 1.0, 2.0, 3.0):
            assert abs((theta < tt).mean() - (tt - np.sin(tt)) / np.pi) < 0.005
        assert (t >= np.array(lo[4:])).all() and (t < np.array(hi[4:])).all()
        assert np.abs(t.mean(axis=0) - np.array([0, 0.5, 6])).max() < 0.03
        assert np.abs(t.var(axis=0) - np.array([400, 1, 4]) / 12).max() < 0.5
        # streams: disjoint sample numbers are the same samples whatever the batching; other seeds differ
        again = np.concatenate([oracle.sample(se3, lo, hi, 12345, 0, 1000), oracle.sample(se3, lo, hi, 12345, 1000, 500)])
        assert np.array_equal(again, oracle.sample(se3, lo, hi, 12345, 0, 1500))
        assert not np.array_equal(oracle.sample(se3, lo, hi, 12346, 0, 100), again[:100])
    # goal bias: the goal state replaces about goal_bias of the samples
    lp = m.lp_space(2, 2, m.F64)
    g = oracle.sample(lp, [0, 0], [1, 1], 7, 0, 100_000, goal=[5.0, 5.0], goal_bias=0.05)
    assert abs((g[:, 0] == 5.0).mean() - 0.05) < 0.003


def test_grid_semantics(oracle):
    """demo/png_2d_scenario.hpp:104-117: round-half-up indexing, row wrap at x == width, out of range -> obstacle."""
    occ = np.zeros((4, 6), dtype=np.uint8)
    occ[2, 3] = 1
    g = oracle.grid(occ)
    st = np.array([[3.4, 2.4], [2.6, 1.6], [3.5, 2.0], [5.6, 0.0], [5.6, 3.0], [0.0, 3.6], [0, 0]], dtype=np.float64)
    #            inside cell   rounds to (3,2)  (4,2) free  x=6 wraps to (0,1) free, x=6,y=3 -> idx 24 out of range
    assert list(g.valid(st)) == [0, 0, 1, 1, 0, 0, 1]
    a = np.array([[0.0, 2.0], [0.0, 0.0]], dtype=np.float64)
    b = np.array([[5.0, 2.0], [5.0, 0.0]], dtype=np.float64)
    assert list(g.link(a, b)) == [0, 1]


def test_linkarm_and_shapes_semantics(oracle):
    lengths, radius, circles = [1.0, 1.0], 0.1, [[1.5, 0.0, 0.2]]
    arm = oracle.link_arm(lengths, radius, circles)
    st = np.array([[0.0, 0.0], [np.pi / 2, 0.0]], dtype=np.float64)
    assert list(arm.valid(st)) == [0, 1]  # straight arm passes through the circle
    assert list(arm.link(st[1:], np.array([[np.pi / 2, 0.3]]))) == [1]
    assert list(arm.link(st[1:], np.array([[-np.pi / 2, 0.0]]))) == [0]  # sweeps through the obstacle
    sh = oracle.shapes(2, [[5.0, 5.0]], [1.0], rects=[[8, 8, 9, 9]])
    assert list(sh.valid(np.array([[5.0, 5.5], [0, 0], [8.5, 8.5]], dtype=np.float64))) == [0, 1, 0]
    assert list(sh.link(np.array([[0.0, 5.0], [0.0, 0.0], [7.0, 8.5]]), np.array([[10.0, 5.0], [10.0, 0.0], [10.0, 8.5]]))) == [0, 1, 0]


def test_mesh_oracle_margin(oracle):
    """Two unit triangles: collide when overlapping, margin ~ the gap when separated."""
    sp = m.se3_space(50, 1)
    robot = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], dtype=np.float32)
    env = np.array([[[0.2, 0.2, -0.5], [0.2, 0.2, 0.5], [0.8, 0.8, 0.5]]], dtype=np.float32)
    g = oracle.mesh_pair(robot, env, sp, 0.5)
    st = np.array([[0, 0, 0, 1, 0, 0, 0], [0, 0, 0, 1, 0, 0, 2.0]], dtype=np.float32)
    ok, margin = g.valid(st, with_margin=True)
    assert list(ok) == [0, 1]
    assert margin[0] <= 0
    assert list(g.link(st[1:], st[:1])) == [0]
    assert list(g.link(st[1:], np.array([[0, 0, 0, 1, 0, 0, 3.0]], dtype=np.float32))) == [1]


def test_dmv_counts_states_like_reference(oracle):
    """DiscreteMotionValidator: valid(to) + (steps-1) interior states for a free edge (:75-82)."""
    sp = m.se3_space(50, 1)
    robot = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], dtype=np.float32)
    env = np.array([[[100, 100, 100], [101, 100, 100], [100, 101, 100]]], dtype=np.float32)
    g = oracle.mesh_pair(robot, env, sp, 0.25)
    a = np.array([[0, 0, 0, 1, 0, 0, 0]], dtype=np.float32)
    b = np.array([[0, 0, 0, 1, 2.0, 0, 0]], dtype=np.float32)  # distance 2 -> steps 8 -> 1 + 7 states
    assert list(g.link(a, b)) == [1] and g.last_states == 8


def test_golden_fixtures(oracle):
    """Committed fixtures (tests/golden/make_golden.py) pin the oracle's outputs across refactors."""
    g = np.load(ROOT / "tests" / "golden" / "golden.npz")
    sp = m.se3_space(50, 1)
    idx, dist, _ = oracle.knn(sp, g["se3_pts"], g["se3_q"], 16)
    assert np.array_equal(idx, g["se3_knn_idx"]) and np.array_equal(dist, g["se3_knn_dist"])
    l1 = m.lp_space(8, 1)
    idx, dist, _ = oracle.knn(l1, g["l1_pts"], g["l1_q"], 5)
    assert np.array_equal(idx, g["l1_knn_idx"]) and np.array_equal(dist, g["l1_knn_dist"])
    assert np.array_equal(oracle.interpolate(sp, g["se3_q"][:32], g["se3_pts"][:32], g["interp_t"]), g["se3_interp"])
    grid = oracle.grid(g["grid_occ"])
    assert np.array_equal(grid.link(g["grid_a"], g["grid_b"]), g["grid_link"])
    arm = oracle.link_arm(g["arm_lengths"], float(g["arm_radius"]), g["arm_circles"])
    assert np.array_equal(arm.link(g["arm_a"], g["arm_b"]), g["arm_link"])
    mesh = oracle.mesh_pair(g["mesh_robot"], g["mesh_env"], sp, float(g["mesh_step"]))
    assert np.array_equal(mesh.valid(g["mesh_states"]), g["mesh_valid"])
    assert np.array_equal(mesh.link(g["mesh_a"], g["mesh_b"]), g["mesh_link"])


def test_cabi_exports_every_declared_symbol():
    """include/mptg/mptg.h vs libmptg.so: every declared function is exported and bound (no compute)."""
    header = (ROOT / "include" / "mptg" / "mptg.h").read_text()
    declared = set(re.findall(r"\b(mptg_[a-z0-9_]+)\s*\(", header))
    lib = L.load()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    assert lib.mptg_abi_version() == 2
    sp = m.se3_space(50, 1)
    assert sp.scalars == 7 and sp.dimensions == 6


def test_no_cpu_fallback_without_gpu():
    """On a machine without a CUDA device the product fails loudly instead of computing on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(m.MptgError) as e:
        m.Context(0)
    assert e.value.code == L.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_oracle():
    """Nothing under mpt_b200/ or include/ may include, import, load or link the oracle."""
    pat = re.compile(r"^\s*(#\s*include|import|from)\b.*oracle|liboracle|oracle_binding|oracle/_build|CDLL\(.*oracle", re.M)
    for p in list((ROOT / "mpt_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if p.is_file() and p.suffix in {".py", ".cu", ".cuh", ".h", ".hpp", ".cpp"}:
            hit = pat.search(p.read_text())
            assert hit is None, (p, hit.group(0))


# ------------------------------------------------------------------ contact: two formulations of triangle-triangle
def test_tri_tri_formulations_agree_outside_the_contact_band(oracle):
    """FCL is absent, so the 17-axis SAT (oracle.hpp triTriIntersect, what the kernels restate) is pinned the only way
    left: an independent formulation from orientation predicates (triTriPredicates, double) must give the same answer
    on every pair that is not within the contact band, and on the constructed touching cases both must give the answer
    known by construction (closed sets: touching is contact)."""
    from tests import contact_cases as cc

    rng = np.random.default_rng(77)
    n = 300_000
    P = rng.normal(size=(n, 3, 3))
    Q = rng.normal(size=(n, 3, 3)) + rng.normal(size=(n, 1, 3)) * 0.8
    # near-degenerate families: slivers, shared vertices, pairs pushed to first contact along a random direction
    P[:20000, 2] = P[:20000, 0] + (P[:20000, 1] - P[:20000, 0]) * 0.5 + rng.normal(size=(20000, 3)) * 1e-6
    Q[20000:40000, 0] = P[20000:40000, 1]
    sat, margin, pred = oracle.tri_pairs(P, Q)
    band = 1e-9 * 4.0  # double arithmetic on unit-size triangles
    differ = sat != pred
    inside = np.abs(margin) < band
    print(f"tri-tri: {sat.mean():.3f} intersect, {int(inside.sum())} pairs within the band, {int(differ.sum())} differ "
          f"({int((differ & inside).sum())} of them inside)")
    assert not (differ & ~inside).any()
    assert inside.sum() > 0  # the shared-vertex family touches by construction
    sp = m.se3_space(50, 1)
    for name, case in cc.CASES.items():
        st = cc.states_for(case)
        og = oracle.mesh_pair(np.array([case["robot"]], np.float32), cc.ENV, sp, 0.5)
        want = cc.expected_contact(name)
        ok_sat = og.valid(st)
        og.use_predicates(True)
        ok_pred = og.valid(st)
        assert np.array_equal(1 - ok_sat, want), (name, ok_sat, want)
        assert np.array_equal(1 - ok_pred, want), (name, ok_pred, want)
