"""Parity against the REFERENCE'S OWN code.  tests/golden/reference_golden.npz holds outputs of the reference's
headers (interpolate, DiscreteMotionValidator, PNG / holonomic / link-arm scenario checks, GoalState) compiled
from /root/reference against stand-in Eigen/Nigh headers (oracle/ref_driver.cpp, oracle/shim/).  The CPU tests
hold the oracle to those vectors (and, where /root/reference exists, re-run the reference live); the GPU tests
hold the CUDA kernels to them.  Scalar libm calls differ (the reference uses libm sin/cos/acos, we use
mptg_fpmath.h), so interpolated rotations are compared to a stated tolerance; decisions are compared exactly."""
from pathlib import Path

import numpy as np
import pytest

import mpt_b200 as m
from mpt_b200 import workloads as W
from tests import kats, reference_binding

ROOT = Path(__file__).resolve().parent.parent
G = np.load(ROOT / "tests" / "golden" / "reference_golden.npz")
HOLO_CIRCLES = [[170, 140, 80], [800, 70, 50], [900, 380, 70]]
HOLO_RECTS = [[375, 140, 520, 220], [200, 320, 390, 390], [600, 200, 680, 450]]
ARM5 = ([10.0, 12.0, 8.0, 6.0, 4.0], 0.5, [[20, -20, 8], [-20, -30, 5], [0, 25, 10], [30, 10, 10], [-30, 10, 8]])
ROT_TOL_F32, ROT_TOL_F64 = 3e-7, 1e-15  # interpolated quaternion coefficients: libm vs mptg_fpmath.h


def arm_scene(n):
    return ARM5 if n == 5 else W.link_arm_scene(n)


def check_interpolation(interp):
    """interp(space, a, b, t) -> states"""
    assert np.array_equal(interp(m.lp_space(3, 2, m.F64), G["l2_a"], G["l2_b"], G["l2_t"]), G["l2_out"])
    assert np.array_equal(interp(m.so2_space(1, 1, m.F64), G["so2_a"], G["so2_b"], G["l2_t"]).ravel(), G["so2_out"])
    o32 = interp(m.se3_space(50, 1, m.F32), G["se3_a"], G["se3_b"], G["se3_t"])
    assert np.array_equal(o32[:, 4:], G["se3_out_f32"][:, 4:])
    assert np.abs(o32[:, :4] - G["se3_out_f32"][:, :4]).max() <= ROT_TOL_F32
    o64 = interp(m.se3_space(50, 1, m.F64), G["se3_a"].astype(np.float64), G["se3_b"].astype(np.float64), G["l2_t"])
    assert np.array_equal(o64[:, 4:], G["se3_out_f64"][:, 4:])
    assert np.abs(o64[:, :4] - G["se3_out_f64"][:, :4]).max() <= ROT_TOL_F64


def check_scenarios(make_grid, make_shapes, make_arm):
    g = make_grid(G["grid_occ"])
    assert np.array_equal(g.valid(G["grid_a"]), G["grid_valid"])
    assert np.array_equal(g.link(G["grid_a"], G["grid_b"]), G["grid_link"])
    c = np.asarray(HOLO_CIRCLES, dtype=np.float64)
    h = make_shapes(2, c[:, :2], c[:, 2], HOLO_RECTS)
    assert np.array_equal(h.valid(G["holo_a"]), G["holo_valid"])
    assert np.array_equal(h.link(G["holo_a"], G["holo_b"]), G["holo_link"])
    for n in (5, 8, 16, 32):
        arm = make_arm(*arm_scene(n))
        assert np.array_equal(arm.valid(G[f"arm{n}_a"]), G[f"arm{n}_valid"]), n
        assert np.array_equal(arm.link(G[f"arm{n}_a"], G[f"arm{n}_b"]), G[f"arm{n}_link"]), n


# ------------------------------------------------------------------ CPU: oracle vs the reference's vectors
def test_oracle_interpolation_matches_reference(oracle):
    check_interpolation(oracle.interpolate)


def test_oracle_scenarios_match_reference(oracle):
    check_scenarios(oracle.grid, oracle.shapes, oracle.link_arm)


def test_oracle_dmv_matches_reference(oracle):
    """The oracle's DiscreteMotionValidator restatement vs the reference's loop, same state validator:
    identical decisions and identical number of states checked on every edge."""
    sp = m.se3_space(50, 1)
    mesh = oracle.mesh_pair(G["dmv_robot"], G["dmv_env"], sp, float(G["dmv_step"]))
    ok = mesh.link(G["dmv_a"], G["dmv_b"])
    assert np.array_equal(ok, G["dmv_ok"])
    assert mesh.last_states == int(G["dmv_states"].sum())


def test_goal_state_semantics():
    """src/mpt/goal_state.hpp:64-69: (true, 0) within the radius, else (false, d - radius)."""
    d = np.sqrt(((G["goal_q"] - 0.25) ** 2).sum(axis=1))
    assert np.array_equal(G["goal_is"], (d <= 1e-6).astype(np.uint8)) and G["goal_is"][0] == 1
    assert np.allclose(G["goal_dist"], np.where(d <= 1e-6, 0.0, d - 1e-6), rtol=0, atol=1e-15)


def test_planner_loop_restatement_builds_the_reference_planners_tree(oracle):
    """Row a11.  The reference's own Planner<Scenario, PRRT<single_threaded>> -- its Worker::solve / addSample loop,
    UniformBoxSampler, goal-biased sampling, GoalState, interpolate, PNG2dScenario::valid / link, compiled from
    /root/reference (oracle/ref_planner.cpp), fed with the product's sample stream -- and the oracle's restatement of
    that loop (one sample per wave) build the same tree: states bit-identical, same parents, same first goal node."""
    occ, lo, hi, start, goal = kats.prrt_scene()
    sp, og = m.lp_space(2, 2, m.F64), oracle.grid(occ)
    for i, (rng, n, goal_radius, goal_bias, seed) in enumerate(kats.PRRT_CASES):
        states, parents, goal_node = kats.replay_prrt(oracle, og, sp, lo, hi, start, goal, goal_radius, goal_bias, rng, seed, n, 1)
        assert states.shape[0] > 200
        assert np.array_equal(states, G[f"prrt{i}_states"]) and np.array_equal(parents, G[f"prrt{i}_parents"]), i
        assert goal_node == int(G[f"prrt{i}_goal"]), i
    assert int(G["prrt0_goal"]) != kats.NO_INDEX and int(G["prrt2_goal"]) == kats.NO_INDEX  # both outcomes are covered


@pytest.mark.skipif(not reference_binding.REFERENCE.exists(), reason="/root/reference not present (GPU box)")
def test_reference_planner_live_reproduces_committed_tree(oracle):
    ref = reference_binding.load()
    occ, lo, hi, start, goal = kats.prrt_scene()
    rng, n, goal_radius, goal_bias, seed = kats.PRRT_CASES[3]
    u = oracle.sample_uniforms(m.F64, seed, 0, n, 3)
    states, parents, goal_node = ref.prrt_grid(occ, lo, hi, start, goal, goal_radius, goal_bias, rng, u)
    assert np.array_equal(states, G["prrt3_states"]) and np.array_equal(parents, G["prrt3_parents"]) and goal_node == int(G["prrt3_goal"])


@pytest.mark.skipif(not reference_binding.REFERENCE.exists(), reason="/root/reference not present (GPU box)")
def test_reference_live_reproduces_committed_vectors():
    """Where the reference tree exists, rebuild oracle/_ref from its sources and re-run it: the committed
    vectors are what the reference computes."""
    ref = reference_binding.load()
    assert np.array_equal(ref.interpolate("l2_3", G["l2_a"], G["l2_b"], G["l2_t"]), G["l2_out"])
    assert np.array_equal(ref.interpolate("se3_f32", G["se3_a"], G["se3_b"], G["se3_t"]), G["se3_out_f32"])
    va, ln = ref.grid(G["grid_occ"], G["grid_a"], G["grid_b"])
    assert np.array_equal(va, G["grid_valid"]) and np.array_equal(ln, G["grid_link"])
    va, ln = ref.linkarm(*arm_scene(8), G["arm8_a"], G["arm8_b"])
    assert np.array_equal(va, G["arm8_valid"]) and np.array_equal(ln, G["arm8_link"])


# ------------------------------------------------------------------ GPU: kernels vs the reference's vectors
@pytest.mark.gpu
def test_device_interpolation_matches_reference(ctx):
    check_interpolation(ctx.interpolate)


@pytest.mark.gpu
def test_device_scenarios_match_reference(ctx):
    check_scenarios(lambda occ: m.Scenario.grid(ctx, occ, m.F64),
                    lambda dim, c, r, rects: m.Scenario.shapes(ctx, dim, c, r, rects, m.F64),
                    lambda lengths, radius, circles: m.Scenario.link_arm(ctx, lengths, radius, circles, m.F64))


@pytest.mark.gpu
def test_device_dmv_matches_reference(ctx):
    sp = m.se3_space(50, 1)
    sc = m.Scenario.mesh_pair(ctx, G["dmv_robot"], G["dmv_env"], sp, float(G["dmv_step"]))
    ok = sc.link(G["dmv_a"], G["dmv_b"])
    assert np.array_equal(ok, G["dmv_ok"])
    # valid edges touch exactly the states the reference's validator touched; an invalid edge touches between one
    # state and its whole set (coarse-to-fine order over all edges instead of the reference's per-edge queue)
    states = sc.last_stats()["states"]
    okm = G["dmv_ok"] == 1
    dist = ctx.distance(sp, G["dmv_a"], G["dmv_b"]).astype(np.float32)
    full = np.maximum(np.ceil(dist * (np.float32(1.0) / np.float32(G["dmv_step"]))).astype(np.int64), 1)
    assert np.array_equal(full[okm], G["dmv_states"][okm])
    assert int(full[okm].sum()) + int((~okm).sum()) <= states <= int(full.sum())


@pytest.mark.gpu
def test_device_prrt_builds_the_reference_planners_tree(ctx):
    """Row a11 on the device: the device-resident PRRT, one sample per wave, builds the tree the reference's own PRRT
    class built from the same sample stream (committed vectors) -- node for node, bit for bit."""
    occ, lo, hi, start, goal = kats.prrt_scene()
    sp, sc = m.lp_space(2, 2, m.F64), m.Scenario.grid(ctx, occ, m.F64)
    for i in (1, 3):
        rng, n, goal_radius, goal_bias, seed = kats.PRRT_CASES[i]
        pl = m.DevicePRRT(sc, sp, lo, hi, range=rng, goal=goal, goal_radius=goal_radius, goal_bias=goal_bias, seed=seed, capacity=8192, max_wave=64)
        pl.add_start(start)
        for _ in range(n):
            pl.wave(1)
        states, parents = pl.tree()
        assert np.array_equal(states, G[f"prrt{i}_states"]) and np.array_equal(parents, G[f"prrt{i}_parents"]), i
        assert pl.goal_node == int(G[f"prrt{i}_goal"]), i
        pl.close()
