"""The Nao-cup scenario (reference demo/nao_cup/src/{naocup,collide,linear}.hpp,
demo/nao_cup_planning.cpp:146-152).

tests/golden/nao_golden.npz holds the decisions of the REFERENCE'S OWN code (compiled from /root/reference against
oracle/shim/Eigen by oracle/ref_nao.cpp; script tests/golden/make_nao_golden.py): nao_clear on 16,384 configurations and
nao_link on 3,328 edges, float and double.  CPU tests hold the oracle (oracle/oracle_nao.hpp, general isometry products)
to them and, where /root/reference exists, re-run the reference live on fresh inputs; GPU tests hold the CUDA validator
(mpt_b200/csrc/nao.cuh, closed forms) to the oracle bit for bit on much larger sets and to the reference's vectors.
The reference calls libm sin/cos, the oracle and the kernels mptg_fpmath.h: a decision could differ only where a distance
sits within an ulp or two of its threshold; none does on these inputs (the margins are printed)."""
from pathlib import Path

import numpy as np
import pytest

import mpt_b200 as m
from mpt_b200 import workloads as W
from tests import reference_binding

ROOT = Path(__file__).resolve().parent.parent
G = np.load(ROOT / "tests" / "golden" / "nao_golden.npz")
TAGS = {m.F64: ("f64", np.float64), m.F32: ("f32", np.float32)}


def check_against_reference_vectors(make, scalar):
    tag, _ = TAGS[scalar]
    sc = make(scalar)
    assert np.array_equal(sc.valid(G[f"q_{tag}"]), G[f"clear_{tag}"])
    assert np.array_equal(sc.link(G[f"a_{tag}"], G[f"b_{tag}"]), G[f"link_{tag}"])
    return sc


# ------------------------------------------------------------------ CPU
@pytest.mark.parametrize("scalar", [m.F64, m.F32])
def test_oracle_matches_reference_vectors(oracle, scalar):
    tag, _ = TAGS[scalar]
    sc = check_against_reference_vectors(oracle.nao_cup, scalar)
    ok, margin = sc.valid(G[f"q_{tag}"], with_margin=True)
    assert G[f"clear_{tag}"][:2].tolist() == [1, 1]  # the reference's start and goal configurations are clear
    assert 0.05 < ok.mean() < 0.2
    print(f"{tag}: {int(ok.sum())} of {ok.size} clear; smallest |distance - threshold| {np.sort(margin)[:3]}")


def test_configs_match_reference():
    for scalar, (tag, _) in TAGS.items():
        start, goal, lo, hi = m.Scenario.nao_cup_configs(scalar)
        want = G[f"configs_{tag}"]
        assert np.array_equal(np.stack([start, goal, lo, hi]), want), tag
    assert np.allclose(W.NAO_START, G["configs_f64"][0]) and np.allclose(W.NAO_GOAL, G["configs_f64"][1])
    assert np.allclose(W.NAO_LO, G["configs_f64"][2], rtol=0, atol=1e-15) and np.allclose(W.NAO_HI, G["configs_f64"][3], rtol=0, atol=1e-15)


def test_oracle_link_semantics(oracle):
    """nao_link (naocup.hpp:809-840): an edge shorter than one degree is accepted without a look at anything -- even
    between two configurations in collision -- and the ends of a longer edge are not examined, only its midpoints."""
    sc = oracle.nao_cup(m.F64)
    q = G["q_f64"]
    bad = q[G["clear_f64"] == 0][:64]
    assert np.array_equal(sc.link(bad, bad + 1e-3), np.ones(64, np.uint8))
    deg = np.pi / 180.0
    a = np.tile(W.NAO_START, (4, 1))
    b = a.copy()
    b[:, 0] += np.array([0.99, 1.01, 1.99, 2.01]) * deg
    sc.link(a, b)
    assert sc.last_states == 0 + 1 + 1 + 3  # midpoints checked: none, 1, 1 (halves 0.995 deg), 3
    # an invalid END is not looked at (the planners validate states before linking them, :833-837) ...
    goal_side = W.NAO_START.copy()
    step = bad[:1] - goal_side
    short = goal_side + step * (1.9 * deg / np.linalg.norm(step))  # one midpoint, 0.95 deg from a clear configuration
    if sc.valid(0.5 * (goal_side + short[0]))[0]:
        assert sc.link(goal_side[None], short)[0] == 1
    # ... and a long edge into that configuration fails on some midpoint
    assert sc.link(goal_side[None], bad[:1])[0] == 0


@pytest.mark.skipif(not reference_binding.REFERENCE.exists(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("scalar", [m.F64, m.F32])
def test_reference_live(oracle, scalar):
    """Where the reference tree exists: rebuild oracle/_ref from its sources, reproduce the committed vectors, and compare
    the oracle with it on fresh inputs (65,536 configurations, 4,096 edges from clear configurations)."""
    tag, dt = TAGS[scalar]
    ref = reference_binding.load()
    ok, col = ref.nao_clear(G[f"q_{tag}"], scalar)
    assert np.array_equal(ok, G[f"clear_{tag}"]) and np.array_equal(col, G[f"collision_{tag}"])
    assert np.array_equal(ref.nao_link(G[f"a_{tag}"], G[f"b_{tag}"], scalar), G[f"link_{tag}"])
    sc = oracle.nao_cup(scalar)
    q = W.nao_states(65536, 909, dtype=dt)
    want, _ = ref.nao_clear(q, scalar)
    got, margin = sc.valid(q, with_margin=True)
    assert np.array_equal(got, want), f"{(got != want).sum()} decisions differ, margins {margin[got != want]}"
    a = q[want == 1][:4096]
    b = np.clip(a + np.random.default_rng(5).normal(0, 0.1, a.shape), W.NAO_LO, W.NAO_HI).astype(dt)
    assert np.array_equal(sc.link(a, b), ref.nao_link(a, b, scalar))


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("scalar", [m.F64, m.F32])
def test_device_matches_reference_vectors(ctx, scalar):
    check_against_reference_vectors(lambda s: m.Scenario.nao_cup(ctx, s), scalar)


@pytest.mark.gpu
@pytest.mark.parametrize("scalar", [m.F64, m.F32])
def test_device_matches_oracle(ctx, oracle, scalar):
    """Closed-form kinematics and root-free distance tests of the kernel against the general products of the oracle:
    identical decisions on 262,144 configurations (uniform and planner-like) and on edges of every length."""
    _, dt = TAGS[scalar]
    sc, og = m.Scenario.nao_cup(ctx, scalar), oracle.nao_cup(scalar)
    q = W.nao_states(262144, 11, dtype=dt)
    got, (want, margin) = sc.valid(q), og.valid(q, with_margin=True)
    assert np.array_equal(got, want), f"{(got != want).sum()} of {q.shape[0]} differ; margins {margin[got != want][:8]}"
    assert 0.05 < got.mean() < 0.2
    clear = q[got == 1]
    rng = np.random.default_rng(12)
    for reach, n in ((0.05, 8192), (0.3, 8192), (1.5, 2048)):
        a = clear[rng.integers(0, clear.shape[0], n)]
        b = np.clip(a + rng.normal(0, reach / np.sqrt(10), a.shape), W.NAO_LO, W.NAO_HI).astype(dt)
        got, want = sc.link(a, b), og.link(a, b)
        assert np.array_equal(got, want), f"reach {reach}: {(got != want).sum()} of {n} edges differ"
        probes = sc.last_stats()["prim_tests"]
        print(f"reach {reach}: {int(got.sum())} of {n} edges valid, {og.last_states} midpoints (oracle, early exit), {probes} (device)")
    # long edges across the joint range (up to 2^10 midpoints), clear or not at the ends
    a, b = W.nao_states(1024, 13, dtype=dt), W.nao_states(1024, 14, dtype=dt)
    assert np.array_equal(sc.link(a, b), og.link(a, b))
    # degenerate: a == b, and edges just below / above one degree
    assert np.array_equal(sc.link(q[:256], q[:256]), np.ones(256, np.uint8))
    d = np.zeros((512, 10))
    d[:, 3] = np.linspace(0.9, 1.1, 512) * np.pi / 180.0
    a = np.tile(clear[:1], (512, 1))
    b = (a.astype(np.float64) + d).astype(dt)
    assert np.array_equal(sc.link(a, b), og.link(a, b))


@pytest.mark.gpu
def test_device_edges_are_order_free(ctx, oracle):
    """An edge's decision is the AND over the midpoints of its recursion: the same from either end's point of view only
    if the midpoints coincide, which (a+b)/2 guarantees -- reversed edges give the same answers on the device and the oracle."""
    sc, og = m.Scenario.nao_cup(ctx, m.F64), oracle.nao_cup(m.F64)
    a, b = W.nao_edges(4096, 21)
    fwd, rev = sc.link(a, b), sc.link(b, a)
    assert np.array_equal(fwd, og.link(a, b)) and np.array_equal(rev, og.link(b, a))


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["nao", "arm"])
def test_flat_edge_check_limits(ctx, oracle, which):
    """flatLinkKernel (geom.cu): NaN ends behave as in the reference's recursion (its stop test never holds, the first
    midpoint fails: the edge is invalid; an arm edge fails on its invalid end), zero-length and sub-threshold edges need no
    midpoint, and an edge whose recursion would be deeper than 24 levels is an ERROR (MPTG_ERR_CAPACITY), not a silent cut."""
    if which == "nao":
        sc, og, D = m.Scenario.nao_cup(ctx, m.F64), oracle.nao_cup(m.F64), 10
        base = np.tile(W.NAO_START, (8, 1))
        huge = 1e8
    else:
        lengths, radius, circles = W.link_arm_scene(8)
        sc, og, D = m.Scenario.link_arm(ctx, lengths, radius, circles, m.F64), oracle.link_arm(lengths, radius, circles), 8
        st = W.box_states(4096, 8, 3, -np.pi, np.pi)
        base = np.tile(st[sc.valid(st) == 1][:1], (8, 1))
        huge = 1e7
    a, b = base.copy(), base.copy()
    b[1, 0] += 0.3
    a[2, 1] = np.nan
    b[3, D - 1] = np.nan
    a[4], b[4] = np.nan, np.nan
    b[5, 2] += 1e-9
    b[6, 0] += np.inf
    got, want = sc.link(a, b), og.link(a, b)
    assert np.array_equal(got, want) and got[0] == 1 and got[2] == 0 and got[3] == 0 and got[4] == 0 and got[5] == 1 and got[6] == 0
    b[7, 0] += huge  # 2^24 midpoints and more
    with pytest.raises(m.MptgError) as e:
        sc.link(a, b)
    assert e.value.code == -5  # MPTG_ERR_CAPACITY
    assert np.array_equal(sc.link(a[:6], b[:6]), want[:6])  # the geometry stays usable after the error


@pytest.mark.gpu
@pytest.mark.parametrize("scalar", [m.F64, m.F32])
def test_device_matches_oracle_at_bench_size(ctx, oracle, scalar):
    """The wave bench.py times (secondary.nao_cup_edges_*): 65,536 edges between clear configurations 0.3 rad apart, two
    lists (coarse to fine) on the device -- every decision equal to the oracle's, and the same edges answered in small
    batches (one list) give the same bytes."""
    _, dt = TAGS[scalar]
    sc, og = m.Scenario.nao_cup(ctx, scalar), oracle.nao_cup(scalar)
    pool = W.nao_states(1 << 20, 3, dtype=dt)
    pool = pool[sc.valid(pool) == 1]
    rng = np.random.default_rng(4)
    a = np.ascontiguousarray(pool[rng.integers(0, pool.shape[0], 65536)])
    b = np.ascontiguousarray(np.clip(a + rng.normal(0, 0.3 / np.sqrt(10), a.shape), W.NAO_LO, W.NAO_HI).astype(dt))
    got = sc.link(a, b)
    assert np.array_equal(got, og.link(a, b))
    small = np.concatenate([sc.link(a[i:i + 2048], b[i:i + 2048]) for i in range(0, 16384, 2048)])
    assert np.array_equal(small, got[:16384])
    assert 0.5 < got.mean() < 0.75
