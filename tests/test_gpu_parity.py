"""GPU parity tests: the CUDA path, called through the C ABI (libmptg.so), against the CPU oracle on
the same seeded inputs, against the committed golden fixtures and against the reference's KATs.

Bars: kNN indices AND distances bit-identical; grid / shapes / link-arm decisions bit-identical;
mesh decisions identical except states the oracle flags as within 1e-6 (relative to the scene
diagonal) of contact, which are counted and reported, never silently dropped."""
from pathlib import Path

import numpy as np
import pytest

import mpt_b200 as m
from mpt_b200 import _lib as L
from mpt_b200 import workloads as W
from tests import kats
from tests.test_oracle import SPACES, random_states

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def assert_knn_equal(got, want):
    gi, gd, gc = got
    wi, wd, wc = want
    assert np.array_equal(gc, wc), "neighbour counts differ"
    bad = np.nonzero((gi != wi).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} queries differ, first {bad[0]}: {gi[bad[0]]} vs {wi[bad[0]]}; d {gd[bad[0]]} vs {wd[bad[0]]}"
    assert np.array_equal(gd, wd), "distances not bit-identical"


# ------------------------------------------------------------------ metric
def test_reference_kats_on_device(ctx):
    """The reference's metric / interpolation known-answer tests, evaluated by the CUDA kernels (f64)."""
    assert kats.run_distance_kats(ctx.distance) == []
    assert kats.run_interpolate_kats(ctx.interpolate) == []


@pytest.mark.parametrize("name", sorted(SPACES))
@pytest.mark.parametrize("scalar", [m.F32, m.F64])
def test_distance_interpolate_bit_exact(ctx, oracle, name, scalar):
    sp0 = SPACES[name]()
    parts = [(("lp", "so2", "so3")[sp0.desc.part[i].kind - 1], sp0.desc.part[i].p, sp0.desc.part[i].dim, sp0.desc.part[i].weight)
             for i in range(sp0.desc.n_parts)]
    sp = m.Space(parts, scalar)
    a, b = random_states(sp, 4096, 11), random_states(sp, 4096, 12)
    b[:16] = a[:16]          # zero distance
    if name.startswith("se3") or name == "so3":
        b[16:32, :4] = -a[16:32, :4]  # antipodal quaternions: same rotation (so3_space.hpp:61 quirk path)
    assert np.array_equal(ctx.distance(sp, a, b), oracle.distance(sp, a, b))
    t = np.random.default_rng(5).random(4096).astype(sp.dtype)
    t[:4] = [0, 1, 0.5, 0.25]
    gi, oi = ctx.interpolate(sp, a, b, t), oracle.interpolate(sp, a, b, t)
    assert np.array_equal(gi, oi, equal_nan=True)
    d = oracle.distance(sp, a, b)
    gs, gd = ctx.steer(sp, a, b, d, 2.5, with_distance=True)
    os_, od = oracle.steer(sp, a, b, d, 2.5, with_distance=True)
    assert np.array_equal(gs, os_, equal_nan=True) and np.array_equal(gd, od, equal_nan=True)


# ------------------------------------------------------------------ kNN
@pytest.mark.parametrize("name", sorted(SPACES))
def test_knn_brute_matches_oracle(ctx, oracle, name):
    sp = SPACES[name]()
    pts = random_states(sp, 5000, 1)
    pts[100] = pts[7]
    pts[200] = pts[7]
    q = random_states(sp, 300, 2)
    q[0] = pts[7]
    nn = m.Nearest(ctx, sp, 8192, m.KNN_BRUTE)
    assert nn.insert(pts[:1000]) == 0 and nn.insert(pts[1000:]) == 1000 and nn.size() == 5000
    assert np.array_equal(nn.states(990, 20), pts[990:1010])
    for k, radius in ((1, -1.0), (16, -1.0), (33, -1.0), (100, -1.0), (16, 3.0)):
        assert_knn_equal(nn.nearest(q, k, radius), oracle.knn(sp, pts, q, k, radius))
    idx, dist, cnt = nn.nearest(q[:1], 3)
    assert list(idx[0]) == [7, 100, 200] and (dist[0] == 0).all()
    nn.close()


@pytest.mark.parametrize("name", sorted(SPACES))
def test_knn_bvh_matches_oracle(ctx, oracle, name):
    """The 32-ary box tree returns exactly what the scan returns (indices, distances, ties)."""
    sp = SPACES[name]()
    pts = random_states(sp, 40_000, 1)
    pts[100] = pts[7]
    pts[200] = pts[7]
    q = random_states(sp, 400, 2)
    q[0] = pts[7]
    nn = m.Nearest(ctx, sp, 65536, m.KNN_BVH)
    nn.insert(pts)
    tree = oracle.tree(sp, pts)
    for k, radius in ((1, -1.0), (16, -1.0), (49, -1.0), (100, -1.0), (16, 2.0)):
        assert_knn_equal(nn.nearest(q, k, radius), tree.knn(q, k, radius))
    st = nn.last_stats()
    assert st["strategy"] == m.KNN_BVH and st["indexed"] == 40_000 and 0 < st["distance_evals"] < 400 * 40_000
    idx, dist, cnt = nn.nearest(q[:1], 3)
    assert list(idx[0]) == [7, 100, 200] and (dist[0] == 0).all()
    nn.close()


def test_knn_bvh_with_unindexed_tail(ctx, oracle):
    """Points inserted after the build are scanned and merged until the index is refreshed."""
    sp = m.se3_space(50, 1)
    pts = W.se3_states(60_000, 9)
    q = W.se3_states(500, 10)
    nn = m.Nearest(ctx, sp, 65536, m.KNN_BVH)
    nn.insert(pts[:36_000])
    tree = oracle.tree(sp, pts[:36_000])
    assert_knn_equal(nn.nearest(q, 16), tree.knn(q, 16))
    nn.insert(pts[36_000:41_000])       # tail of 5000 <= count/2: index kept, tail searched separately
    assert_knn_equal(nn.nearest(q, 16), oracle.tree(sp, pts[:41_000]).knn(q, 16))
    assert nn.last_stats()["indexed"] == 36_000
    nn.insert(pts[41_000:])             # tail of 24000 > count/2: rebuilt
    assert_knn_equal(nn.nearest(q, 16), oracle.tree(sp, pts).knn(q, 16))
    assert nn.last_stats()["indexed"] == 60_000
    nn.close()


@pytest.mark.parametrize("name", ["se3_f32", "l2_2d_f64", "l1_8d_f64", "so2_3_f32"])
def test_knn_tail_leaves_and_raw_remainder(ctx, oracle, name):
    """The planner's access pattern on a tree-indexed set: batches arrive between searches; batches of >= 1024 points
    become Morton-sorted leaves (several chunks), smaller ones are scanned; everything is merged with the tree's answer.
    Results stay bit-identical to the exhaustive oracle, for k = 1, a large k, a radius, duplicates of tree points in
    the tail, and a sharded index map."""
    sp = {"se3_f32": m.se3_space(50, 1), "l2_2d_f64": m.lp_space(2, 2, m.F64), "l1_8d_f64": m.lp_space(8, 1, m.F64),
          "so2_3_f32": m.so2_space(3, 1, m.F32)}[name]
    pts = random_states(sp, 52_000, 21)
    pts[41_000:41_050] = pts[100:150]      # duplicates of indexed points inside the tail: ties broken by index
    q = random_states(sp, 300, 22)
    q[:20] = pts[41_000:41_020]
    nn = m.Nearest(ctx, sp, 65536, m.KNN_BVH)
    nn.insert(pts[:40_000])
    nn.nearest(q[:8], 1)                   # builds the tree over 40,000 points
    n = 40_000
    for batch, k, radius in ((1500, 16, None), (300, 1, None), (2500, 48, None), (40, 16, None), (3000, 100, None), (1200, 16, 0.5)):
        nn.insert(pts[n:n + batch])
        n += batch
        if radius is not None:
            d16 = oracle.knn(sp, pts[:n], q, 16)[1]
            radius = float(np.median(d16[:, 8]))   # about half of the neighbours fall inside
        got = nn.nearest(q, k, radius) if radius is not None else nn.nearest(q, k)
        want = oracle.knn(sp, pts[:n], q, k, radius if radius is not None else -1.0)
        assert_knn_equal(got, want)
        assert nn.last_stats()["indexed"] == 40_000   # still the first tree: the tail did the rest
    nn.close()
    # sharded index map (global = local * 4 + 1), as on 4 GPUs
    nn = m.Nearest(ctx, sp, 65536, m.KNN_BVH)
    nn.set_index_map(4, 1)
    nn.insert(pts[:40_000])
    nn.nearest(q[:8], 1)
    nn.insert(pts[40_000:43_000])
    gi, gd, gc = nn.nearest(q, 16)
    wi, wd, wc = oracle.knn(sp, pts[:43_000], q, 16)
    assert np.array_equal(gi, wi * 4 + 1) and np.array_equal(gd, wd) and np.array_equal(gc, wc)
    nn.close()


def test_knn_auto_strategy_and_ragged_sizes(ctx, oracle):
    sp = m.se3_space(50, 1)
    for n in (1, 31, 32, 33, 1023, 1025, 16384, 33_000):
        pts = W.se3_states(n, 100 + n)
        q = W.se3_states(65, 7)
        for strat in (m.KNN_AUTO, m.KNN_BVH):
            nn = m.Nearest(ctx, sp, max(n, 64), strat)
            nn.insert(pts)
            assert_knn_equal(nn.nearest(q, 16), oracle.knn(sp, pts, q, 16))
            nn.close()


def test_knn_bvh_double(ctx, oracle):
    sp = m.se3_space(50, 1, m.F64)
    pts = W.se3_states(20_000, 1, dtype=np.float64)
    q = W.se3_states(100, 2, dtype=np.float64)
    nn = m.Nearest(ctx, sp, 32768, m.KNN_BVH)
    nn.insert(pts)
    assert_knn_equal(nn.nearest(q, 16), oracle.knn(sp, pts, q, 16))


@pytest.mark.parametrize("scalar", [m.F32, m.F64])
def test_knn_index_device_build_equals_host_build(ctx, oracle, scalar, monkeypatch):
    """The index built on the device (knn_build.cu; double-precision sets are partitioned by their float roundings
    and emitted exactly) and the host build of the same tree give the same, exact, answers."""
    dt = np.float32 if scalar == m.F32 else np.float64
    for sp, pts, q in (
        (m.se3_space(50, 1, scalar), W.se3_states(70_000, 11, dtype=dt), W.se3_states(300, 12, dtype=dt)),
        (m.lp_space(2, 2, scalar), W.box_states(50_000, 2, 13, 0.0, [3976, 2603], dt), W.box_states(300, 2, 14, 0.0, [3976, 2603], dt)),
        (m.lp_space(8, 1, scalar), W.box_states(40_000, 8, 15, -np.pi, np.pi, dt), W.box_states(200, 8, 16, -np.pi, np.pi, dt)),
    ):
        want = oracle.knn(sp, pts, q, 20)
        results = []
        for host in ("0", "1"):
            monkeypatch.setenv("MPTG_KNN_HOST_BUILD", host)
            nn = m.Nearest(ctx, sp, 1 << 17, m.KNN_BVH)
            nn.insert(pts)
            nn.build_index()
            results.append(nn.nearest(q, 20))
            nn.close()
        assert_knn_equal(results[0], want)
        assert_knn_equal(results[1], want)


@pytest.mark.parametrize("scalar", [m.F32, m.F64])
def test_knn_l1_scan_several_queries_per_warp(ctx, oracle, scalar):
    """Arm-space scan with several queries per warp (knnBruteL1Kernel; waves of >= 512 queries, k <= 64): bit-exact
    indices and distances against the oracle and against the one-query-per-warp kernel, over dimensions that are not
    multiples of anything, ragged wave sizes, k in both register layouts, radius searches, duplicates, a NaN query, and the
    split scan (few query blocks -> several partial lists per query merged)."""
    dt = np.float32 if scalar == m.F32 else np.float64
    for dim, n_pts, n_q in ((5, 3000, 515), (8, 9000, 1000), (16, 2500, 512), (32, 1100, 777), (7, 33, 600)):
        sp = m.lp_space(dim, 1, scalar)
        pts = W.box_states(n_pts, dim, 100 + dim, -np.pi, np.pi, dt)
        pts[n_pts // 2:n_pts // 2 + 40] = pts[3]  # ties: index order decides
        q = W.box_states(n_q, dim, 200 + dim, -np.pi, np.pi, dt)
        q[5] = pts[3]
        q[7, dim - 1] = np.nan
        nn = m.Nearest(ctx, sp, 16384, m.KNN_BRUTE)
        nn.insert(pts)
        for k, radius in ((1, -1.0), (16, -1.0), (37, -1.0), (64, -1.0), (20, 0.35 * dim), (100, -1.0)):
            got, want = nn.nearest(q, k, radius), oracle.knn(sp, pts, q, k, radius)
            assert got[2][7] == 0
            assert_knn_equal(got, want)
            small = nn.nearest(q[:100], k, radius)  # below the threshold: the one-query-per-warp kernel
            assert np.array_equal(small[0], got[0][:100]) and np.array_equal(small[1].view(np.uint8), got[1][:100].view(np.uint8))
        nn.close()


def test_knn_l1_double_sets_extreme_values(ctx, oracle):
    """Double-precision L1 sets on inputs a reduced-precision shortcut would get wrong (written for a float-prefilter
    variant of the scan that was measured and dropped; kept for the shipped kernel): coordinates of
    magnitude 1e4 / 1e7 whose DIFFERENCES are far below the float resolution, neighbours separated by 1e-12, a set that
    grows between searches, values beyond the float range, NaN and infinite coordinates in points and queries, ties."""
    rng = np.random.default_rng(8)
    for dim, scale, spread in ((8, 1e4, 1e-3), (16, 1e7, 1e-2), (8, 1.0, 1e-12), (12, 3.0, 1.0)):
        sp = m.lp_space(dim, 1, m.F64)
        centre = rng.uniform(-scale, scale, dim)
        pts = centre + rng.uniform(-spread, spread, (6000, dim))
        pts[100:140] = pts[7]
        q = centre + rng.uniform(-spread, spread, (700, dim))
        q[3] = pts[7]
        nn = m.Nearest(ctx, sp, 16384, m.KNN_BRUTE)
        nn.insert(pts[:2500])
        for k, radius in ((16, -1.0), (40, -1.0), (5, 0.4 * spread * dim)):
            assert_knn_equal(nn.nearest(q, k, radius), oracle.knn(sp, pts[:2500], q, k, radius))
        nn.insert(pts[2500:])
        for k, radius in ((1, -1.0), (16, -1.0), (64, -1.0)):
            assert_knn_equal(nn.nearest(q, k, radius), oracle.knn(sp, pts, q, k, radius))
        nn.close()
    # out-of-range and non-finite values
    sp = m.lp_space(8, 1, m.F64)
    pts = rng.uniform(-3, 3, (4000, 8))
    pts[10, 2] = 1e39       # beyond float: infinite in the mirror, finite in the set
    pts[11, 0] = np.inf
    pts[12, 5] = np.nan
    pts[13] = -1e39
    q = rng.uniform(-3, 3, (600, 8))
    q[0, 1] = np.nan
    q[1, 1] = 1e39
    q[2, 1] = np.inf
    q[4] = -1e39
    nn = m.Nearest(ctx, sp, 8192, m.KNN_BRUTE)
    nn.insert(pts)
    for k in (4, 16):
        got, want = nn.nearest(q, k), oracle.knn(sp, pts, q, k)
        assert got[2][0] == 0 and want[2][0] == 0
        assert_knn_equal(got, want)
    nn.close()


def test_knn_degenerate_point_sets(ctx, oracle):
    """All points identical / on a line / two clusters: zero-extent boxes, massive ties (index order decides)."""
    sp = m.se3_space(50, 1)
    base = W.se3_states(1, 3)
    same = np.repeat(base, 20_000, axis=0)
    line = np.repeat(base, 20_000, axis=0)
    line[:, 4] = np.linspace(-50, 50, 20_000, dtype=np.float32)
    two = np.concatenate([np.repeat(W.se3_states(1, 4), 10_000, axis=0), np.repeat(W.se3_states(1, 5), 10_000, axis=0)])
    q = np.concatenate([base, W.se3_states(63, 6)])
    for pts in (same, line, two):
        for strat in (m.KNN_BRUTE, m.KNN_BVH):
            nn = m.Nearest(ctx, sp, 32768, strat)
            nn.insert(pts)
            for k in (1, 16, 40):
                assert_knn_equal(nn.nearest(q, k), oracle.knn(sp, pts, q, k))
            nn.close()


def test_knn_nan_and_far_queries(ctx, oracle):
    sp = m.se3_space(50, 1)
    pts = W.se3_states(30_000, 1)
    q = W.se3_states(16, 2)
    q[0, 4] = np.nan          # NaN query: no distance compares true, no neighbours
    q[1, 4:] = 1e6            # far away: still exact
    q[2, :4] = 0              # zero quaternion: |dot| = 0, rotation term pi/2 * w everywhere
    for strat in (m.KNN_BRUTE, m.KNN_BVH):
        nn = m.Nearest(ctx, sp, 32768, strat)
        nn.insert(pts)
        got, want = nn.nearest(q, 8), oracle.knn(sp, pts, q, 8)
        assert got[2][0] == 0 and want[2][0] == 0
        assert_knn_equal(got, want)
        nn.close()


def test_mesh_empty_and_far(ctx, oracle):
    sp = m.se3_space(50, 1)
    tri = [[[0, 0, 0], [1, 0, 0], [0, 1, 0]]]
    none = np.zeros((0, 3, 3), np.float32)
    st = W.se3_states(100, 1, -5, 5)
    for robot, env in ((tri, none), (none, tri), (none, none)):
        sc = m.Scenario.mesh_pair(ctx, robot, env, sp, 0.5)
        assert sc.valid(st).all() and sc.link(st[:50], st[50:]).all()
    far = st.copy()
    far[:, 4:] += 1e4
    sc = m.Scenario.mesh_pair(ctx, tri, tri, sp, 0.5)
    assert sc.valid(far).all()


def test_knn_edge_cases(ctx, oracle):
    sp = m.se3_space(50, 1)
    nn = m.Nearest(ctx, sp, 64, m.KNN_BRUTE)
    pts = W.se3_states(5, 3)
    nn.insert(pts)
    q = W.se3_states(3, 4)
    idx, dist, cnt = nn.nearest(q, 8)          # k > n: padded rows
    assert (cnt == 5).all() and (idx[:, 5:] == L.NO_INDEX).all() and np.isinf(dist[:, 5:]).all()
    assert_knn_equal((idx, dist, cnt), oracle.knn(sp, pts, q, 8))
    idx, dist, cnt = nn.nearest(q, 4, radius=1e-3)  # nothing within the radius
    assert (cnt == 0).all() and (idx == L.NO_INDEX).all()
    idx, dist, cnt = nn.nearest(np.zeros((0, 7), np.float32), 4)  # empty query batch
    assert idx.shape == (0, 4)
    with pytest.raises(m.MptgError):
        nn.nearest(q, 129)
    with pytest.raises(m.MptgError):
        nn.insert(W.se3_states(100, 5))  # over capacity
    nn.close()


def test_knn_few_queries_split_scan(ctx, oracle):
    """Few queries over many points: the scan is split over CTAs and merged."""
    sp = m.se3_space(50, 1)
    pts = W.se3_states(200_000, 7)
    q = W.se3_states(24, 8)
    nn = m.Nearest(ctx, sp, 200_000, m.KNN_BRUTE)
    nn.insert(pts)
    for k in (1, 16, 49):
        assert_knn_equal(nn.nearest(q, k), oracle.tree(sp, pts).knn(q, k))
    nn.close()


def test_knn_golden(ctx):
    g = np.load(ROOT / "tests" / "golden" / "golden.npz")
    sp = m.se3_space(50, 1)
    nn = m.Nearest(ctx, sp, 4096)
    nn.insert(g["se3_pts"])
    idx, dist, _ = nn.nearest(g["se3_q"], 16)
    assert np.array_equal(idx, g["se3_knn_idx"]) and np.array_equal(dist, g["se3_knn_dist"])
    l1 = m.lp_space(8, 1)
    nn2 = m.Nearest(ctx, l1, 1024)
    nn2.insert(g["l1_pts"])
    idx, dist, _ = nn2.nearest(g["l1_q"], 5)
    assert np.array_equal(idx, g["l1_knn_idx"]) and np.array_equal(dist, g["l1_knn_dist"])
    assert np.array_equal(ctx.interpolate(sp, g["se3_q"][:32], g["se3_pts"][:32], g["interp_t"]), g["se3_interp"])


def test_knn_double_precision(ctx, oracle):
    sp = m.se3_space(50, 1, m.F64)
    pts = W.se3_states(3000, 1, dtype=np.float64)
    q = W.se3_states(100, 2, dtype=np.float64)
    nn = m.Nearest(ctx, sp, 4096, m.KNN_BRUTE)
    nn.insert(pts)
    assert_knn_equal(nn.nearest(q, 16), oracle.knn(sp, pts, q, 16))


@pytest.mark.parametrize("strategy", [m.KNN_BRUTE, m.KNN_BVH])
def test_knn_f64_planar_like_the_png_demo(ctx, oracle, strategy):
    """L2 in the plane, double precision (the PNG / holonomic demos), k as PRRT* asks for it."""
    sp = m.lp_space(2, 2, m.F64)
    pts = W.box_states(20_000, 2, 3, 0.0, [3976, 2603], np.float64)
    q = W.box_states(700, 2, 4, 0.0, [3976, 2603], np.float64)
    q[:50] = pts[:50]
    nn = m.Nearest(ctx, sp, 32768, strategy)
    nn.insert(pts)
    for k, radius in ((1, -1.0), (16, -1.0), (44, -1.0), (100, -1.0), (128, 300.0)):
        assert_knn_equal(nn.nearest(q, k, radius), oracle.knn(sp, pts, q, k, radius))
    nn.close()


def test_knn_incremental_like_a_planner(ctx, oracle):
    """Insert in waves, query between waves (the planner's access pattern)."""
    sp = m.lp_space(2, 2)
    allpts = W.box_states(6000, 2, 3, 0.0, 1000.0, np.float32)
    nn = m.Nearest(ctx, sp, 8192)
    n = 0
    for wave in (1, 31, 480, 2000, 3488):
        nn.insert(allpts[n:n + wave])
        n += wave
        q = W.box_states(257, 2, 100 + n, 0.0, 1000.0, np.float32)
        k = min(int(np.ceil(1.1 * np.e * 1.5 * np.log(n + 1))), 128)
        assert_knn_equal(nn.nearest(q, 1), oracle.knn(sp, allpts[:n], q, 1))
        assert_knn_equal(nn.nearest(q, k), oracle.knn(sp, allpts[:n], q, k))


def test_knn_merge_sharded_equals_single(ctx, oracle):
    """Tree points dealt round-robin to 4 shards (as on 4 GPUs): per-shard top-k with global indices,
    then mptg_knn_merge_dev == the single-structure answer."""
    import torch

    sp = m.se3_space(50, 1)
    pts = W.se3_states(20_000, 5)
    q = W.se3_states(512, 6)
    G, k = 4, 16
    shards = []
    for r in range(G):
        nn = m.Nearest(ctx, sp, 8192, m.KNN_BRUTE)
        nn.set_index_map(G, r)
        nn.insert(pts[r::G])
        shards.append(nn)
    dq = torch.from_numpy(q).cuda()
    idx_parts = torch.empty((G, 512, k), dtype=torch.int32, device="cuda")
    dist_parts = torch.empty((G, 512, k), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for r, nn in enumerate(shards):
        nn.nearest_dev(dq.data_ptr(), 512, k, -1.0, idx_parts[r].data_ptr(), dist_parts[r].data_ptr())
    idx = torch.empty((512, k), dtype=torch.int32, device="cuda")
    dist = torch.empty((512, k), dtype=torch.float32, device="cuda")
    cnt = torch.empty(512, dtype=torch.int32, device="cuda")
    m.knn_merge_dev(ctx, m.F32, G, 512, k, idx_parts.data_ptr(), dist_parts.data_ptr(), idx.data_ptr(), dist.data_ptr(), cnt.data_ptr())
    ctx.sync()
    got = (idx.cpu().numpy().view(np.uint32), dist.cpu().numpy(), cnt.cpu().numpy().view(np.uint32))
    assert_knn_equal(got, oracle.tree(sp, pts).knn(q, k))


def test_knn_merge_constructed_lists(ctx):
    """mptg_knn_merge_dev on constructed lists against a plain sort: 1 to 64 lists (one and two heads per lane of the
    head-selection kernel, and the insertion kernel for two or three short lists), k on both sides of every register
    layout, lists of ragged length with their empty slots at the end, distances drawn from a handful of values so that
    ties within and across lists are decided by the index, all-empty queries, zero and infinite distances; float32 and
    float64."""
    import torch

    rng = np.random.default_rng(12)
    NO = np.uint32(0xFFFFFFFF)
    for dt, scalar in ((np.float32, m.F32), (np.float64, m.F64)):
        for parts, k, Q in ((1, 16, 50), (2, 1, 64), (3, 16, 33), (4, 16, 100), (8, 36, 77), (16, 36, 40), (33, 5, 31), (64, 49, 20), (64, 128, 9),
                            (7, 64, 25), (32, 32, 17)):
            values = np.concatenate(([0.0, np.inf], rng.random(6) * 10)).astype(dt)
            idx = np.full((parts, Q, k), NO, dtype=np.uint32)
            dist = np.full((parts, Q, k), np.inf, dtype=dt)
            want_i = np.full((Q, k), NO, dtype=np.uint32)
            want_d = np.full((Q, k), np.inf, dtype=dt)
            want_c = np.zeros(Q, dtype=np.uint32)
            for q in range(Q):
                ids = rng.permutation(parts * k + 7).astype(np.uint32)  # distinct indices over the lists of a query
                rows = []
                for p_ in range(parts):
                    n = 0 if q % 11 == 0 else int(rng.integers(0, k + 1))
                    d = rng.choice(values, n) if q % 3 else (rng.random(n) * 5).astype(dt)
                    i = ids[p_ * k:p_ * k + n]
                    o = np.lexsort((i, d))
                    idx[p_, q, :n], dist[p_, q, :n] = i[o], d[o]
                    rows.append((d[o], i[o]))
                d_all, i_all = np.concatenate([r[0] for r in rows]), np.concatenate([r[1] for r in rows])
                o = np.lexsort((i_all, d_all))[:k]
                want_i[q, :o.size], want_d[q, :o.size], want_c[q] = i_all[o], d_all[o], o.size
            t_i, t_d = torch.from_numpy(idx.view(np.int32)).cuda(), torch.from_numpy(dist).cuda()
            o_i = torch.empty((Q, k), dtype=torch.int32, device="cuda")
            o_d = torch.empty((Q, k), dtype=t_d.dtype, device="cuda")
            o_c = torch.empty(Q, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            m.knn_merge_dev(ctx, scalar, parts, Q, k, t_i.data_ptr(), t_d.data_ptr(), o_i.data_ptr(), o_d.data_ptr(), o_c.data_ptr())
            ctx.sync()
            got = (o_i.cpu().numpy().view(np.uint32), o_d.cpu().numpy(), o_c.cpu().numpy().view(np.uint32))
            assert np.array_equal(got[2], want_c), (parts, k, "counts")
            assert np.array_equal(got[0], want_i), (parts, k, np.nonzero((got[0] != want_i).any(axis=1))[0][:5])
            assert np.array_equal(got[1], want_d), (parts, k, "distances")


# ------------------------------------------------------------------ BASELINE.json configs[4] at full size
def test_c5_full_size_knn_properties(ctx, oracle):
    """N = 1,048,576 tree points, Q = 65,536 queries, k = 16 (the bench workload).  Size-independent properties on the
    full wave; the two device strategies against each other on 2,048 queries; the oracle's exhaustive scan on 128 queries;
    and (r2) the oracle's box tree -- an independent exact search, held to the exhaustive scan by tests/test_oracle.py --
    on ALL 65,536 queries: indices, distances and counts bit-identical."""
    sp = m.se3_space(50, 1)
    N, Q, k = 1 << 20, 1 << 16, 16
    pts, q = W.se3_states(N, W.TREE_SEED), W.se3_states(Q, W.QUERY_SEED)
    nn = m.Nearest(ctx, sp, N, m.KNN_BVH)
    nn.insert(pts)
    idx, dist, cnt = nn.nearest(q, k)
    # every query gets k neighbours, sorted by (distance, index), all indices valid and distinct per query
    assert (cnt == k).all() and (idx < N).all()
    assert (np.diff(dist, axis=1) >= 0).all()
    ties = np.diff(dist, axis=1) == 0
    assert (np.diff(idx.astype(np.int64), axis=1)[ties] > 0).all()
    assert (np.sort(idx, axis=1)[:, 1:] != np.sort(idx, axis=1)[:, :-1]).all()
    # the reported distances ARE the metric: recompute them with the batched distance entry point
    sel = np.arange(0, Q, 97)
    for j in (0, 7, 15):
        assert np.array_equal(ctx.distance(sp, q[sel], pts[idx[sel, j]]), dist[sel, j])
    # idempotence, and independence from the rest of the wave (query ordering pass, split scans)
    idx2, dist2, _ = nn.nearest(q, k)
    assert np.array_equal(idx, idx2) and np.array_equal(dist, dist2)
    sub = np.arange(5, Q, 211)
    si, sd, _ = nn.nearest(q[sub], k)
    assert np.array_equal(si, idx[sub]) and np.array_equal(sd, dist[sub])
    # a tree point queried against the tree finds itself first (distance 0 up to the rounding of |q.q| vs 1)
    self_i, self_d, _ = nn.nearest(pts[:4096], 1)
    own = ctx.distance(sp, pts[:4096], pts[:4096])
    assert (self_d[:, 0] <= own).all() and (self_i[:, 0] == np.arange(4096)).mean() > 0.99
    # exhaustive scan of all 1M points (the other device strategy) on 2,048 queries: bit-identical
    brute = m.Nearest(ctx, sp, N, m.KNN_BRUTE)
    brute.insert(pts)
    bsel = np.arange(0, Q, 32)
    assert_knn_equal(brute.nearest(q[bsel], k), (idx[bsel], dist[bsel], cnt[bsel]))
    # and the CPU oracle's exhaustive scan on 128 queries
    osel = np.arange(3, Q, 512)
    assert_knn_equal((idx[osel], dist[osel], cnt[osel]), oracle.knn(sp, pts, q[osel], k))
    # the whole wave against the oracle's box-tree search (CPU, all host threads; about a second per 64 K queries)
    assert_knn_equal((idx, dist, cnt), oracle.tree(sp, pts).knn(q, k))
    # radius form: exactly the neighbours within r, same order
    r = float(np.median(dist[:, 7]))
    ri, rd, rc = nn.nearest(q[bsel], k, r)
    inside = dist[bsel] <= np.float32(r)
    assert np.array_equal(rc, inside.sum(axis=1).astype(np.uint32))
    assert np.array_equal(ri[inside], idx[bsel][inside]) and (ri[~inside] == m.NO_INDEX).all()


def test_c5_full_size_edge_properties(ctx, oracle):
    """E = 65,536 SE(3) edges against the bench mesh pair (~1k vs ~4k triangles)."""
    sp = m.se3_space(50, 1)
    robot, env, vmin, vmax = W.alpha_puzzle_like(env_tris_target=4000, robot_tris_target=1000)
    step = W.se3_step_size(vmin, vmax, 50.0)
    sc = m.Scenario.mesh_pair(ctx, robot, env, sp, step)
    E = 1 << 16
    a, b = W.se3_edges(E, W.EDGE_SEED, -45.0, 45.0, 12.0, 0.5)
    ok = sc.link(a, b)
    states = sc.last_stats()["states"]
    assert 0.3 < ok.mean() < 0.95
    # the schedule (work pulling, donation between warps) must not leak into the answer
    for _ in range(3):
        assert np.array_equal(sc.link(a, b), ok)
    sub = np.arange(11, E, 37)
    assert np.array_equal(sc.link(a[sub], b[sub]), ok[sub])
    # an edge is valid only if its end state is (discrete_motion_validator.hpp:75); a zero-length edge is valid(to)
    vb = sc.valid(b)
    assert not (ok.astype(bool) & ~vb.astype(bool)).any()
    assert np.array_equal(sc.link(b[:4096], b[:4096]), vb[:4096])
    # number of states touched: all of a valid edge's, at least one of an invalid edge's
    full = _dmv_state_counts(ctx, sp, a, b, step)
    assert int(full[ok == 1].sum()) + int((ok == 0).sum()) <= states <= int(full.sum())
    # the oracle on 2,048 of the edges with its near-contact pass (exhaustive, slow) ...
    osel = np.arange(0, E, 32)
    omesh = oracle.mesh_pair(robot, env, sp, step)
    want, near = omesh.link(a[osel], b[osel], with_near_contact=True)
    assert not ((ok[osel] != want) & (near == 0)).any()
    # ... and (r2) its decisions on ALL 65,536 edges; the few that differ must be flagged near contact by the product itself
    want_all = omesh.link(a, b)
    differ = np.flatnonzero(ok != want_all)
    if differ.size:
        _, flagged = sc.link(a[differ], b[differ], with_near_contact=True)
        _, onear = omesh.link(a[differ], b[differ], with_near_contact=True)
        assert (flagged == 1).all() and (onear == 1).all(), f"{differ.size} edges differ from the oracle outside the contact band"
    print(f"C5 edges: {E} compared with the oracle, {differ.size} differ (all inside the reported contact band)")


# ------------------------------------------------------------------ sampling and the device-resident PRRT
def test_device_sampler_matches_oracle_and_reference_kats(ctx, oracle):
    from tests.test_oracle import sampler_kats

    assert sampler_kats(lambda sp, lo, hi, u: m.sample_from_uniforms(ctx, sp, lo, hi, u)) == []
    cases = [
        (m.se3_space(50, 1, m.F32), [0, 0, 0, 0, -100, -100, -100], [0, 0, 0, 0, 100, 100, 100]),
        (m.se3_space(50, 1, m.F64), [0, 0, 0, 0, -45, -45, -45], [0, 0, 0, 0, 45, 45, 45]),
        (m.lp_space(2, 2, m.F64), [0, 0], [3976, 2603]),
        (m.lp_space(8, 1, m.F64), -np.pi, np.pi),
        (m.so2_space(5, 1, m.F32), 0, 0),
        (m.se2_space(1, 1, m.F32), [-5, -5, 0], [5, 5, 0]),
        (m.so3_space(m.F64), 0, 0),
    ]
    for sp, lo, hi in cases:
        for first, n in ((0, 4097), (123456789012, 1000)):
            got = m.sample(ctx, sp, lo, hi, 20261017, first, n)
            want = oracle.sample(sp, lo, hi, 20261017, first, n)
            assert np.array_equal(got, want), f"sampler differs for {sp.parts} scalar {sp.scalar}"


def test_device_prrt_replays_the_reference_loop_on_a_grid(ctx, oracle):
    """PNG-style occupancy grid, planar L2 double states (png_2d_scenario.hpp): the device-resident tree must be
    the tree the reference's addSample loop builds from the same samples -- states bit-identical, same parents,
    same goal node."""
    occ = W.synthetic_grid(500, 400, seed=2)
    sp = m.lp_space(2, 2, m.F64)
    sc, og = m.Scenario.grid(ctx, occ, m.F64), oracle.grid(occ)
    free = np.argwhere(occ == 0)
    start = free[len(free) // 7][::-1].astype(np.float64)
    goal = free[-len(free) // 9][::-1].astype(np.float64)
    lo, hi = [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1]
    for rng, waves, W_ in ((25.0, 6, 512), (float("inf"), 3, 300)):
        pl = m.DevicePRRT(sc, sp, lo, hi, range=rng, goal=goal, goal_radius=12.0, goal_bias=0.05, seed=99, capacity=8192, max_wave=512)
        pl.add_start(start)
        for _ in range(waves):
            pl.wave(W_)
        states, parents = pl.tree()
        want_states, want_parents, want_goal = kats.replay_prrt(oracle, og, sp, lo, hi, start, goal, 12.0, 0.05, rng, 99, waves, W_)
        assert pl.samples_drawn == waves * W_
        assert states.shape == want_states.shape and states.shape[0] > 50
        assert np.array_equal(states, want_states) and np.array_equal(parents, want_parents)
        assert pl.goal_node == want_goal
        if pl.solved():
            path = pl.solution()
            assert np.array_equal(path[0], start) and np.linalg.norm(path[-1] - goal) <= 12.0
            assert og.link(path[:-1], path[1:]).all()
        pl.close()


def test_device_prrt_on_meshes_builds_a_valid_tree(ctx, oracle):
    """SE(3) rigid body among meshes (se3_rigid_body_scenario.hpp): every node valid, every tree edge a valid motion
    within `range` of its parent (oracle; near-contact items excepted and counted), parents precede children, and
    the run is reproducible."""
    sp = m.se3_space(50, 1)
    robot, env, vmin, vmax = W.alpha_puzzle_like(env_tris_target=1200, robot_tris_target=400)
    step = W.se3_step_size(vmin, vmax)
    sc, og = m.Scenario.mesh_pair(ctx, robot, env, sp, step), oracle.mesh_pair(robot, env, sp, step)
    lo, hi = [0, 0, 0, 0, -45, -45, -45], [0, 0, 0, 0, 45, 45, 45]
    cand = W.se3_states(64, 5, -45.0, 45.0)
    start = cand[np.nonzero(og.valid(cand))[0][0]]
    trees = []
    for _ in range(2):
        pl = m.DevicePRRT(sc, sp, lo, hi, range=30.0, seed=7, capacity=1 << 15, max_wave=2048)
        pl.add_start(start)
        for _ in range(6):
            pl.wave(2048)
        trees.append(pl.tree())
        pl.close()
    (states, parents), (s2, p2) = trees
    assert np.array_equal(states, s2) and np.array_equal(parents, p2)
    n = states.shape[0]
    assert n > 2000 and parents[0] == m.NO_INDEX and (parents[1:] < np.arange(1, n)).all()
    child, par = states[1:], states[parents[1:]]
    ok, near = og.link(par, child, with_near_contact=True)
    assert not ((ok == 0) & (near == 0)).any()
    edge = oracle.distance(sp, par, child).astype(np.float64)
    assert edge.max() <= 30.0 * (1 + 1e-4), f"longest tree edge {edge.max()}"


def test_device_prrtstar_replays_its_wave_semantics_on_the_oracle(ctx, oracle):
    """Device-resident PRRT* (mptg_prrtstar_*) against the same wave-parallel loop restated on the oracle: states
    bit-identical, same parents, same costs, same best goal node, same number of rewires -- and the tree invariants:
    parents precede or were re-parented to later nodes without cycles, every edge is a valid motion, cost(node) =
    cost(parent) + distance up to rounding."""
    occ = W.synthetic_grid(500, 400, seed=2)
    sp = m.lp_space(2, 2, m.F64)
    sc, og = m.Scenario.grid(ctx, occ, m.F64), oracle.grid(occ)
    free = np.argwhere(occ == 0)
    start, goal = free[len(free) // 7][::-1].astype(np.float64), free[-len(free) // 9][::-1].astype(np.float64)
    lo, hi = [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1]
    # radius rewiring: r_rrg = rewireFactor (2 (1 + 1/d) measure / unit_ball)^(1/d) (rrg_rewire_neighbors.hpp:102-122), d = 2
    r_rrg = 1.1 * (2 * 1.5 * (occ.shape[1] - 1) * (occ.shape[0] - 1) / np.pi) ** 0.5
    for rng, waves, W_, rr in ((30.0, 8, 256, None), (float("inf"), 4, 200, None), (40.0, 300, 1, None), (30.0, 8, 256, r_rrg)):
        pl = m.DevicePRRTStar(sc, sp, lo, hi, range=rng, goal=goal, goal_radius=12.0, goal_bias=0.05, rewire_radius=rr, seed=31, capacity=8192,
                              max_wave=256)
        pl.add_start(start)
        for _ in range(waves):
            pl.wave(W_)
        st, pa, co = pl.tree(with_costs=True)
        ws, wp, wc, wg, wr = kats.replay_prrtstar(oracle, og, sp, lo, hi, start, goal, 12.0, 0.05, rng, 1.1, 31, waves, W_, 128, rr)
        assert st.shape[0] > 50 and np.array_equal(st, ws), (rng, st.shape, ws.shape)
        assert np.array_equal(pa, wp), f"{(pa != wp).sum()} parents differ"
        assert np.array_equal(co, wc), f"{(co != wc).sum()} costs differ, max {np.abs(co - wc).max()}"
        assert pl.goal_node == wg and pl.rewires == wr and (W_ == 1 or wr > 0)
        # invariants
        n = st.shape[0]
        assert pa[0] == m.NO_INDEX and co[0] == 0 and (pa[1:] < n).all()
        depth = np.zeros(n, dtype=np.int64)
        for i in range(1, n):  # no cycles: every node reaches the root
            a, steps = i, 0
            while a != 0:
                a, steps = int(pa[a]), steps + 1
                assert steps <= n
            depth[i] = steps
        edge = oracle.distance(sp, st[pa[1:]], st[1:])
        assert np.abs(co[pa[1:]] + edge - co[1:]).max() < 1e-9 * max(1.0, co.max())
        assert og.link(st[pa[1:]], st[1:]).all()
        if pl.solved():
            path = pl.solution()
            assert np.array_equal(path[0], start) and np.linalg.norm(path[-1] - goal) <= 12.0
            assert abs(np.linalg.norm(np.diff(path, axis=0), axis=1).sum() - pl.solution_cost()) < 1e-6
        pl.close()


def test_device_prrtstar_queued_waves_build_the_same_tree(ctx, oracle, monkeypatch):
    """Waves of up to 1,024 samples run without host synchronisation between their steps (plan.cu starWaveQueuedT: counts
    stay on the device, launches cover upper bounds); MPTG_STAR_QUEUED_MAX=0 keeps the synchronised wave.  Same seed,
    same waves: states, parents, costs, rewires and goal node must be identical -- on the grid (k nearest and radius
    rewiring) and on the 8-link arm (flat edge check behind the same calls)."""
    occ = W.synthetic_grid(500, 400, seed=2)
    free = np.argwhere(occ == 0)
    start, goal = free[len(free) // 7][::-1].astype(np.float64), free[-len(free) // 9][::-1].astype(np.float64)
    r_rrg = 1.1 * (2 * 1.5 * (occ.shape[1] - 1) * (occ.shape[0] - 1) / np.pi) ** 0.5
    lengths, radius, circles = W.link_arm_scene(8)
    cand = W.box_states(256, 8, 3, -np.pi, np.pi)
    free8 = cand[oracle.link_arm(lengths, radius, circles).valid(cand) != 0]
    cases = [
        ("grid", lambda: m.Scenario.grid(ctx, occ, m.F64), m.lp_space(2, 2, m.F64), [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1], start, goal, 12.0, 30.0, None,
         (64, 128, 256, 512, 512, 1024, 37, 1, 512)),
        ("grid, radius", lambda: m.Scenario.grid(ctx, occ, m.F64), m.lp_space(2, 2, m.F64), [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1], start, goal, 12.0, 30.0,
         r_rrg, (64, 128, 256, 256, 256)),
        ("arm", lambda: m.Scenario.link_arm(ctx, lengths, radius, circles, m.F64), m.lp_space(8, 1, m.F64), [-np.pi] * 8, [np.pi] * 8, free8[0],
         free8[1], 0.5, 2.0, None, (64, 128, 256, 512, 512)),
    ]
    for name, make, sp, lo, hi, s0, g, g_rad, rng, rr, waves in cases:
        trees = []
        for queued_max in ("0", "1024"):
            monkeypatch.setenv("MPTG_STAR_QUEUED_MAX", queued_max)
            sc = make()
            pl = m.DevicePRRTStar(sc, sp, lo, hi, range=rng, goal=g, goal_radius=g_rad, goal_bias=0.05, rewire_radius=rr, seed=5, capacity=1 << 13,
                                  max_wave=1024)
            pl.add_start(s0)
            for w in waves:
                pl.wave(w)
            trees.append(pl.tree(with_costs=True) + (pl.rewires, pl.goal_node, pl.samples_drawn))
            pl.close()
            sc.close()
        (st, pa, co, rew, gn, drawn), other = trees
        assert st.shape[0] > 100 and (rew > 0 or name == "arm"), (name, st.shape, rew)  # (8-D: few offers lower a cost this early)
        assert np.array_equal(st, other[0]) and np.array_equal(pa, other[1]) and np.array_equal(co, other[2]), name
        assert (rew, gn, drawn) == other[3:], name


def test_device_prrtstar_and_pprm_on_meshes_se3_f32(ctx, oracle):
    """SE(3) rigid body among meshes, float32 states (BASELINE configs[2]): the device-resident PRRT* and PPRM build
    graphs whose every node is valid and every edge a valid motion on the oracle (near-contact items excepted and
    counted), PRRT* costs are consistent and runs are reproducible."""
    sp = m.se3_space(50, 1)
    robot, env, vmin, vmax = W.alpha_puzzle_like(env_tris_target=1200, robot_tris_target=400)
    step = W.se3_step_size(vmin, vmax)
    sc, og = m.Scenario.mesh_pair(ctx, robot, env, sp, step), oracle.mesh_pair(robot, env, sp, step)
    lo, hi = [0, 0, 0, 0, -45, -45, -45], [0, 0, 0, 0, 45, 45, 45]
    cand = W.se3_states(64, 5, -45.0, 45.0)
    free = cand[np.nonzero(og.valid(cand))[0]]
    start, goal = free[0], free[1]
    runs = []
    for _ in range(2):
        pl = m.DevicePRRTStar(sc, sp, lo, hi, range=30.0, goal=goal, goal_radius=8.0, goal_bias=0.05, seed=7, capacity=1 << 14, max_wave=1024)
        pl.add_start(start)
        for _ in range(5):
            pl.wave(1024)
        runs.append(pl.tree(with_costs=True) + (pl.rewires, pl.goal_node))
        pl.close()
    (st, pa, co, rew, gn), second = runs
    assert all(np.array_equal(a, b) for a, b in zip((st, pa, co), second[:3])) and (rew, gn) == second[3:]
    n = st.shape[0]
    assert n > 1000 and rew > 0 and pa[0] == m.NO_INDEX and (pa[1:] < n).all()
    ok, near = og.link(st[pa[1:]], st[1:], with_near_contact=True)
    assert not ((ok == 0) & (near == 0)).any()
    edge = oracle.distance(sp, st[pa[1:]], st[1:]).astype(np.float64)
    # (tree edges may be longer than `range`: parents and rewiring come from the k nearest, prrt_star.hpp:553-556)
    assert np.abs(co[pa[1:]].astype(np.float64) + edge - co[1:]).max() < 2e-4 * max(1.0, float(co.max()))  # float32 costs
    for i in range(1, n, 7):  # no cycles
        a, steps = i, 0
        while a != 0:
            a, steps = int(pa[a]), steps + 1
            assert steps <= n
    # PPRM on the same scene
    pp = m.DevicePPRM(sc, sp, lo, hi, goal=goal, goal_radius=8.0, seed=9, capacity=1 << 13, max_wave=512)
    assert pp.add_start(start) == 0 and pp.add_goal(goal) == 1
    for _ in range(4):
        pp.wave(512)
    gs, ei, ed, mk, cp = pp.graph()
    assert gs.shape[0] > 300 and og.valid(gs).all()
    rows, cols = np.nonzero(ei != m.NO_INDEX)
    ok, near = og.link(gs[rows], gs[ei[rows, cols]], with_near_contact=True)
    assert rows.size > 1000 and not ((ok == 0) & (near == 0)).any()
    assert np.array_equal(ed[rows, cols], oracle.distance(sp, gs[rows], gs[ei[rows, cols]]))
    want = kats.components(ei)
    assert np.array_equal(cp, want)
    pp.close()


def test_device_planner_argument_errors(ctx):
    """Misuse is an error, never a silent fallback: waves before a start, waves larger than max_wave, a full tree."""
    occ = W.synthetic_grid(200, 150, seed=3)
    sp = m.lp_space(2, 2, m.F64)
    sc = m.Scenario.grid(ctx, occ, m.F64)
    free = np.argwhere(occ == 0)[0][::-1].astype(np.float64)
    for cls in (m.DevicePRRT, m.DevicePRRTStar):
        pl = cls(sc, sp, [0, 0], [199, 149], range=20.0, seed=1, capacity=64, max_wave=128)
        with pytest.raises(m.MptgError):
            pl.wave(16)  # "there are no valid initial states" (prrt.hpp:197-198)
        pl.add_start(free)
        with pytest.raises(m.MptgError):
            pl.wave(129)
        for _ in range(6):
            pl.wave(128)
        assert pl.size == 64  # capacity: the tree stops growing, no overrun
        st, pa = pl.tree()
        assert (pa[1:] < 64).all()
        pl.close()
    pp = m.DevicePPRM(sc, sp, [0, 0], [199, 149], seed=1, capacity=64, max_wave=128)
    with pytest.raises(m.MptgError):
        pp.wave(129)
    obstacle = np.argwhere(occ != 0)[0][::-1].astype(np.float64)
    assert pp.add_start(obstacle) == m.NO_INDEX  # invalid state: rejected (pprm.hpp:299-300)
    assert pp.add_start(free) == 0
    for _ in range(4):
        pp.wave(128)
    assert pp.size == 64
    pp.close()


def _check_device_pprm(ctx, oracle, sp, sc, og, lo, hi, start, goal, goal_radius, seed, waves, W, spanner_stretch=None, spanner_capacity=0):
    pl = m.DevicePPRM(sc, sp, lo, hi, goal=goal, goal_radius=goal_radius, seed=seed, capacity=1 << 14, max_wave=W,
                      spanner_stretch=spanner_stretch or 0.0, spanner_capacity=spanner_capacity)
    assert pl.add_start(start) == 0 and pl.add_goal(goal) == 1
    assert pl.add_goal(goal) == m.NO_INDEX  # closer than epsilon to a node: rejected (pprm.hpp:306-308)
    for _ in range(waves):
        pl.wave(W)
    st, ei, ed, mk, cp = pl.graph()
    ws, wi, wd, wm = kats.replay_pprm(oracle, og, sp, lo, hi, [start], [goal], goal, goal_radius, seed, waves, W, pl.row_stride, spanner_stretch)
    assert pl.samples_drawn == waves * W and st.shape[0] > 100
    assert np.array_equal(st, ws) and np.array_equal(ei, wi) and np.array_equal(ed, wd) and np.array_equal(mk, wm)
    # components: same partition as a host union-find over the same edges; solved() <=> start and goal connected
    want = kats.components(wi)
    assert np.array_equal(cp == cp[:, None], want == want[:, None]) if st.shape[0] <= 2048 else np.array_equal(cp, want)
    starts, goals = np.nonzero(mk & 1)[0], np.nonzero(mk & 2)[0]
    assert pl.solved() == bool(np.isin(want[goals], want[starts]).any())
    if pl.solved():
        path = pl.solution()
        assert np.array_equal(path[0], np.asarray(start, dtype=sp.dtype)) and (mk[np.nonzero((st == path[-1]).all(axis=1))[0][0]] & 2)
        assert og.link(path[:-1], path[1:]).all()
    n_edges = int((ei != m.NO_INDEX).sum())
    pl.close()
    return st.shape[0], n_edges


def test_device_pprm_replays_the_reference_loop(ctx, oracle):
    """Device-resident PPRM (mptg_pprm_*): the roadmap after every wave equals the one PPRM's addSample loop builds on the
    oracle from the same samples -- states, edge rows (neighbour, distance), marks, components, solved()."""
    # occupancy grid, planar L2 doubles (BASELINE configs[1] geometry)
    occ = W.synthetic_grid(500, 400, seed=2)
    sp = m.lp_space(2, 2, m.F64)
    free = np.argwhere(occ == 0)
    start, goal = free[len(free) // 7][::-1].astype(np.float64), free[-len(free) // 9][::-1].astype(np.float64)
    n, e = _check_device_pprm(ctx, oracle, sp, m.Scenario.grid(ctx, occ, m.F64), oracle.grid(occ), [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1],
                              start, goal, 1e-6, 5, 5, 256)
    assert e > 3 * n
    # 8-link arm, L1 over [-pi, pi]^8 (BASELINE configs[3]: PPRM for the link manipulator)
    lengths, radius, circles = W.link_arm_scene(8)
    sp8 = m.lp_space(8, 1, m.F64)
    arm, oarm = m.Scenario.link_arm(ctx, lengths, radius, circles, m.F64), oracle.link_arm(lengths, radius, circles)
    cand = W.box_states(256, 8, 3, -np.pi, np.pi)
    free8 = cand[oarm.valid(cand) != 0]
    _check_device_pprm(ctx, oracle, sp8, arm, oarm, -np.pi, np.pi, free8[0], free8[1], 1e-6, 11, 4, 200)


def test_device_pprm_irs_replays_the_reference_spanner(ctx, oracle):
    """PPRM-IRS on the device (mptg_pprm_set_spanner): the SPARSE roadmap after every wave equals the one the reference's
    addSample / addEdge / ShortestPathCheck build on the oracle from the same samples (impl/pprm_irs/pprm_irs.hpp:300-368,
    shortest_path_check.hpp:111-225) -- states, sparse edge rows, distances, marks, components, solved(); waves of one sample
    (the reference's own order of events) and of 128 / 200 samples; stretch 5 (the default) and 2."""
    occ = W.synthetic_grid(500, 400, seed=2)
    sp = m.lp_space(2, 2, m.F64)
    free = np.argwhere(occ == 0)
    start, goal = free[len(free) // 7][::-1].astype(np.float64), free[-len(free) // 9][::-1].astype(np.float64)
    grid, ogrid = m.Scenario.grid(ctx, occ, m.F64), oracle.grid(occ)
    lo, hi = [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1]
    n1, e1 = _check_device_pprm(ctx, oracle, sp, grid, ogrid, lo, hi, start, goal, 1e-6, 5, 300, 1, spanner_stretch=5.0)
    n2, e2 = _check_device_pprm(ctx, oracle, sp, grid, ogrid, lo, hi, start, goal, 1e-6, 5, 6, 128, spanner_stretch=5.0)
    n3, e3 = _check_device_pprm(ctx, oracle, sp, grid, ogrid, lo, hi, start, goal, 1e-6, 5, 6, 128, spanner_stretch=2.0)
    n0, e0 = _check_device_pprm(ctx, oracle, sp, grid, ogrid, lo, hi, start, goal, 1e-6, 5, 6, 128)
    assert n0 == n2 == n3 and e2 < e3 < e0 and e2 < 4 * n2  # the spanner is sparse; a smaller stretch keeps more
    print(f"PPRM-IRS device: {n2} nodes, sparse edges {e2} (stretch 5) / {e3} (stretch 2) of {e0} valid ones; one-sample waves {n1} nodes, {e1} edges")
    lengths, radius, circles = W.link_arm_scene(8)
    sp8 = m.lp_space(8, 1, m.F64)
    arm, oarm = m.Scenario.link_arm(ctx, lengths, radius, circles, m.F64), oracle.link_arm(lengths, radius, circles)
    cand = W.box_states(256, 8, 3, -np.pi, np.pi)
    free8 = cand[oarm.valid(cand) != 0]
    # 8 dimensions: a search within 5 x the neighbour distance reaches most of the roadmap; started with room for 64 labelled
    # nodes per search, the waves are rerun with more storage until every search fits -- same roadmap
    _check_device_pprm(ctx, oracle, sp8, arm, oarm, -np.pi, np.pi, free8[0], free8[1], 1e-6, 11, 4, 200, spanner_stretch=5.0, spanner_capacity=64)
    _check_device_pprm(ctx, oracle, sp, grid, ogrid, lo, hi, start, goal, 1e-6, 5, 6, 128, spanner_stretch=5.0, spanner_capacity=64)


# ------------------------------------------------------------------ grid / shapes / link arm
@pytest.mark.parametrize("scalar", [m.F64, m.F32])
def test_grid_matches_oracle(ctx, oracle, scalar):
    occ = W.synthetic_grid(1000, 700, seed=4, n_blobs=125)
    dt = np.float64 if scalar == m.F64 else np.float32
    sc = m.Scenario.grid(ctx, occ, scalar)
    og = oracle.grid(occ, scalar)
    st = W.box_states(20000, 2, 8, 0.0, [1000, 700], dt)
    st[:4] = [[999.6, 10], [10, 699.6], [999.6, 699.6], [0, 0]]  # x == width wrap, out of range
    assert np.array_equal(sc.valid(st), og.valid(st))
    for max_len in (None, 64.0, 5.0, 0.5):
        a, b = W.grid_edges(8192, 1000, 700, 9, max_len, dt)
        got, want = sc.link(a, b), og.link(a, b)
        assert np.array_equal(got, want), f"max_len={max_len}: {(got != want).sum()} differ"
        if max_len == 5.0:
            assert 0.2 < want.mean() < 0.8


def test_grid_golden_and_empty(ctx):
    g = np.load(ROOT / "tests" / "golden" / "golden.npz")
    sc = m.Scenario.grid(ctx, g["grid_occ"])
    assert np.array_equal(sc.link(g["grid_a"], g["grid_b"]), g["grid_link"])
    assert sc.link(np.zeros((0, 2)), np.zeros((0, 2))).shape == (0,)
    assert sc.valid(np.zeros((0, 2))).shape == (0,)


@pytest.mark.parametrize("scalar", [m.F64, m.F32])
def test_shapes_match_oracle(ctx, oracle, scalar):
    dt = np.float64 if scalar == m.F64 else np.float32
    # the holonomic demo scene (demo/holonomic_2d_point_planning.cpp:67-79)
    centres, radii = [[170, 140], [800, 70], [900, 380]], [80, 50, 70]
    rects = [[375, 140, 520, 220], [200, 320, 390, 390], [600, 200, 680, 450]]
    sc = m.Scenario.shapes(ctx, 2, centres, radii, rects, scalar)
    og = oracle.shapes(2, centres, radii, rects, scalar)
    st = W.box_states(20000, 2, 3, 0.0, [1024, 512], dt)
    assert np.array_equal(sc.valid(st), og.valid(st))
    for max_len in (None, 100.0, 8.0):
        a, b = W.grid_edges(8192, 1024, 512, 5, max_len, dt)
        assert np.array_equal(sc.link(a, b), og.link(a, b))
    # the 3-D sphere scenario of test/planner_integration_test.hpp:128-150
    r = float(np.sqrt(2.0) * 0.95)
    sc3 = m.Scenario.shapes(ctx, 3, [[0, 0, 0]], [r], (), scalar)
    og3 = oracle.shapes(3, [[0, 0, 0]], [r], (), scalar)
    a, b = W.box_states(8192, 3, 1, -1.0, 1.0, dt), W.box_states(8192, 3, 2, -1.0, 1.0, dt)
    assert np.array_equal(sc3.valid(a), og3.valid(a))
    assert np.array_equal(sc3.link(a, b), og3.link(a, b))
    sc3b = m.Scenario.shapes(ctx, 3, [[0, 0, 0], [0.5, 0.5, 0.5]], [0.5, 0.2], (), scalar)
    og3b = oracle.shapes(3, [[0, 0, 0], [0.5, 0.5, 0.5]], [0.5, 0.2], (), scalar)
    got = sc3b.link(a, b)
    assert np.array_equal(got, og3b.link(a, b)) and 0.05 < got.mean() < 0.95


@pytest.mark.parametrize("n_links", [5, 8, 16, 32])
@pytest.mark.parametrize("scalar", [m.F64, m.F32])
def test_linkarm_matches_oracle(ctx, oracle, n_links, scalar):
    dt = np.float64 if scalar == m.F64 else np.float32
    if n_links == 5:  # the shipped demo (demo/link_manipulator_planning.cpp:62-76)
        lengths, radius = [10.0, 12.0, 8.0, 6.0, 4.0], 0.5
        circles = [[20, -20, 8], [-20, -30, 5], [0, 25, 10], [30, 10, 10], [-30, 10, 8]]
    else:
        lengths, radius, circles = W.link_arm_scene(n_links)
    sc = m.Scenario.link_arm(ctx, lengths, radius, circles, scalar)
    og = oracle.link_arm(lengths, radius, circles, scalar)
    st = W.box_states(8192, n_links, 3, -np.pi, np.pi, dt)
    gv = sc.valid(st)
    assert np.array_equal(gv, og.valid(st))
    for delta in (0.5, 0.05, 3.0):
        a, b = W.arm_edges(4096, n_links, 7, delta, dt)
        got, want = sc.link(a, b), og.link(a, b)
        assert np.array_equal(got, want), f"delta={delta}: {(got != want).sum()} differ"


def _reach_boundary_scenes(dt):
    """Arm scenes with ONE circle each, its centre at (reach + radius) * (1 + delta) from a joint: delta from -1e-3 to +1e-3
    through 0 and a few ulps, along, against and across the link or in a random direction; every sixth arm is 100 times
    longer, so that the coordinates' rounding errors are large against the margins."""
    eps = np.finfo(dt).eps
    rng = np.random.default_rng(77)
    for scene in range(20):
        n_links = int(rng.integers(1, 6))
        lengths = rng.uniform(0.5, 6.0, n_links) * (100.0 if scene % 6 == 5 else 1.0)
        radius = float(rng.uniform(0.05, 0.8))
        pose = rng.uniform(-np.pi, np.pi, n_links)
        ang = np.cumsum(pose)
        joints = np.concatenate([[[0.0, 0.0]], np.cumsum(np.stack([lengths * np.cos(ang), lengths * np.sin(ang)], axis=1), axis=0)])
        for d in (-1e-3, -64 * eps, -eps, 0.0, eps, 64 * eps, 1024 * eps, 4096 * eps, 1e-3):
            i = int(rng.integers(0, n_links))
            r = float(rng.uniform(0.1, 2.0))
            along = (joints[i + 1] - joints[i]) / lengths[i]
            direction = [along, -along, np.array([-along[1], along[0]]), rng.normal(size=2)][int(rng.integers(0, 4))]
            direction = direction / np.linalg.norm(direction)
            c = joints[i] + direction * (lengths[i] + r + radius) * (1.0 + d)
            st = np.concatenate([pose[None, :], pose[None, :] + rng.normal(scale=1e-7, size=(63, n_links)),
                                 pose[None, :] + rng.normal(scale=1e-3, size=(64, n_links)),
                                 pose[None, :] + rng.normal(scale=0.2, size=(128, n_links))]).astype(dt)
            yield lengths, radius, [[c[0], c[1], r]], st


@pytest.mark.parametrize("scalar", [m.F64, m.F32])
def test_linkarm_circles_at_the_reach_boundary(ctx, oracle, scalar):
    """The arm validator passes circles farther from a link's start than the link's reach without the exact segment test;
    on scenes that sit on that boundary its decisions must still equal the oracle's (states and edges)."""
    dt = np.float64 if scalar == m.F64 else np.float32
    seen = set()
    for n, (lengths, radius, circles, st) in enumerate(_reach_boundary_scenes(dt)):
        sc = m.Scenario.link_arm(ctx, lengths, radius, circles, scalar)
        og = oracle.link_arm(lengths, radius, circles, scalar)
        want = og.valid(st)
        seen.update(int(x) for x in want[:128])
        assert np.array_equal(sc.valid(st), want), f"scene {n}"
        assert np.array_equal(sc.link(st[:128], st[128:]), og.link(st[:128], st[128:])), f"scene {n} (edges)"
        sc.close()
    assert seen == {0, 1}


def test_fp32_probe(ctx):
    """mptg_probe_fp32_tflops: the FFMA yardstick of bench.py's fp32 rooflines lands near the nominal 74.4 TFLOP/s of a B200."""
    t = ctx.probe_fp32_tflops()
    assert 40.0 < t < 80.0, t


def test_linkarm_golden(ctx):
    g = np.load(ROOT / "tests" / "golden" / "golden.npz")
    sc = m.Scenario.link_arm(ctx, g["arm_lengths"], float(g["arm_radius"]), g["arm_circles"])
    assert np.array_equal(sc.link(g["arm_a"], g["arm_b"]), g["arm_link"])


# ------------------------------------------------------------------ mesh
def _mesh_scene(env_t=1200, robot_t=400):
    robot, env, vmin, vmax = W.alpha_puzzle_like(env_tris_target=env_t, robot_tris_target=robot_t)
    return robot, env, W.se3_step_size(vmin, vmax)


def test_mesh_valid_matches_oracle(ctx, oracle):
    sp = m.se3_space(50, 1)
    robot, env, step = _mesh_scene()
    sc = m.Scenario.mesh_pair(ctx, robot, env, sp, step)
    og = oracle.mesh_pair(robot, env, sp, step)
    st = W.se3_states(6000, 21, -45.0, 45.0)
    got = sc.valid(st)
    want, margin = og.valid(st, with_margin=True)
    near = np.abs(margin) < 1e-6 * float(np.linalg.norm(env.reshape(-1, 3).max(0) - env.reshape(-1, 3).min(0)))
    diff = got != want
    print(f"mesh valid: {want.mean():.3f} free, {int(near.sum())} states within the contact band, {int(diff.sum())} differ")
    assert not (diff & ~near).any(), f"{int((diff & ~near).sum())} decisions differ outside the near-contact band"
    assert 0.3 < want.mean() < 0.97
    stats = sc.last_stats()
    assert stats["states"] == 6000 and stats["bv_tests"] > 0 and stats["prim_tests"] > 0


def test_mesh_link_matches_oracle(ctx, oracle):
    sp = m.se3_space(50, 1)
    robot, env, step = _mesh_scene()
    sc = m.Scenario.mesh_pair(ctx, robot, env, sp, step)
    og = oracle.mesh_pair(robot, env, sp, step)
    for max_trans, max_angle in ((12.0, 0.5), (40.0, 2.0), (0.5, 0.01)):
        a, b = W.se3_edges(1500, 23, -45.0, 45.0, max_trans, max_angle)
        got = sc.link(a, b)
        want, near = og.link(a, b, with_near_contact=True)
        diff = got != want
        print(f"mesh link {max_trans}/{max_angle}: {want.mean():.3f} valid, {int(near.sum())} near-contact edges, {int(diff.sum())} differ")
        assert not (diff & (near == 0)).any()
        # valid edges touch exactly the reference's set of states (`to` + steps-1 interior ones); an invalid edge
        # touches between one state and its whole set (the visiting order is coarse-to-fine, not the reference's)
        if not diff.any():
            got_states = sc.last_stats()["states"]
            full = _dmv_state_counts(ctx, sp, a, b, step)
            assert int(full[want == 1].sum()) + int((want == 0).sum()) <= got_states <= int(full.sum())


def _dmv_state_counts(ctx, sp, a, b, step):
    """states of each edge under discrete_motion_validator.hpp:78 -- max(1, ceil(distance * (1/stepSize)))"""
    dist = ctx.distance(sp, a, b).astype(np.float32)
    steps = np.ceil(dist * (np.float32(1.0) / np.float32(step))).astype(np.int64)
    return np.maximum(steps, 1)


def test_mesh_float64_states_match_oracle(ctx, oracle):
    """SE3RigidBodyScenario<double> (the demos build both scalar types, demo/CMakeLists.txt:39): the edge
    discretisation runs in double (steps, interpolation parameters, slerp), the collision test in float on the
    rounded pose; the double-precision oracle may only disagree inside its near-contact band."""
    sp = m.se3_space(50, 1, m.F64)
    robot, env, step = _mesh_scene()
    sc = m.Scenario.mesh_pair(ctx, robot, env, sp, step)
    og = oracle.mesh_pair(robot, env, sp, step)
    st = W.se3_states(4000, 27, -45.0, 45.0, dtype=np.float64)
    got = sc.valid(st)
    want, margin = og.valid(st, with_margin=True)
    near = np.abs(margin) < 1e-6 * float(np.linalg.norm(env.reshape(-1, 3).max(0) - env.reshape(-1, 3).min(0)))
    assert not ((got != want) & ~near).any()
    assert 0.3 < want.mean() < 0.97
    for max_trans, max_angle in ((12.0, 0.5), (40.0, 2.0)):
        a, b = W.se3_edges(1200, 29, -45.0, 45.0, max_trans, max_angle, dtype=np.float64)
        got = sc.link(a, b)
        want, nearc = og.link(a, b, with_near_contact=True)
        diff = got != want
        print(f"mesh link f64 {max_trans}/{max_angle}: {want.mean():.3f} valid, {int(nearc.sum())} near-contact edges, {int(diff.sum())} differ")
        assert not (diff & (nearc == 0)).any()
        assert 0.05 < want.mean() < 0.98
        if not diff.any():  # valid edges touch exactly ceil(distance / step) states, computed in double
            full = np.maximum(np.ceil(ctx.distance(sp, a, b) * (1.0 / step)).astype(np.int64), 1)
            assert int(full[want == 1].sum()) + int((want == 0).sum()) <= sc.last_stats()["states"] <= int(full.sum())


def test_mesh_golden_and_degenerate(ctx):
    g = np.load(ROOT / "tests" / "golden" / "golden.npz")
    sp = m.se3_space(50, 1)
    sc = m.Scenario.mesh_pair(ctx, g["mesh_robot"], g["mesh_env"], sp, float(g["mesh_step"]))
    assert np.array_equal(sc.valid(g["mesh_states"]), g["mesh_valid"])
    assert np.array_equal(sc.link(g["mesh_a"], g["mesh_b"]), g["mesh_link"])
    # zero-length edge and an empty batch
    a = g["mesh_a"][:8]
    assert np.array_equal(sc.link(a, a), sc.valid(a))
    assert sc.link(a[:0], a[:0]).shape == (0,)
    # single triangles
    tri = m.Scenario.mesh_pair(ctx, [[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], [[[0.2, 0.2, -0.5], [0.2, 0.2, 0.5], [0.8, 0.8, 0.5]]], sp, 0.5)
    st = np.array([[0, 0, 0, 1, 0, 0, 0], [0, 0, 0, 1, 0, 0, 2.0]], dtype=np.float32)
    assert list(tri.valid(st)) == [0, 1]


# ------------------------------------------------------------------ contact band ("0 near contact" everywhere)
def test_mesh_constructed_contact_cases(ctx, oracle):
    """Touching configurations built so that float arithmetic is exact (tests/contact_cases.py): vertex on face, edge on
    edge, edge in face, coplanar overlap, coplanar corner, vertex on vertex, each displaced by 0, +-1, +-16, +-1024,
    +-4096 ulp along the separating direction.  For float states the kernels and the oracle run the same arithmetic, so
    the decisions must be equal INSIDE the band too, equal to the answer known by construction, and the exported
    near-contact flags must cover every state the oracle places inside the band."""
    from tests import contact_cases as cc

    sp = m.se3_space(50, 1)
    total_near = 0
    for name, case in cc.CASES.items():
        robot = np.array([case["robot"]], np.float32)
        sc = m.Scenario.mesh_pair(ctx, robot, cc.ENV, sp, 0.5)
        og = oracle.mesh_pair(robot, cc.ENV, sp, 0.5)
        assert abs(sc.contact_band() - cc.BAND) < 1e-6 * cc.BAND
        st = cc.states_for(case)
        ok, near = sc.valid(st, with_near_contact=True)
        want, margin = og.valid(st, with_margin=True)
        onear = np.abs(margin) < cc.BAND
        print(f"contact {name}: collide {list(1 - ok)}, near (gpu) {list(near)}, near (oracle) {list(onear.astype(int))}")
        assert np.array_equal(ok, want) and np.array_equal(1 - ok, cc.expected_contact(name)), name
        assert np.array_equal(sc.valid(st), ok)                     # the plain kernel decides the same
        assert (near[onear] == 1).all(), name                        # the export covers the oracle's band
        assert (near[cc.in_band()] == 1).all(), name                 # +-16 ulp of touching is inside the band
        total_near += int(near.sum())
        # as edges from a far, free state: `to` is always checked (discrete_motion_validator.hpp:75)
        far = np.tile(np.array([[0, 0, 0, 1, 2.0, 2.0, 9.0]], np.float32), (len(st), 1))
        eok, enear = sc.link(far, st, with_near_contact=True)
        wok, wnear = og.link(far, st, with_near_contact=True)
        assert np.array_equal(eok, wok), name
        assert (enear[(wnear == 1) & (eok == 1)] == 1).all(), name   # valid edges: every state was examined
        assert (eok[cc.expected_contact(name) == 1] == 0).all(), name
    assert total_near > 0
    # nothing near: flags stay 0
    sc = m.Scenario.mesh_pair(ctx, np.array([cc.CASES["vertex_on_face"]["robot"]], np.float32), cc.ENV, sp, 0.5)
    ok, near = sc.valid(np.array([[0, 0, 0, 1, 2.0, 2.0, 9.0]], np.float32), with_near_contact=True)
    assert list(ok) == [1] and list(near) == [0]


def _first_contact(og, start, direction, lo=0.0, hi=1.0, iters=60):
    """bisection on the oracle: largest s in [lo, hi] with state(start + s * direction) free (translations only)"""
    def at(s):
        q = start.copy()
        q[4:7] += s * direction
        return q
    assert og.valid(at(lo)[None])[0] == 1 and og.valid(at(hi)[None])[0] == 0
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        if og.valid(at(mid)[None])[0] == 1:
            lo = mid
        else:
            hi = mid
    return lo, hi, at


@pytest.mark.parametrize("scalar", [m.F32, m.F64])
def test_mesh_sliding_through_first_contact(ctx, oracle, scalar):
    """Generic rotations: the robot is pushed along random directions until the oracle reports first contact (bisection
    to the last bit), then states are laid out at -1e-3 ... +1e-3 of that point, densest around it.  Double states are
    where the kernels (float collision test on the rounded pose) and the double-precision oracle may disagree: only
    inside the band, and every disagreement must carry the exported flag."""
    dt = np.float32 if scalar == m.F32 else np.float64
    sp = m.se3_space(50, 1, scalar)
    robot, env, step = _mesh_scene()
    sc = m.Scenario.mesh_pair(ctx, robot, env, sp, step)
    og = oracle.mesh_pair(robot, env, sp, step)
    band = sc.contact_band()
    rng = np.random.default_rng(31)
    cand = W.se3_states(400, 33, -45.0, 45.0, dtype=dt)
    free = cand[og.valid(cand) == 1]
    states, n_contacts = [], 0
    for q0 in free[:40]:
        d = rng.normal(size=3)
        d = (d / np.linalg.norm(d) * 120.0).astype(dt)
        end = q0.copy()
        end[4:7] += d
        if og.valid(end[None])[0] == 1:
            # look for a colliding point along the ray
            ss = np.linspace(0, 1, 65)[1:]
            hits = [s for s in ss if og.valid((q0 + np.concatenate([np.zeros(4, dt), (s * d).astype(dt)]))[None])[0] == 0]
            if not hits:
                continue
            hi = hits[0]
        else:
            hi = 1.0
        lo, hi, at = _first_contact(og, q0.astype(dt), d, 0.0, hi)
        n_contacts += 1
        for off in (0.0, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3):
            for sgn in (-1.0, 1.0):
                states.append(at(lo + sgn * off / 120.0))   # off in length units along the ray
    st = np.asarray(states, dtype=dt)
    assert n_contacts >= 10
    ok, near = sc.valid(st, with_near_contact=True)
    want, margin = og.valid(st, with_margin=True)
    onear = np.abs(margin) < band
    diff = ok != want
    print(f"first contact ({'f32' if scalar == m.F32 else 'f64'}): {len(st)} states around {n_contacts} contacts, {int(onear.sum())} inside the band "
          f"(oracle), {int(near.sum())} flagged by the kernel, {int(diff.sum())} decisions differ")
    assert onear.sum() > 0 and near.sum() > 0
    assert not (diff & ~onear).any()          # outside the band: identical
    assert (near[diff] == 1).all()            # a differing decision is always a reported one
    if scalar == m.F32:
        assert not diff.any()                 # same arithmetic: identical inside the band too
    assert np.array_equal(sc.valid(st), ok)
    # the same as edge end points
    a = np.repeat(free[:1], len(st), axis=0).astype(dt)
    eok, enear = sc.link(a, st, with_near_contact=True)
    wok, wnear = og.link(a, st, with_near_contact=True)
    ediff = eok != wok
    print(f"   as edges: {int(wnear.sum())} near-contact edges (oracle), {int(enear.sum())} flagged, {int(ediff.sum())} differ")
    assert not (ediff & (wnear == 0)).any()
    assert (enear[ediff] == 1).all()
