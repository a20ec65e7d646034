"""The reference's own known-answer tests for the hot-path arithmetic,
transcribed as data so that the CPU oracle AND the CUDA kernels are checked against the same table.

Each entry cites the reference test it restates (paths relative to the reference root).  Expected
values are computed with the same double-precision expressions the reference test writes
(Python floats are IEEE doubles).  `tol` None means exact equality, as in the reference's EXPECT ==.
"""
from __future__ import annotations

import math

import numpy as np

import mpt_b200 as m

F64 = m.F64
PI = math.pi


def angle_axis(angle, axis):
    """Eigen::AngleAxisd(angle, axis) -> quaternion coeffs (x,y,z,w)."""
    s, c = math.sin(angle / 2), math.cos(angle / 2)
    return [axis[0] * s, axis[1] * s, axis[2] * s, c]


def to_angle_axis(q):
    """Eigen::AngleAxisd(Quaterniond) -> (angle, axis)."""
    n = math.sqrt(q[0] ** 2 + q[1] ** 2 + q[2] ** 2)
    angle = 2 * math.atan2(n, abs(q[3]))
    if q[3] < 0:
        n = -n
    return angle, [q[0] / n, q[1] / n, q[2] / n]


AXIS = [v / math.sqrt(14.0) for v in (-1.0, 2.0, 3.0)]
SQRT20 = math.sqrt(0.0 + 2.0 * 2.0 + 4.0 * 4.0)


def distance_kats():
    """-> list of (name, space, a, b, expected, tol)"""
    k = []
    k.append(("lp_space_test.cpp:39-50", m.lp_space(3, 2, F64), [1, 2, 3], [1, 0, -1], SQRT20, None))
    k.append(("scaled_space_test.cpp:40-52", m.lp_space(3, 2, F64, 5.0 / 2.0), [1, 2, 3], [1, 0, -1], SQRT20 * 5 / 2, None))
    so2 = m.so2_space(1, 1, F64)
    for a, b, e in ((1.0, 1.0, 0.0), (0.0, 2.0, 2.0), (2.0, 0.0, 2.0), (-1.0, 3.0, 2 * PI - 4.0), (3.0, -1.0, 2 * PI - 4.0)):
        k.append(("so2_space_test.cpp:46-50", so2, [a], [b], e, None))
    k.append(("so2_space_test.cpp:54-65", m.so2_space(3, 1, F64), [1, 2, 3], [1, 0, -1], 0.0 + 2.0 + 2 * PI - 4, None))
    k.append(("so3_space_test.cpp:39-56", m.so3_space(F64), angle_axis(-1.0, AXIS), angle_axis(2.0, AXIS), 3.0 / 2, 1e-10))
    for so3w, l2w in ((1, 1), (5, 2), (11, 1), (1, 13)):
        a = angle_axis(-1.0, AXIS) + [1, 2, 3]
        b = angle_axis(2.0, AXIS) + [1, 0, -1]
        k.append((f"se3_space_test.cpp:45-90 ({so3w},{l2w})", m.se3_space(so3w, l2w, F64), a, b, 3.0 / 2 * so3w + SQRT20 * l2w, 1e-9))
        k.append((f"se2_space_test.cpp:45-84 ({so3w},{l2w})", m.se2_space(so3w, l2w, F64), [2, 3, -1.0], [0, -1, 2.0],
                  3.0 * so3w + SQRT20 * l2w, None))
    return k


def interpolate_kats():
    """-> list of (name, space, a, b, t, checker(result) -> bool)"""
    k = []
    k.append(("lp_space_test.cpp:53-66", m.lp_space(3, 2, F64), [1, 2, 3], [1, 0, -1], 0.1,
              lambda c: c[0] == 1.0 and c[1] == 1.8 and c[2] == 2.6))
    so2 = m.so2_space(1, 1, F64)
    for a, b, t, e in (
        (1.0, 1.0, 0.0, 1.0), (1.0, 1.0, 3.0, 1.0), (1.0, 2.0, 0.5, 1.5), (1.0, 2.0, -3.0, -2.0),
        (-1.0, 2.0, 4.0, 11.0 - 4 * PI), (5 * PI / 6, -5 * PI / 6, 1.0, -5 * PI / 6),
        (5 * PI / 6, -5 * PI / 6, 2.0, -3 * PI / 6), (-5 * PI / 6, 5 * PI / 6, 2.0, 3 * PI / 6),
    ):
        k.append(("so2_space_test.cpp:67-84", so2, [a], [b], t, (lambda e: lambda c: c[0] == e)(e)))

    def slerp_ok(c):
        ang, ax = to_angle_axis(c[:4])
        return sum((x - y) ** 2 for x, y in zip(ax, AXIS)) < 1e-15 and abs(ang - (0.5 + 2 * 0.1)) < 1e-10

    k.append(("so3_space_test.cpp:58-79", m.so3_space(F64), angle_axis(0.5, AXIS), angle_axis(2.5, AXIS), 0.1, slerp_ok))
    k.append(("se3_space_test.cpp:92-118", m.se3_space(1, 1, F64), angle_axis(0.5, AXIS) + [1, 2, 3],
              angle_axis(2.5, AXIS) + [1, 0, -1], 0.1,
              lambda c: slerp_ok(c) and c[4] == 1.0 and c[5] == 1.8 and c[6] == 2.6))
    return k


def run_distance_kats(distance_fn):
    """distance_fn(space, a[1,D], b[1,D]) -> array[1].  Returns list of failure strings."""
    bad = []
    for name, sp, a, b, exp, tol in distance_kats():
        got = float(distance_fn(sp, np.asarray([a], dtype=np.float64), np.asarray([b], dtype=np.float64))[0])
        ok = got == exp if tol is None else abs(got - exp) < tol
        if not ok:
            bad.append(f"{name}: got {got!r}, expected {exp!r}")
    return bad


def run_interpolate_kats(interp_fn):
    bad = []
    for name, sp, a, b, t, chk in interpolate_kats():
        got = interp_fn(sp, np.asarray([a], dtype=np.float64), np.asarray([b], dtype=np.float64), t)[0]
        if not chk([float(x) for x in got]):
            bad.append(f"{name}: got {got!r}")
    return bad


NO_INDEX = 0xFFFFFFFF


def replay_prrt(oracle, og, sp, lo, hi, start, goal, goal_radius, goal_bias, rng, seed, waves, W):
    """Worker::addSample (src/mpt/impl/prrt/prrt.hpp:411-452) on the oracle, one wave at a time, on the same samples."""
    nodes = [np.asarray(start, dtype=sp.dtype).reshape(1, -1)]
    parents = [np.array([NO_INDEX], dtype=np.uint32)]
    goal_node, drawn = NO_INDEX, 0
    for _ in range(waves):
        tree = np.concatenate(nodes)
        biased = goal is not None and goal_bias > 0 and goal_node == NO_INDEX
        smp = oracle.sample(sp, lo, hi, seed, drawn, W, goal if biased else None, goal_bias)
        drawn += W
        idx, dist, cnt = oracle.knn(sp, tree, smp, 1)
        near, d = tree[idx[:, 0]], dist[:, 0]
        to = oracle.steer(sp, near, smp, d, rng)
        keep = (cnt > 0) & (d != 0) & (og.valid(to) != 0) & (og.link(near, to) != 0)
        fresh = to[keep]
        if goal is not None and goal_node == NO_INDEX and len(fresh):
            hit = np.nonzero(oracle.distance(sp, fresh, np.broadcast_to(np.asarray(goal, dtype=sp.dtype), fresh.shape)) <= sp.dtype(goal_radius))[0]
            if hit.size:
                goal_node = tree.shape[0] + int(hit[0])
        nodes.append(fresh)
        parents.append(idx[keep, 0].astype(np.uint32))
    return np.concatenate(nodes), np.concatenate(parents), goal_node


# The reference's own PRRT (oracle/ref_planner.cpp) vs the planner-loop restatement above vs the device-resident PRRT:
# (range, samples, goal radius, goal bias, seed) on prrt_scene(); results in tests/golden/reference_golden.npz.
PRRT_CASES = [(25.0, 3000, 12.0, 0.05, 99), (float("inf"), 800, 12.0, 0.05, 99), (40.0, 2000, 1e-6, 0.0, 7), (60.0, 2500, 5.0, 0.2, 2026)]


def prrt_scene():
    from mpt_b200 import workloads as W

    occ = W.synthetic_grid(500, 400, seed=2)
    free = np.argwhere(occ == 0)
    start = free[len(free) // 7][::-1].astype(np.float64)
    goal = free[-len(free) // 9][::-1].astype(np.float64)
    return occ, [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1], start, goal


def pprm_k(space, n):
    """k = ceil(kRRG * ln(n + 1)), kRRG = e + e / dimensions, in the space's scalar type (src/mpt/impl/pprm/pprm.hpp:146,302-303)."""
    dt = space.dtype
    e = dt(math.e)
    return max(1, int(np.ceil((e + e / dt(space.dimensions)) * dt(math.log(n + 1.0)))))


def spanner_keeps(adj, kept, v, target, state):
    """ShortestPathCheck::operator() (src/mpt/impl/pprm_irs/shortest_path_check.hpp:111-225) for the node being added: True
    when no path of sparse edges to `v` is shorter than `target`.  adj: node -> [(neighbour, length)] of the roadmap before
    the wave; state = (cost dict, heap list) of the search, resumed from check to check; kept edges enter it at once."""
    import heapq

    cost, heap = state
    if v in cost and cost[v] < target:
        return False
    while heap:
        priority, top = heap[0]
        path_cost = cost[top]
        if path_cost >= target:
            break
        heapq.heappop(heap)
        if path_cost != priority:
            continue
        found = top == v
        for nbr, length in adj.get(top, ()):
            c = path_cost + length  # numpy scalar of the space's type: rounded as the device rounds it
            if nbr == v and c < target:
                found = True
            if nbr in cost and not (c < cost[nbr]):
                continue
            cost[nbr] = c
            heapq.heappush(heap, (c, nbr))
        if found:
            return False
    return True


def replay_pprm(oracle, og, sp, lo, hi, starts, goals, goal, goal_radius, seed, waves, W, stride, spanner_stretch=None):
    """PPRM's Worker::addSample (src/mpt/impl/pprm/pprm.hpp:298-339) on the oracle, one wave at a time, on the same
    samples: -> (states, edge rows [n, stride] of neighbour indices, edge distances, marks).  spanner_stretch: PPRM-IRS
    (impl/pprm_irs/pprm_irs.hpp:300-368) -- a validated edge is recorded only if the spanner needs it, every sample of a wave
    judged against the sparse roadmap as it stood when the wave began plus its own kept edges (mptg_pprm_set_spanner)."""
    D = sp.scalars
    states = np.empty((0, D), dtype=sp.dtype)
    rows_i, rows_d, marks = [], [], []

    def process(samples, forced):
        nonlocal states
        samples = np.asarray(samples, dtype=sp.dtype).reshape(-1, D)
        cand = samples[og.valid(samples) != 0]
        if cand.shape[0] == 0:
            return
        n = states.shape[0]
        if n > 0:
            k = min(pprm_k(sp, n), stride)
            idx, dist, cnt = oracle.knn(sp, states, cand, k)
            keep = ~((cnt > 0) & (dist[:, 0] < np.finfo(sp.dtype).eps))
            cand, idx, dist, cnt = cand[keep], idx[keep], dist[keep], cnt[keep]
        adj = {}
        if spanner_stretch is not None and n > 0:  # the sparse roadmap before this wave, both directions
            for r, row in enumerate(rows_i):
                for t in np.nonzero(row != NO_INDEX)[0]:
                    adj.setdefault(r, []).append((int(row[t]), rows_d[r][t]))
                    adj.setdefault(int(row[t]), []).append((r, rows_d[r][t]))
        for s in range(cand.shape[0]):
            ri = np.full(stride, NO_INDEX, dtype=np.uint32)
            rd = np.zeros(stride, dtype=sp.dtype)
            if n > 0 and cnt[s] > 0:
                c = int(cnt[s])
                ok = og.link(np.repeat(cand[s:s + 1], c, axis=0), states[idx[s, :c]]) != 0
                if spanner_stretch is not None:
                    search = ({}, [])
                    for j in range(c):
                        if not ok[j]:
                            continue
                        v, d = int(idx[s, j]), dist[s, j]
                        if spanner_keeps(adj, None, v, sp.dtype(spanner_stretch) * d, search):
                            import heapq
                            search[0][v] = d
                            heapq.heappush(search[1], (d, v))
                        else:
                            ok[j] = False
                ri[:c][ok] = idx[s, :c][ok]
                rd[:c][ok] = dist[s, :c][ok]
            mk = forced
            if goal is not None and not (mk & 2):
                if oracle.distance(sp, cand[s:s + 1], np.asarray(goal, dtype=sp.dtype).reshape(1, D))[0] <= sp.dtype(goal_radius):
                    mk |= 2
            rows_i.append(ri), rows_d.append(rd), marks.append(mk)
        states = np.concatenate([states, cand])

    for q in starts:
        process([q], 1)
    for q in goals:
        process([q], 2)
    drawn = 0
    for _ in range(waves):
        process(oracle.sample(sp, lo, hi, seed, drawn, W), 0)
        drawn += W
    return states, np.stack(rows_i), np.stack(rows_d), np.asarray(marks, dtype=np.uint8)


def components(edge_idx):
    """Connected-component label (smallest node index) per node of the undirected union of the edge rows."""
    n = edge_idx.shape[0]
    comp = np.arange(n)

    def find(x):
        while comp[x] != x:
            comp[x] = comp[comp[x]]
            x = comp[x]
        return x

    for r, c in zip(*np.nonzero(edge_idx != NO_INDEX)):
        a, b = find(int(r)), find(int(edge_idx[r, c]))
        if a != b:
            comp[max(a, b)] = min(a, b)
    return np.array([find(i) for i in range(n)])


def prrtstar_k(space, rewire_factor, n):
    """k = ceil(kRRG ln(n + 1)), kRRG = rewireFactor e (1 + 1/d) (src/mpt/impl/rrg_rewire_neighbors.hpp:53-61)."""
    dt = space.dtype
    k_rrg = dt(rewire_factor) * dt(math.e) * (dt(1) + dt(1) / dt(space.dimensions))
    return max(1, int(np.ceil(k_rrg * np.log(dt(n + 1.0)))))


def replay_prrtstar(oracle, og, sp, lo, hi, start, goal, goal_radius, goal_bias, rng, rewire_factor, seed, waves, W, stride, r_rrg=None):
    """The wave-parallel PRRT* of include/mptg/mptg.h (mptg_prrtstar_*) restated on the oracle: Worker::addSample of
    src/mpt/impl/prrt_star/prrt_star.hpp:510-657 per sample against the tree at the start of the wave; rewiring offers
    evaluated on the costs before the wave's rewiring step, best valid offer per node, applied at once, decreases
    pushed down the new ancestor chains.  -> (states, parents, costs, best goal node, rewires applied)"""
    dt = sp.dtype
    D = sp.scalars
    states = np.asarray(start, dtype=dt).reshape(1, D)
    parent = [NO_INDEX]
    cost = [dt(0)]
    goals, best_goal, drawn, rewires = [], NO_INDEX, 0, 0
    for _ in range(waves):
        tree = states
        biased = goal is not None and goal_bias > 0 and best_goal == NO_INDEX
        smp = oracle.sample(sp, lo, hi, seed, drawn, W, goal if biased else None, goal_bias)
        drawn += W
        idx, dist, cnt = oracle.knn(sp, tree, smp, 1)
        near, d = tree[idx[:, 0]], dist[:, 0]
        to, dre = oracle.steer(sp, near, smp, d, rng, with_distance=True)
        dnew = np.where(d > dt(rng), dre, d)
        keep = (cnt > 0) & (d != 0) & (og.valid(to) != 0) & (og.link(near, to) != 0)
        fresh, near_of, d_of = to[keep], idx[keep, 0], dnew[keep]
        S = fresh.shape[0]
        if S == 0:
            continue
        n0 = tree.shape[0]
        k = min(prrtstar_k(sp, rewire_factor, n0), stride)
        radius = -1.0
        if r_rrg is not None:  # rewire_r_nearest (src/mpt/impl/rrg_rewire_neighbors.hpp:102-128)
            n1 = dt(n0 + 1.0)
            radius = float(dt(r_rrg) * np.power(np.log(n1) / n1, dt(1) / dt(sp.dimensions)))
            k = stride
        nidx, ndist, ncnt = oracle.knn(sp, tree, fresh, k, radius)
        c0 = np.asarray(cost, dtype=dt)
        # candidate parents in cost + distance order up to the near node or the cut-off, one link batch
        orders, cands = [], []
        for s in range(S):
            c = c0[nidx[s, :ncnt[s]]] + ndist[s, :ncnt[s]]
            order = np.argsort(c, kind="stable")
            orders.append((order, c))
            parent_cost = c0[near_of[s]] + d_of[s]
            for j in order:
                if c[j] > parent_cost or nidx[s, j] == near_of[s]:
                    break
                cands.append((s, int(j)))
        ok = og.link(tree[[nidx[s, j] for s, j in cands]], fresh[[s for s, _ in cands]]) if cands else np.zeros(0, np.uint8)
        ok_of = {c: bool(o) for c, o in zip(cands, ok)}
        checked = np.zeros((S, k), dtype=bool)
        for s in range(S):
            order, c = orders[s]
            par, pc = int(near_of[s]), c0[near_of[s]] + d_of[s]
            parent_cost = pc
            for j in order:
                if c[j] > parent_cost:
                    break
                checked[s, j] = True
                if nidx[s, j] == near_of[s]:
                    pc = c[j]
                    break
                if ok_of[(s, int(j))]:
                    par, pc = int(nidx[s, j]), c[j]
                    break
            parent.append(par)
            cost.append(dt(pc))
            if goal is not None and oracle.distance(sp, fresh[s:s + 1], np.asarray(goal, dtype=dt).reshape(1, D))[0] <= dt(goal_radius):
                goals.append(n0 + s)
        states = np.concatenate([states, fresh])
        # rewiring offers on the costs as they are now
        c1 = np.asarray(cost, dtype=dt)
        offers = [(s, j) for s in range(S) for j in range(int(ncnt[s])) if not checked[s, j] and c1[n0 + s] + ndist[s, j] < c1[nidx[s, j]]]
        if offers:
            okr = og.link(fresh[[s for s, _ in offers]], tree[[nidx[s, j] for s, j in offers]])
            best = {}
            for e, ((s, j), o) in enumerate(zip(offers, okr)):
                if not o:
                    continue
                nb, new_cost = int(nidx[s, j]), c1[n0 + s] + ndist[s, j]
                key = (float(new_cost), e)
                if nb not in best or key < best[nb][0]:
                    best[nb] = (key, n0 + s, new_cost)
            delta = np.zeros(len(cost), dtype=dt)
            for nb, (_key, frm, new_cost) in best.items():
                parent[nb] = frm
                delta[nb] = c1[nb] - new_cost
            rewires += len(best)
            if best:
                for i in range(len(cost)):
                    c, a = c1[i], i
                    while a != NO_INDEX:
                        if delta[a] > 0:
                            c = dt(c - delta[a])
                        a = parent[a]
                    cost[i] = c
        if goals:
            cg = np.asarray(cost, dtype=dt)[goals]
            best_goal = goals[int(np.lexsort((goals, cg))[0])]
    return states, np.asarray(parent, dtype=np.uint32), np.asarray(cost, dtype=dt), best_goal, rewires
