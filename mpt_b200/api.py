"""Host-side mirror of the reference interfaces for the hot path, on top of the C ABI.

Names follow the reference (UNC-Robotics/mpt):
  Space / se3_space / lp_space ...      metric tags of src/mpt/impl/metrics.hpp, se3_space.hpp:91-113
  Nearest.insert / size / nearest       nigh::Nigh as used at impl/prrt/prrt.hpp:186,406-409,447 and
                                        impl/rrg_rewire_neighbors.hpp:65-67,125-128 (batched)
  Scenario.valid / link                 the Scenario concept (impl/prrt/prrt.hpp:439,454-457), batched
Host arrays are numpy (AoS, one state per row).  `*_dev` methods take raw device pointers (ints).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib as L


def _dtype(scalar: int):
    return np.float32 if scalar == L.F32 else np.float64


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Space:
    """A Cartesian product of weighted metric parts (mptg_space_desc)."""

    def __init__(self, parts: Sequence[Tuple], scalar: int = L.F32):
        """parts: sequence of (kind, p, dim, weight); kind in {'lp','so2','so3'}; p in {1,2,0(=inf)}."""
        self.desc = L.SpaceDesc()
        self.desc.n_parts = len(parts)
        self.desc.scalar = scalar
        kinds = {"lp": L.PART_LP, "so2": L.PART_SO2, "so3": L.PART_SO3}
        for i, (kind, p, dim, w) in enumerate(parts):
            self.desc.part[i].kind = kinds[kind]
            self.desc.part[i].p = p
            self.desc.part[i].dim = 4 if kind == "so3" else dim
            self.desc.part[i].weight = float(w)
        self.scalar = scalar
        self.dtype = _dtype(scalar)
        lib = L.load()
        self.scalars = lib.mptg_space_scalars(C.byref(self.desc))
        if self.scalars <= 0:
            raise ValueError("malformed space")
        self.dimensions = lib.mptg_space_dimensions(C.byref(self.desc))

    @property
    def ref(self):
        return C.byref(self.desc)


def se3_space(so3_weight: float = 1.0, l2_weight: float = 1.0, scalar: int = L.F32) -> Space:
    """SE3Space<Scalar, so3wt, l2wt>: rotation (x,y,z,w) first, then translation (se3_space.hpp:91-113)."""
    return Space([("so3", 0, 4, so3_weight), ("lp", 2, 3, l2_weight)], scalar)


def se2_space(so2_weight: float = 1.0, l2_weight: float = 1.0, scalar: int = L.F32) -> Space:
    """SE2Space: translation first, then the angle (se2_space.hpp:62-85)."""
    return Space([("lp", 2, 2, l2_weight), ("so2", 1, 1, so2_weight)], scalar)


def lp_space(dim: int, p: int = 2, scalar: int = L.F32, weight: float = 1.0) -> Space:
    return Space([("lp", p, dim, weight)], scalar)


def so2_space(dim: int = 1, p: int = 1, scalar: int = L.F32) -> Space:
    return Space([("so2", p, dim, 1.0)], scalar)


def so3_space(scalar: int = L.F32) -> Space:
    return Space([("so3", 0, 4, 1.0)], scalar)


class Context:
    """One GPU, one stream (mptg_ctx).  Single owner."""

    def __init__(self, device: int = -1):
        self.lib = L.load()
        h = C.c_void_p()
        L.check(self.lib.mptg_ctx_create(device, C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.mptg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        L.check(self.lib.mptg_sync(self.h), self.h)

    @property
    def stream(self) -> int:
        return self.lib.mptg_ctx_stream(self.h) or 0

    @property
    def launches(self) -> int:
        return int(self.lib.mptg_ctx_launch_count(self.h))

    @property
    def sm_count(self) -> int:
        return int(self.lib.mptg_ctx_sm_count(self.h))

    def probe_fp32_tflops(self) -> float:
        """FFMA rate of this GPU in TFLOP/s (mptg_probe_fp32_tflops)."""
        out = C.c_double(0.0)
        L.check(self.lib.mptg_probe_fp32_tflops(self.h, C.byref(out)), self.h)
        return float(out.value)

    # ---- metric helpers (a4, a5, steer)
    def _arr(self, space: Space, a, cols=None):
        a = np.ascontiguousarray(a, dtype=space.dtype)
        return a

    def distance(self, space: Space, a, b) -> np.ndarray:
        a = self._arr(space, a).reshape(-1, space.scalars)
        b = self._arr(space, b).reshape(-1, space.scalars)
        out = np.empty(a.shape[0], dtype=space.dtype)
        L.check(self.lib.mptg_distance_batch(self.h, space.ref, _ptr(a), _ptr(b), a.shape[0], _ptr(out)), self.h)
        return out

    def interpolate(self, space: Space, a, b, t) -> np.ndarray:
        a = self._arr(space, a).reshape(-1, space.scalars)
        b = self._arr(space, b).reshape(-1, space.scalars)
        t = np.ascontiguousarray(np.broadcast_to(np.asarray(t, dtype=space.dtype), (a.shape[0],)))
        out = np.empty_like(a)
        L.check(self.lib.mptg_interpolate_batch(self.h, space.ref, _ptr(a), _ptr(b), _ptr(t), a.shape[0], _ptr(out)), self.h)
        return out

    def steer(self, space: Space, near, sample, d, rng: float, with_distance: bool = False):
        near = self._arr(space, near).reshape(-1, space.scalars)
        sample = self._arr(space, sample).reshape(-1, space.scalars)
        d = np.ascontiguousarray(d, dtype=space.dtype)
        out = np.empty_like(near)
        dist = np.empty(near.shape[0], dtype=space.dtype) if with_distance else None
        L.check(self.lib.mptg_steer_batch(self.h, space.ref, _ptr(near), _ptr(sample), _ptr(d), near.shape[0], float(rng),
                                          _ptr(out), _ptr(dist)), self.h)
        return (out, dist) if with_distance else out


class Comm:
    """Communicator of the tree-sharded search (mptg_comm): one per process / GPU, NCCL underneath."""

    UNIQUE_ID_BYTES = 128

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(Comm.UNIQUE_ID_BYTES)
        L.check(L.load().mptg_comm_unique_id(buf))
        return buf.raw

    def __init__(self, ctx: Context, unique_id: bytes, rank: int, world: int):
        self.ctx = ctx
        h = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), Comm.UNIQUE_ID_BYTES)
        L.check(ctx.lib.mptg_comm_init(ctx.h, buf, rank, world, C.byref(h)), ctx.h)
        self.h, self.rank, self.world = h, rank, world

    def slice(self, n: int):
        first, count = C.c_uint32(), C.c_uint32()
        L.check(self.ctx.lib.mptg_comm_slice(self.h, n, C.byref(first), C.byref(count)), self.ctx.h)
        return first.value, count.value

    def sync(self, shard: "Nearest"):
        """Collective, after inserting: index the shard and exchange the shards' top-level boxes (mptg_knn_shard_sync)."""
        L.check(self.ctx.lib.mptg_knn_shard_sync(self.h, shard.h), self.ctx.h)

    def nearest(self, shard: "Nearest", queries, k: int = 1, radius: float = -1.0):
        """Collective: all ranks pass the same queries; returns this rank's slice (idx, dist, count) with global indices."""
        q = np.ascontiguousarray(queries, dtype=shard.space.dtype).reshape(-1, shard.space.scalars)
        _, n = self.slice(q.shape[0])
        idx = np.empty((n, k), dtype=np.uint32)
        dist = np.empty((n, k), dtype=shard.space.dtype)
        cnt = np.empty(n, dtype=np.uint32)
        r = float(radius) if radius is not None and math.isfinite(radius) else -1.0
        L.check(self.ctx.lib.mptg_knn_query_sharded(self.h, shard.h, _ptr(q), q.shape[0], k, r, _ptr(idx), _ptr(dist), _ptr(cnt)), self.ctx.h)
        return idx, dist, cnt

    def nearest_host_into(self, shard: "Nearest", q_ptr: int, Q: int, k: int, radius: float, idx_ptr: int, dist_ptr: int, cnt_ptr: int = 0):
        """host pointers in and out (this rank's slice of the results), no numpy allocation: bench.py's e2e path"""
        L.check(self.ctx.lib.mptg_knn_query_sharded(self.h, shard.h, C.c_void_p(q_ptr), Q, k, float(radius), C.c_void_p(idx_ptr),
                                                    C.c_void_p(dist_ptr), C.c_void_p(cnt_ptr) if cnt_ptr else None), self.ctx.h)

    def nearest_dev(self, shard: "Nearest", q_ptr: int, Q: int, k: int, radius: float, idx_ptr: int, dist_ptr: int, cnt_ptr: int = 0):
        L.check(self.ctx.lib.mptg_knn_query_sharded_dev(self.h, shard.h, C.c_void_p(q_ptr), Q, k, float(radius), C.c_void_p(idx_ptr),
                                                        C.c_void_p(dist_ptr), C.c_void_p(cnt_ptr) if cnt_ptr else None), self.ctx.h)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.mptg_comm_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Nearest:
    """Device-resident batched nearest-neighbour structure (mptg_knn)."""

    def __init__(self, ctx: Context, space: Space, capacity: int, strategy: int = L.KNN_AUTO):
        self.ctx, self.space = ctx, space
        h = C.c_void_p()
        L.check(ctx.lib.mptg_knn_create(ctx.h, space.ref, capacity, C.byref(h)), ctx.h)
        self.h = h
        if strategy != L.KNN_AUTO:
            self.set_strategy(strategy)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.mptg_knn_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_strategy(self, strategy: int):
        L.check(self.ctx.lib.mptg_knn_set_strategy(self.h, strategy), self.ctx.h)

    def set_index_map(self, mul: int, add: int):
        L.check(self.ctx.lib.mptg_knn_set_index_map(self.h, mul, add), self.ctx.h)

    def size(self) -> int:
        return int(self.ctx.lib.mptg_knn_size(self.h))

    def insert(self, states) -> int:
        s = np.ascontiguousarray(states, dtype=self.space.dtype).reshape(-1, self.space.scalars)
        first = C.c_uint32()
        L.check(self.ctx.lib.mptg_knn_insert(self.h, _ptr(s), s.shape[0], C.byref(first)), self.ctx.h)
        return first.value

    def insert_ids(self, states, ids):
        """Shard insert: results report ids[i] for states[i] (mptg_knn_insert_ids)."""
        s = np.ascontiguousarray(states, dtype=self.space.dtype).reshape(-1, self.space.scalars)
        i = np.ascontiguousarray(ids, dtype=np.uint32)
        assert i.shape[0] == s.shape[0]
        L.check(self.ctx.lib.mptg_knn_insert_ids(self.h, _ptr(s), _ptr(i), s.shape[0]), self.ctx.h)

    def insert_dev(self, ptr: int, count: int) -> int:
        first = C.c_uint32()
        L.check(self.ctx.lib.mptg_knn_insert_dev(self.h, C.c_void_p(ptr), count, C.byref(first)), self.ctx.h)
        return first.value

    def states(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.size() - first if count is None else count
        out = np.empty((count, self.space.scalars), dtype=self.space.dtype)
        L.check(self.ctx.lib.mptg_knn_get_states(self.h, first, count, _ptr(out)), self.ctx.h)
        return out

    def build_index(self):
        L.check(self.ctx.lib.mptg_knn_build_index(self.h), self.ctx.h)

    def nearest(self, queries, k: int = 1, radius: float = -1.0):
        """-> (idx [Q,k] uint32, dist [Q,k], count [Q] uint32), rows ascending by (distance, index)."""
        q = np.ascontiguousarray(queries, dtype=self.space.dtype).reshape(-1, self.space.scalars)
        Q = q.shape[0]
        idx = np.empty((Q, k), dtype=np.uint32)
        dist = np.empty((Q, k), dtype=self.space.dtype)
        cnt = np.empty(Q, dtype=np.uint32)
        r = float(radius) if radius is not None and math.isfinite(radius) else -1.0
        L.check(self.ctx.lib.mptg_knn_query(self.h, _ptr(q), Q, k, r, _ptr(idx), _ptr(dist), _ptr(cnt)), self.ctx.h)
        return idx, dist, cnt

    def nearest_host_into(self, q_ptr: int, Q: int, k: int, radius: float, idx_ptr: int, dist_ptr: int, cnt_ptr: int):
        """Raw host-pointer form (pinned buffers): the reference-facing C-ABI call timed by bench.py's e2e."""
        L.check(self.ctx.lib.mptg_knn_query(self.h, C.c_void_p(q_ptr), Q, k, float(radius), C.c_void_p(idx_ptr),
                                            C.c_void_p(dist_ptr), C.c_void_p(cnt_ptr) if cnt_ptr else None), self.ctx.h)

    def nearest_dev(self, q_ptr: int, Q: int, k: int, radius: float, idx_ptr: int, dist_ptr: int, cnt_ptr: int = 0):
        L.check(self.ctx.lib.mptg_knn_query_dev(self.h, C.c_void_p(q_ptr), Q, k, float(radius), C.c_void_p(idx_ptr),
                                                C.c_void_p(dist_ptr), C.c_void_p(cnt_ptr) if cnt_ptr else None), self.ctx.h)

    def last_stats(self):
        out = (C.c_uint64 * 4)()
        L.check(self.ctx.lib.mptg_knn_last_stats(self.h, out), self.ctx.h)
        return {"distance_evals": out[0], "nodes_visited": out[1], "indexed": out[2], "strategy": out[3]}


def _bounds(space: Space, lo, hi):
    lo = np.ascontiguousarray(np.broadcast_to(np.asarray(lo, dtype=np.float64), (space.scalars,)))
    hi = np.ascontiguousarray(np.broadcast_to(np.asarray(hi, dtype=np.float64), (space.scalars,)))
    return lo, hi


def sample(ctx: Context, space: Space, lo, hi, seed: int, first: int, n: int) -> np.ndarray:
    """n uniform samples of `space` (sample numbers first .. first+n-1 of the stream `seed`); lo / hi: one bound per
    scalar of the state (ignored for SO2 / SO3 parts).  UniformSampler of the reference, counter-based generator."""
    lo, hi = _bounds(space, lo, hi)
    out = np.empty((n, space.scalars), dtype=space.dtype)
    L.check(ctx.lib.mptg_sample_batch(ctx.h, space.ref, _ptr(lo), _ptr(hi), seed, first, n, _ptr(out)), ctx.h)
    return out


def sample_from_uniforms(ctx: Context, space: Space, lo, hi, uniforms) -> np.ndarray:
    """The deterministic half of the sampler: rows of uniforms in [0,1) -> states."""
    lo, hi = _bounds(space, lo, hi)
    per = ctx.lib.mptg_space_uniforms(space.ref)
    u = np.ascontiguousarray(uniforms, dtype=space.dtype).reshape(-1, per)
    out = np.empty((u.shape[0], space.scalars), dtype=space.dtype)
    L.check(ctx.lib.mptg_sample_transform_batch(ctx.h, space.ref, _ptr(lo), _ptr(hi), _ptr(u), u.shape[0], _ptr(out)), ctx.h)
    return out


class DevicePRRT:
    """Device-resident PRRT (mptg_prrt_*): Planner<Scenario, PRRT> with the tree kept on the GPU."""

    def __init__(self, scenario: "Scenario", space: Space, lo, hi, *, range: float = float("inf"), goal=None, goal_radius: float = 0.0,
                 goal_bias: float = 0.01, seed: int = 1, capacity: int = 1 << 20, max_wave: int = 1 << 16):
        self.ctx, self.scenario, self.space = scenario.ctx, scenario, space
        self._lo, self._hi = _bounds(space, lo, hi)
        self._goal = None if goal is None else np.ascontiguousarray(goal, dtype=space.dtype).reshape(space.scalars)
        prm = L.PrrtParams(C.pointer(space.desc), self._lo.ctypes.data, self._hi.ctypes.data, min(float(range), 1.7e308), float(goal_bias),
                           None if self._goal is None else self._goal.ctypes.data, float(goal_radius), float(scenario.step or 0.0),
                           int(seed), int(capacity), int(max_wave))
        self.h = C.c_void_p()
        L.check(self.ctx.lib.mptg_prrt_create(self.ctx.h, scenario.h, C.byref(prm), C.byref(self.h)), self.ctx.h)
        self.goal_node = L.NO_INDEX

    def add_start(self, state):
        s = np.ascontiguousarray(state, dtype=self.space.dtype).reshape(self.space.scalars)
        L.check(self.ctx.lib.mptg_prrt_add_start(self.h, _ptr(s)), self.ctx.h)

    def wave(self, n_samples: int) -> int:
        size, goal = C.c_uint32(), C.c_uint32()
        L.check(self.ctx.lib.mptg_prrt_wave(self.h, n_samples, C.byref(size), C.byref(goal)), self.ctx.h)
        self.goal_node = goal.value
        return size.value

    @property
    def size(self) -> int:
        return self.ctx.lib.mptg_prrt_size(self.h)

    @property
    def samples_drawn(self) -> int:
        return self.ctx.lib.mptg_prrt_samples_drawn(self.h)

    def solved(self) -> bool:
        return self.goal_node != L.NO_INDEX

    def tree(self, first: int = 0, count: int | None = None):
        """-> (states [n, D], parents [n] uint32; NO_INDEX for a start node)"""
        count = self.size - first if count is None else count
        st = np.empty((count, self.space.scalars), dtype=self.space.dtype)
        pa = np.empty(count, dtype=np.uint32)
        L.check(self.ctx.lib.mptg_prrt_get_tree(self.h, first, count, _ptr(st), _ptr(pa)), self.ctx.h)
        return st, pa

    def solution(self) -> np.ndarray:
        """Planner::solution(): states from the start to the first goal node (prrt.hpp:232-249)."""
        if not self.solved():
            return np.empty((0, self.space.scalars), dtype=self.space.dtype)
        st, pa = self.tree()
        path, n = [], self.goal_node
        while n != L.NO_INDEX:
            path.append(st[n])
            n = int(pa[n])
        return np.stack(path[::-1])

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.mptg_prrt_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DevicePRRTStar(DevicePRRT):
    """Device-resident PRRT* (mptg_prrtstar_*): Planner<Scenario, PRRTStar> with tree, costs and rewiring on the GPU."""

    def __init__(self, scenario: "Scenario", space: Space, lo, hi, *, range: float = float("inf"), goal=None, goal_radius: float = 0.0,
                 goal_bias: float = 0.01, rewire_factor: float = 1.1, rewire_radius: float | None = None, seed: int = 1, capacity: int = 1 << 20,
                 max_wave: int = 1 << 14):
        """rewire_radius: r_rrg of rewire_r_nearest (mptg_prrtstar_set_rewire_radius); None = k-nearest rewiring."""
        self.ctx, self.scenario, self.space = scenario.ctx, scenario, space
        self._lo, self._hi = _bounds(space, lo, hi)
        self._goal = None if goal is None else np.ascontiguousarray(goal, dtype=space.dtype).reshape(space.scalars)
        prm = L.PrrtParams(C.pointer(space.desc), self._lo.ctypes.data, self._hi.ctypes.data, min(float(range), 1.7e308), float(goal_bias),
                           None if self._goal is None else self._goal.ctypes.data, float(goal_radius), float(scenario.step or 0.0),
                           int(seed), int(capacity), int(max_wave))
        self.h = C.c_void_p()
        L.check(self.ctx.lib.mptg_prrtstar_create(self.ctx.h, scenario.h, C.byref(prm), float(rewire_factor), C.byref(self.h)), self.ctx.h)
        if rewire_radius is not None:
            L.check(self.ctx.lib.mptg_prrtstar_set_rewire_radius(self.h, float(rewire_radius)), self.ctx.h)
        self.goal_node = L.NO_INDEX

    def add_start(self, state):
        s = np.ascontiguousarray(state, dtype=self.space.dtype).reshape(self.space.scalars)
        L.check(self.ctx.lib.mptg_prrtstar_add_start(self.h, _ptr(s)), self.ctx.h)

    def wave(self, n_samples: int) -> int:
        size, goal = C.c_uint32(), C.c_uint32()
        L.check(self.ctx.lib.mptg_prrtstar_wave(self.h, n_samples, C.byref(size), C.byref(goal)), self.ctx.h)
        self.goal_node = goal.value
        return size.value

    @property
    def size(self) -> int:
        return self.ctx.lib.mptg_prrtstar_size(self.h)

    @property
    def samples_drawn(self) -> int:
        return self.ctx.lib.mptg_prrtstar_samples_drawn(self.h)

    @property
    def rewires(self) -> int:
        return self.ctx.lib.mptg_prrtstar_rewires(self.h)

    def tree(self, first: int = 0, count: int | None = None, with_costs: bool = False):
        """-> (states [n, D], parents [n] uint32[, costs [n]]); the goal node of smallest cost is `goal_node`"""
        count = self.size - first if count is None else count
        st = np.empty((count, self.space.scalars), dtype=self.space.dtype)
        pa = np.empty(count, dtype=np.uint32)
        co = np.empty(count, dtype=self.space.dtype)
        L.check(self.ctx.lib.mptg_prrtstar_get_tree(self.h, first, count, _ptr(st), _ptr(pa), _ptr(co)), self.ctx.h)
        return (st, pa, co) if with_costs else (st, pa)

    def solution_cost(self) -> float:
        return float("nan") if not self.solved() else float(self.tree(self.goal_node, 1, with_costs=True)[2][0])

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.mptg_prrtstar_destroy(self.h)
        self.h = None


class DevicePPRM:
    """Device-resident PPRM (mptg_pprm_*): Planner<Scenario, PPRM> with the roadmap kept on the GPU."""

    START, GOAL = 1, 2

    def __init__(self, scenario: "Scenario", space: Space, lo, hi, *, goal=None, goal_radius: float = 0.0, seed: int = 1,
                 capacity: int = 1 << 20, max_wave: int = 1 << 14, max_k: int = 0, spanner_stretch: float = 0.0, spanner_capacity: int = 0):
        """spanner_stretch > 0: PPRM-IRS (mptg_pprm_set_spanner) -- only the sparse edges of the roadmap spanner are kept."""
        self.ctx, self.scenario, self.space = scenario.ctx, scenario, space
        self._lo, self._hi = _bounds(space, lo, hi)
        self._goal = None if goal is None else np.ascontiguousarray(goal, dtype=space.dtype).reshape(space.scalars)
        prm = L.PprmParams(C.pointer(space.desc), self._lo.ctypes.data, self._hi.ctypes.data, None if self._goal is None else self._goal.ctypes.data,
                           float(goal_radius), float(scenario.step or 0.0), int(seed), int(capacity), int(max_wave), int(max_k))
        self.h = C.c_void_p()
        L.check(self.ctx.lib.mptg_pprm_create(self.ctx.h, scenario.h, C.byref(prm), C.byref(self.h)), self.ctx.h)
        if spanner_stretch > 0:
            L.check(self.ctx.lib.mptg_pprm_set_spanner(self.h, float(spanner_stretch), int(spanner_capacity)), self.ctx.h)
        self._solved = False

    def add_state(self, state, marks: int) -> int:
        """addStart / addGoal (pprm.hpp:156-169); returns the node or NO_INDEX when the state was rejected."""
        s = np.ascontiguousarray(state, dtype=self.space.dtype).reshape(self.space.scalars)
        node = C.c_uint32()
        L.check(self.ctx.lib.mptg_pprm_add_state(self.h, _ptr(s), marks, C.byref(node)), self.ctx.h)
        return node.value

    def add_start(self, state) -> int:
        return self.add_state(state, self.START)

    def add_goal(self, state) -> int:
        return self.add_state(state, self.GOAL)

    def wave(self, n_samples: int) -> int:
        size, solved = C.c_uint32(), C.c_uint32()
        L.check(self.ctx.lib.mptg_pprm_wave(self.h, n_samples, C.byref(size), C.byref(solved)), self.ctx.h)
        self._solved = bool(solved.value)
        return size.value

    @property
    def size(self) -> int:
        return self.ctx.lib.mptg_pprm_size(self.h)

    @property
    def samples_drawn(self) -> int:
        return self.ctx.lib.mptg_pprm_samples_drawn(self.h)

    @property
    def row_stride(self) -> int:
        return self.ctx.lib.mptg_pprm_row_stride(self.h)

    def solved(self) -> bool:
        return self._solved

    def graph(self, first: int = 0, count: int | None = None):
        """-> (states [n, D], edge_idx [n, stride] (NO_INDEX = unused), edge_dist [n, stride], marks [n], component [n])"""
        count = self.size - first if count is None else count
        K = self.row_stride
        st = np.empty((count, self.space.scalars), dtype=self.space.dtype)
        ei = np.empty((count, K), dtype=np.uint32)
        ed = np.empty((count, K), dtype=self.space.dtype)
        mk = np.empty(count, dtype=np.uint8)
        cp = np.empty(count, dtype=np.uint32)
        L.check(self.ctx.lib.mptg_pprm_get_graph(self.h, first, count, _ptr(st), _ptr(ei), _ptr(ed), _ptr(mk), _ptr(cp)), self.ctx.h)
        return st, ei, ed, mk, cp

    def solution(self) -> np.ndarray:
        """Planner::solution(): shortest roadmap path from a start to a goal (Dijkstra, pprm.hpp:218-246)."""
        import heapq

        st, ei, ed, mk, _ = self.graph()
        n = st.shape[0]
        adj = [[] for _ in range(n)]
        rows, cols = np.nonzero(ei != L.NO_INDEX)
        for r, c in zip(rows.tolist(), cols.tolist()):
            nb, d = int(ei[r, c]), float(ed[r, c])
            adj[r].append((nb, d))
            adj[nb].append((r, d))
        dist, prev = np.full(n, np.inf), np.full(n, -1, dtype=np.int64)
        pq = []
        for s in np.nonzero(mk & self.START)[0].tolist():
            dist[s] = 0.0
            heapq.heappush(pq, (0.0, s))
        hit = -1
        while pq:
            d, u = heapq.heappop(pq)
            if d > dist[u]:
                continue
            if mk[u] & self.GOAL:
                hit = u
                break
            for v, w in adj[u]:
                if d + w < dist[v]:
                    dist[v], prev[v] = d + w, u
                    heapq.heappush(pq, (d + w, v))
        path = []
        while hit >= 0:
            path.append(st[hit])
            hit = int(prev[hit])
        return np.stack(path[::-1]) if path else np.empty((0, self.space.scalars), dtype=self.space.dtype)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.mptg_pprm_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def knn_merge_dev(ctx: Context, scalar: int, parts: int, Q: int, k: int, idx_in: int, dist_in: int, idx_out: int,
                  dist_out: int, cnt_out: int = 0):
    L.check(ctx.lib.mptg_knn_merge_dev(ctx.h, scalar, parts, Q, k, C.c_void_p(idx_in), C.c_void_p(dist_in),
                                       C.c_void_p(idx_out), C.c_void_p(dist_out), C.c_void_p(cnt_out) if cnt_out else None), ctx.h)


class Scenario:
    """Batched scenario.valid(q) / scenario.link(a, b) over a registered geometry (mptg_geom)."""

    def __init__(self, ctx: Context, handle, scalar: int, state_scalars: int, space: Optional[Space] = None,
                 step: float = 0.0):
        self.ctx, self.h, self.scalar, self.D = ctx, handle, scalar, state_scalars
        self.dtype = _dtype(scalar)
        self.space, self.step = space, step

    # ---- factories (one per reference scenario)
    @classmethod
    def grid(cls, ctx: Context, occupancy: np.ndarray, scalar: int = L.F64):
        """PNG2dScenario (demo/png_2d_scenario.hpp): occupancy[h, w] non-zero = obstacle."""
        occ = np.ascontiguousarray(occupancy, dtype=np.uint8)
        hgt, wid = occ.shape
        h = C.c_void_p()
        L.check(ctx.lib.mptg_grid_create(ctx.h, scalar, wid, hgt, _ptr(occ), C.byref(h)), ctx.h)
        return cls(ctx, h, scalar, 2)

    @classmethod
    def shapes(cls, ctx: Context, dim: int, centres, radii, rects=(), scalar: int = L.F64):
        """Holonomic2DPointScenario / sphere test scenario: balls + (2-D) rectangles."""
        c = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, dim) if len(radii) else np.zeros((0, dim))
        r = np.ascontiguousarray(radii, dtype=np.float64)
        rc = np.ascontiguousarray(rects, dtype=np.float64).reshape(-1, 4) if len(rects) else np.zeros((0, 4))
        h = C.c_void_p()
        L.check(ctx.lib.mptg_shapes_create(ctx.h, scalar, dim, r.shape[0], _ptr(c), _ptr(r), rc.shape[0], _ptr(rc), C.byref(h)), ctx.h)
        return cls(ctx, h, scalar, dim)

    @classmethod
    def link_arm(cls, ctx: Context, lengths, link_radius: float, circles, scalar: int = L.F64):
        """LinkManipulatorScenario (demo/link_manipulator_scenario.hpp)."""
        ln = np.ascontiguousarray(lengths, dtype=np.float64)
        cc = np.ascontiguousarray(circles, dtype=np.float64).reshape(-1, 3)
        h = C.c_void_p()
        L.check(ctx.lib.mptg_linkarm_create(ctx.h, scalar, ln.shape[0], _ptr(ln), float(link_radius), cc.shape[0], _ptr(cc), C.byref(h)), ctx.h)
        return cls(ctx, h, scalar, ln.shape[0])

    @classmethod
    def nao_cup(cls, ctx: Context, scalar: int = L.F64):
        """NaoCupScenario (demo/nao_cup_planning.cpp:50-153): 10 joint angles, the reference's robot, cup and obstacles."""
        h = C.c_void_p()
        L.check(ctx.lib.mptg_naocup_create(ctx.h, scalar, C.byref(h)), ctx.h)
        return cls(ctx, h, scalar, 10)

    @staticmethod
    def nao_cup_configs(scalar: int = L.F64):
        """(start, goal, lo, hi) of the reference (naocup.hpp:254-301) as float64 arrays of 10."""
        out = [np.zeros(10) for _ in range(4)]
        L.check(L.load().mptg_naocup_configs(scalar, *[_ptr(a) for a in out]), None)
        return tuple(out)

    @classmethod
    def mesh_pair(cls, ctx: Context, robot_tris, env_tris, space: Space, step: float):
        """SE3RigidBodyScenario (demo/se3_rigid_body_scenario.hpp): triangle soups [n,3,3] float32."""
        rt = np.ascontiguousarray(robot_tris, dtype=np.float32).reshape(-1, 9)
        et = np.ascontiguousarray(env_tris, dtype=np.float32).reshape(-1, 9)
        h = C.c_void_p()
        L.check(ctx.lib.mptg_mesh_pair_create(ctx.h, space.scalar, rt.shape[0], _ptr(rt), et.shape[0], _ptr(et), C.byref(h)), ctx.h)
        return cls(ctx, h, space.scalar, 7, space, step)

    def close(self):
        if getattr(self, "h", None) and getattr(self.ctx, "h", None):
            self.ctx.lib.mptg_geom_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def valid(self, states, with_near_contact=False):
        """scenario.valid for a batch; with_near_contact also returns the near-contact flags (mptg.h)."""
        s = np.ascontiguousarray(states, dtype=self.dtype).reshape(-1, self.D)
        ok = np.empty(s.shape[0], dtype=np.uint8)
        near = np.empty(s.shape[0], dtype=np.uint8) if with_near_contact else None
        L.check(self.ctx.lib.mptg_valid_batch(self.h, _ptr(s), s.shape[0], _ptr(ok), _ptr(near)), self.ctx.h)
        return (ok, near) if with_near_contact else ok

    def link(self, a, b, with_near_contact=False):
        a = np.ascontiguousarray(a, dtype=self.dtype).reshape(-1, self.D)
        b = np.ascontiguousarray(b, dtype=self.dtype).reshape(-1, self.D)
        ok = np.empty(a.shape[0], dtype=np.uint8)
        near = np.empty(a.shape[0], dtype=np.uint8) if with_near_contact else None
        sp = self.space.ref if self.space is not None else None
        L.check(self.ctx.lib.mptg_link_batch(self.h, sp, _ptr(a), _ptr(b), a.shape[0], float(self.step), _ptr(ok), _ptr(near)), self.ctx.h)
        return (ok, near) if with_near_contact else ok

    def contact_band(self) -> float:
        out = C.c_double()
        L.check(self.ctx.lib.mptg_geom_contact_band(self.h, C.byref(out)), self.ctx.h)
        return out.value

    def link_host_into(self, a_ptr: int, b_ptr: int, n: int, ok_ptr: int):
        sp = self.space.ref if self.space is not None else None
        L.check(self.ctx.lib.mptg_link_batch(self.h, sp, C.c_void_p(a_ptr), C.c_void_p(b_ptr), n, float(self.step),
                                             C.c_void_p(ok_ptr), None), self.ctx.h)

    def valid_dev(self, s_ptr: int, n: int, ok_ptr: int):
        L.check(self.ctx.lib.mptg_valid_batch_dev(self.h, C.c_void_p(s_ptr), n, C.c_void_p(ok_ptr), None), self.ctx.h)

    def link_dev(self, a_ptr: int, b_ptr: int, n: int, ok_ptr: int):
        sp = self.space.ref if self.space is not None else None
        L.check(self.ctx.lib.mptg_link_batch_dev(self.h, sp, C.c_void_p(a_ptr), C.c_void_p(b_ptr), n, float(self.step),
                                                 C.c_void_p(ok_ptr), None), self.ctx.h)

    def last_stats(self):
        out = (C.c_uint64 * 4)()
        L.check(self.ctx.lib.mptg_geom_last_stats(self.h, out), self.ctx.h)
        return {"states": out[0], "bv_tests": out[1], "prim_tests": out[2], "items": out[3]}
