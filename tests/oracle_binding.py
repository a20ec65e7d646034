"""ctypes binding of oracle/_build/liboracle.so -- the CPU restatement of the reference path.

TEST INFRASTRUCTURE: used by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs only.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ODIR = ROOT / "oracle"
LIB = ODIR / "_build" / "liboracle.so"

F32, F64 = 4, 8
_P = C.c_void_p


def build(force: bool = False):
    if force or not LIB.exists() or any(
        p.stat().st_mtime > LIB.stat().st_mtime
        for p in list(ODIR.glob("*.cpp")) + list(ODIR.glob("*.hpp")) + list((ROOT / "include" / "mptg").glob("*.h"))
    ):
        subprocess.run(["make", "-C", str(ODIR), "-j4"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


class _Part(C.Structure):
    _fields_ = [("kind", C.c_int32), ("p", C.c_int32), ("dim", C.c_int32), ("_pad", C.c_int32), ("weight", C.c_double)]


class _Desc(C.Structure):  # mptg_space_desc (include/mptg/mptg.h), declared here so that the CPU arm never loads libmptg.so
    _fields_ = [("n_parts", C.c_int32), ("scalar", C.c_int32), ("part", _Part * 8)]


class OracleSpace:
    """A space descriptor for the oracle alone (bench.py --impl reference): same layout as mpt_b200.Space, built
    without touching the product library."""

    def __init__(self, parts, scalar=F32):
        kinds = {"lp": 1, "so2": 2, "so3": 3}
        self.desc = _Desc()
        self.desc.n_parts, self.desc.scalar = len(parts), scalar
        self.scalars = 0
        for i, (kind, p, dim, w) in enumerate(parts):
            self.desc.part[i].kind, self.desc.part[i].p = kinds[kind], p
            self.desc.part[i].dim = 4 if kind == "so3" else dim
            self.desc.part[i].weight = float(w)
            self.scalars += 4 if kind == "so3" else dim
        self.scalar = scalar
        self.dtype = np.float32 if scalar == F32 else np.float64

    @property
    def ref(self):
        return C.byref(self.desc)


def se3_space(so3_weight=1.0, l2_weight=1.0, scalar=F32):
    return OracleSpace([("so3", 0, 4, so3_weight), ("lp", 2, 3, l2_weight)], scalar)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.orc_num_threads.restype = C.c_int
        lib.orc_tree_create.restype = _P
        lib.orc_tree_create.argtypes = [_P, _P, C.c_uint32]
        lib.orc_tree_destroy.argtypes = [_P]
        lib.orc_tree_knn.argtypes = [_P, _P, C.c_uint32, C.c_uint32, C.c_double, _P, _P, _P, _P]
        for f in ("orc_grid_create", "orc_shapes_create", "orc_linkarm_create", "orc_mesh_pair_create", "orc_naocup_create"):
            getattr(lib, f).restype = _P
        lib.orc_grid_create.argtypes = [C.c_int, C.c_int, C.c_int, _P]
        lib.orc_shapes_create.argtypes = [C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P]
        lib.orc_linkarm_create.argtypes = [C.c_int, C.c_int, _P, C.c_double, C.c_int, _P]
        lib.orc_mesh_pair_create.argtypes = [C.c_int, C.c_uint32, _P, C.c_uint32, _P]
        lib.orc_naocup_create.argtypes = [C.c_int]
        lib.orc_geom_destroy.argtypes = [_P]
        lib.orc_valid_batch.argtypes = [_P, _P, C.c_uint32, _P, _P]
        lib.orc_link_batch.argtypes = [_P, _P, _P, _P, C.c_uint32, C.c_double, _P, _P, C.c_double, _P]
        lib.orc_geom_counters.argtypes = [_P, _P]
        lib.orc_distance_batch.argtypes = [_P, _P, _P, C.c_uint32, _P]
        lib.orc_interpolate_batch.argtypes = [_P, _P, _P, _P, C.c_uint32, _P]
        lib.orc_steer_batch.argtypes = [_P, _P, _P, _P, C.c_uint32, C.c_double, _P, _P]
        lib.orc_knn.argtypes = [_P, _P, C.c_uint32, _P, C.c_uint32, C.c_uint32, C.c_double, _P, _P, _P]
        lib.orc_acos01f.restype = C.c_double
        lib.orc_acos01f.argtypes = [C.c_float]
        lib.orc_acos01d.restype = C.c_double
        lib.orc_acos01d.argtypes = [C.c_double]

    @property
    def threads(self) -> int:
        return self.lib.orc_num_threads()

    def set_threads(self, n: int):
        self.lib.orc_set_num_threads(int(n))

    # space: mpt_b200.Space (only its ctypes desc is used -- same struct layout as include/mptg/mptg.h)
    def distance(self, space, a, b):
        a = np.ascontiguousarray(a, dtype=space.dtype).reshape(-1, space.scalars)
        b = np.ascontiguousarray(b, dtype=space.dtype).reshape(-1, space.scalars)
        out = np.empty(a.shape[0], dtype=space.dtype)
        self.lib.orc_distance_batch(space.ref, _ptr(a), _ptr(b), a.shape[0], _ptr(out))
        return out

    def interpolate(self, space, a, b, t):
        a = np.ascontiguousarray(a, dtype=space.dtype).reshape(-1, space.scalars)
        b = np.ascontiguousarray(b, dtype=space.dtype).reshape(-1, space.scalars)
        t = np.ascontiguousarray(np.broadcast_to(np.asarray(t, dtype=space.dtype), (a.shape[0],)))
        out = np.empty_like(a)
        self.lib.orc_interpolate_batch(space.ref, _ptr(a), _ptr(b), _ptr(t), a.shape[0], _ptr(out))
        return out

    def sample(self, space, lo, hi, seed, first, n, goal=None, goal_bias=0.0):
        lo = np.ascontiguousarray(np.broadcast_to(np.asarray(lo, dtype=np.float64), (space.scalars,)))
        hi = np.ascontiguousarray(np.broadcast_to(np.asarray(hi, dtype=np.float64), (space.scalars,)))
        g = None if goal is None else np.ascontiguousarray(goal, dtype=space.dtype).reshape(space.scalars)
        out = np.empty((n, space.scalars), dtype=space.dtype)
        self.lib.orc_sample_batch(space.ref, _ptr(lo), _ptr(hi), C.c_uint64(seed), C.c_uint64(first), C.c_uint32(n), _ptr(g), C.c_double(goal_bias), _ptr(out))
        return out

    def sample_uniforms(self, scalar, seed, first, n, per_state):
        """Raw uniforms of the sample streams: [n, per_state], column 0 = the goal-bias draw."""
        out = np.empty((n, per_state), dtype=np.float32 if scalar == 4 else np.float64)
        self.lib.orc_sample_uniforms(C.c_int(scalar), C.c_uint64(seed), C.c_uint64(first), C.c_uint32(n), C.c_int(per_state), _ptr(out))
        return out

    def sample_from_uniforms(self, space, lo, hi, uniforms, per_state):
        lo = np.ascontiguousarray(np.broadcast_to(np.asarray(lo, dtype=np.float64), (space.scalars,)))
        hi = np.ascontiguousarray(np.broadcast_to(np.asarray(hi, dtype=np.float64), (space.scalars,)))
        u = np.ascontiguousarray(uniforms, dtype=space.dtype).reshape(-1, per_state)
        out = np.empty((u.shape[0], space.scalars), dtype=space.dtype)
        self.lib.orc_sample_from_uniforms(space.ref, _ptr(lo), _ptr(hi), _ptr(u), C.c_int(per_state), C.c_uint32(u.shape[0]), _ptr(out))
        return out

    def steer(self, space, near, sample, d, rng, with_distance=False):
        near = np.ascontiguousarray(near, dtype=space.dtype).reshape(-1, space.scalars)
        sample = np.ascontiguousarray(sample, dtype=space.dtype).reshape(-1, space.scalars)
        d = np.ascontiguousarray(d, dtype=space.dtype)
        out = np.empty_like(near)
        dist = np.empty(near.shape[0], dtype=space.dtype) if with_distance else None
        self.lib.orc_steer_batch(space.ref, _ptr(near), _ptr(sample), _ptr(d), near.shape[0], float(rng), _ptr(out), _ptr(dist))
        return (out, dist) if with_distance else out

    def knn(self, space, pts, queries, k, radius=-1.0):
        pts = np.ascontiguousarray(pts, dtype=space.dtype).reshape(-1, space.scalars)
        q = np.ascontiguousarray(queries, dtype=space.dtype).reshape(-1, space.scalars)
        Q = q.shape[0]
        idx = np.empty((Q, k), dtype=np.uint32)
        dist = np.empty((Q, k), dtype=space.dtype)
        cnt = np.empty(Q, dtype=np.uint32)
        self.lib.orc_knn(space.ref, _ptr(pts), pts.shape[0], _ptr(q), Q, k, float(radius), _ptr(idx), _ptr(dist), _ptr(cnt))
        return idx, dist, cnt

    def tree(self, space, pts):
        return OracleTree(self, space, pts)

    def grid(self, occupancy, scalar=F64):
        occ = np.ascontiguousarray(occupancy, dtype=np.uint8)
        return OracleGeom(self, self.lib.orc_grid_create(scalar, occ.shape[1], occ.shape[0], _ptr(occ)), scalar, 2)

    def shapes(self, dim, centres, radii, rects=(), scalar=F64):
        c = np.ascontiguousarray(centres, dtype=np.float64).reshape(-1, dim) if len(radii) else np.zeros((0, dim))
        r = np.ascontiguousarray(radii, dtype=np.float64)
        rc = np.ascontiguousarray(rects, dtype=np.float64).reshape(-1, 4) if len(rects) else np.zeros((0, 4))
        return OracleGeom(self, self.lib.orc_shapes_create(scalar, dim, r.shape[0], _ptr(c), _ptr(r), rc.shape[0], _ptr(rc)), scalar, dim)

    def link_arm(self, lengths, link_radius, circles, scalar=F64):
        ln = np.ascontiguousarray(lengths, dtype=np.float64)
        cc = np.ascontiguousarray(circles, dtype=np.float64).reshape(-1, 3)
        return OracleGeom(self, self.lib.orc_linkarm_create(scalar, ln.shape[0], _ptr(ln), float(link_radius), cc.shape[0], _ptr(cc)), scalar, ln.shape[0])

    def nao_cup(self, scalar=F64):
        """oracle/oracle_nao.hpp: valid = nao_clear, link = nao_link.  valid(..., with_margin=True) also returns the smallest
        |distance - reach| / |cup axis - threshold| of the state; link(..., with_near_contact=True, tol_rel=t) flags the
        edges with such a margin below t on some midpoint of their recursion."""
        return OracleGeom(self, self.lib.orc_naocup_create(scalar), scalar, 10)

    def tri_pairs(self, P, Q):
        """n triangle pairs [n,3,3] (double): (SAT decision, SAT margin, orientation-predicate decision)."""
        P = np.ascontiguousarray(P, dtype=np.float64).reshape(-1, 9)
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(-1, 9)
        sat, pred = np.empty(P.shape[0], np.uint8), np.empty(P.shape[0], np.uint8)
        margin = np.empty(P.shape[0], np.float64)
        self.lib.orc_tri_pairs(C.c_uint32(P.shape[0]), _ptr(P), _ptr(Q), _ptr(sat), _ptr(margin), _ptr(pred))
        return sat, margin, pred

    def mesh_pair(self, robot_tris, env_tris, space, step):
        rt = np.ascontiguousarray(robot_tris, dtype=np.float32).reshape(-1, 9)
        et = np.ascontiguousarray(env_tris, dtype=np.float32).reshape(-1, 9)
        g = OracleGeom(self, self.lib.orc_mesh_pair_create(space.scalar, rt.shape[0], _ptr(rt), et.shape[0], _ptr(et)), space.scalar, 7)
        g.space, g.step = space, step
        return g


class OracleTree:
    def __init__(self, orc, space, pts):
        self.orc, self.space = orc, space
        pts = np.ascontiguousarray(pts, dtype=space.dtype).reshape(-1, space.scalars)
        self.h = orc.lib.orc_tree_create(space.ref, _ptr(pts), pts.shape[0])
        if not self.h:
            raise RuntimeError("oracle tree supports float32 spaces only")

    def knn(self, queries, k, radius=-1.0):
        q = np.ascontiguousarray(queries, dtype=self.space.dtype).reshape(-1, self.space.scalars)
        Q = q.shape[0]
        idx = np.empty((Q, k), dtype=np.uint32)
        dist = np.empty((Q, k), dtype=self.space.dtype)
        cnt = np.empty(Q, dtype=np.uint32)
        ev = C.c_uint64()
        self.orc.lib.orc_tree_knn(self.h, _ptr(q), Q, k, float(radius), _ptr(idx), _ptr(dist), _ptr(cnt), C.byref(ev))
        self.last_evals = ev.value
        return idx, dist, cnt

    def __del__(self):
        try:
            self.orc.lib.orc_tree_destroy(self.h)
        except Exception:
            pass


class OracleGeom:
    def __init__(self, orc, h, scalar, D):
        self.orc, self.h, self.scalar, self.D = orc, h, scalar, D
        self.dtype = np.float32 if scalar == F32 else np.float64
        self.space, self.step = None, 0.0

    def valid(self, states, with_margin=False):
        s = np.ascontiguousarray(states, dtype=self.dtype).reshape(-1, self.D)
        ok = np.empty(s.shape[0], dtype=np.uint8)
        margin = np.empty(s.shape[0], dtype=np.float64) if with_margin else None
        self.orc.lib.orc_valid_batch(self.h, _ptr(s), s.shape[0], _ptr(ok), _ptr(margin))
        return (ok, margin) if with_margin else ok

    def link(self, a, b, with_near_contact=False, tol_rel=1e-6):
        a = np.ascontiguousarray(a, dtype=self.dtype).reshape(-1, self.D)
        b = np.ascontiguousarray(b, dtype=self.dtype).reshape(-1, self.D)
        ok = np.empty(a.shape[0], dtype=np.uint8)
        nc = np.empty(a.shape[0], dtype=np.uint8) if with_near_contact else None
        states = C.c_uint64()
        sp = self.space.ref if self.space is not None else None
        self.orc.lib.orc_link_batch(self.h, sp, _ptr(a), _ptr(b), a.shape[0], float(self.step), _ptr(ok), _ptr(nc), float(tol_rel), C.byref(states))
        self.last_states = states.value
        return (ok, nc) if with_near_contact else ok

    def use_predicates(self, on=True):
        """Decide triangle pairs with the orientation-predicate formulation (oracle.hpp triTriPredicates)."""
        self.orc.lib.orc_mesh_use_predicates(self.h, int(on))

    def counters(self):
        out = (C.c_uint64 * 4)()
        self.orc.lib.orc_geom_counters(self.h, out)
        return {"states": out[0], "bv_tests": out[1], "tri_tests": out[2], "items": out[3]}

    def __del__(self):
        try:
            self.orc.lib.orc_geom_destroy(self.h)
        except Exception:
            pass


_oracle = None


def load() -> Oracle:
    global _oracle
    if _oracle is None:
        build()
        _oracle = Oracle(C.CDLL(str(LIB)))
    return _oracle
