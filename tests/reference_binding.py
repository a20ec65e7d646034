"""ctypes binding of oracle/_ref/libref.so: the reference's own hot-path headers compiled against the
stand-in Eigen / Nigh / png headers under oracle/shim (see oracle/ref_driver.cpp).  TEST INFRASTRUCTURE.
Only buildable where /root/reference exists; the GPU box uses the golden vectors generated from it
(tests/golden/reference_golden.npz, script tests/golden/make_reference_golden.py)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "oracle" / "_ref" / "libref.so"
LIB_PLANNER = ROOT / "oracle" / "_ref" / "libref_planner.so"
LIB_NAO = ROOT / "oracle" / "_ref" / "libref_nao.so"
REFERENCE = Path("/root/reference")
_P = C.c_void_p
VALID_CB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_float), C.c_void_p)


def available() -> bool:
    return REFERENCE.exists() or LIB.exists()


def load():
    if REFERENCE.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "ref"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return Ref(C.CDLL(str(LIB)), C.CDLL(str(LIB_PLANNER)), C.CDLL(str(LIB_NAO)))


def _p(a):
    return None if a is None else a.ctypes.data_as(_P)


class Ref:
    def __init__(self, lib, planner_lib=None, nao_lib=None):
        self.lib = lib
        self.planner_lib = planner_lib
        self.nao_lib = nao_lib

    # the reference's Nao-cup scenario (oracle/ref_nao.cpp): nao_clear / nao_link of demo/nao_cup/src/naocup.hpp
    def nao_clear(self, q, scalar=8):
        q = np.ascontiguousarray(q, dtype=np.float32 if scalar == 4 else np.float64).reshape(-1, 10)
        ok, col = np.empty(q.shape[0], np.uint8), np.empty(q.shape[0], np.uint8)
        self.nao_lib.ref_nao_clear(C.c_int(scalar), _p(q), C.c_uint32(q.shape[0]), _p(ok), _p(col))
        return ok, col

    def nao_link(self, a, b, scalar=8):
        dt = np.float32 if scalar == 4 else np.float64
        a, b = np.ascontiguousarray(a, dtype=dt).reshape(-1, 10), np.ascontiguousarray(b, dtype=dt).reshape(-1, 10)
        ok = np.empty(a.shape[0], np.uint8)
        self.nao_lib.ref_nao_link(C.c_int(scalar), _p(a), _p(b), C.c_uint32(a.shape[0]), _p(ok))
        return ok

    def nao_configs(self, scalar=8):
        out = [np.zeros(10) for _ in range(4)]
        self.nao_lib.ref_nao_configs(C.c_int(scalar), *[_p(a) for a in out])
        return tuple(out)  # init, lo, hi, target

    def prrt_grid(self, occ, lo, hi, start, goal, goal_radius, goal_bias, rng, uniforms, capacity=1 << 16):
        """The reference's Planner<Scenario, PRRT<single_threaded>> (oracle/ref_planner.cpp) on an occupancy grid, one
        iteration per row of `uniforms` ([n, 3]: goal-bias draw, x, y).  Returns (states, parents, goal_node)."""
        occ = np.ascontiguousarray(occ, dtype=np.uint8)
        lo, hi, start, goal = (np.ascontiguousarray(x, dtype=np.float64) for x in (lo, hi, start, goal))
        u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(-1, 3)
        states = np.empty((capacity, 2), np.float64)
        parents = np.empty(capacity, np.uint32)
        n, goal_node, used = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        rc = self.planner_lib.ref_prrt_grid(occ.shape[1], occ.shape[0], _p(occ), _p(lo), _p(hi), _p(start), _p(goal), C.c_double(goal_radius),
                                            C.c_double(goal_bias), C.c_double(rng), _p(u), C.c_uint32(u.shape[0]), _p(states), _p(parents),
                                            C.c_uint32(capacity), C.byref(n), C.byref(goal_node), C.byref(used))
        assert rc == 0 and used.value == u.shape[0], f"ref_prrt_grid rc={rc}, used {used.value} of {u.shape[0]}"
        return states[: n.value].copy(), parents[: n.value].copy(), goal_node.value

    def interpolate(self, kind, a, b, t):
        dt = np.float32 if kind == "se3_f32" else np.float64
        a, b, t = (np.ascontiguousarray(x, dtype=dt) for x in (a, b, t))
        out = np.empty_like(a)
        getattr(self.lib, f"ref_interpolate_{kind}")(_p(a), _p(b), _p(t), C.c_uint32(t.shape[0]), _p(out))
        return out

    def dmv_se3(self, frm, to, step, valid_fn):
        frm = np.ascontiguousarray(frm, dtype=np.float32)
        to = np.ascontiguousarray(to, dtype=np.float32)
        n = frm.shape[0]
        ok = np.empty(n, dtype=np.uint8)
        cnt = np.empty(n, dtype=np.uint64)

        def cb(ptr, _user):
            return int(valid_fn(np.ctypeslib.as_array(ptr, shape=(7,)).copy()))

        self.lib.ref_dmv_se3_f32(_p(frm), _p(to), C.c_uint32(n), C.c_float(step), VALID_CB(cb), None, _p(ok), _p(cnt))
        return ok, cnt

    def grid(self, occ, a, b):
        occ = np.ascontiguousarray(occ, dtype=np.uint8)
        a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
        va, ln = np.empty(a.shape[0], np.uint8), np.empty(a.shape[0], np.uint8)
        self.lib.ref_grid(occ.shape[1], occ.shape[0], _p(occ), _p(a), _p(b), C.c_uint32(a.shape[0]), _p(va), _p(ln))
        return va, ln

    def holonomic(self, circles, rects, a, b):
        c = np.ascontiguousarray(circles, dtype=np.float64).reshape(-1, 3)
        r = np.ascontiguousarray(rects, dtype=np.float64).reshape(-1, 4)
        a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
        va, ln = np.empty(a.shape[0], np.uint8), np.empty(a.shape[0], np.uint8)
        self.lib.ref_holonomic(c.shape[0], _p(c), r.shape[0], _p(r), _p(a), _p(b), C.c_uint32(a.shape[0]), _p(va), _p(ln))
        return va, ln

    def linkarm(self, lengths, radius, circles, a, b):
        ln_ = np.ascontiguousarray(lengths, dtype=np.float64)
        c = np.ascontiguousarray(circles, dtype=np.float64).reshape(-1, 3)
        a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
        va, ln = np.empty(a.shape[0], np.uint8), np.empty(a.shape[0], np.uint8)
        rc = self.lib.ref_linkarm(ln_.shape[0], _p(ln_), C.c_double(radius), c.shape[0], _p(c), _p(a), _p(b), C.c_uint32(a.shape[0]), _p(va), _p(ln))
        assert rc == 0, "unsupported link count"
        return va, ln

    def goal_l2_3(self, goal, radius, q):
        goal = np.ascontiguousarray(goal, dtype=np.float64)
        q = np.ascontiguousarray(q, dtype=np.float64)
        is_goal, d = np.empty(q.shape[0], np.uint8), np.empty(q.shape[0], np.float64)
        self.lib.ref_goal_l2_3(_p(goal), C.c_double(radius), _p(q), C.c_uint32(q.shape[0]), _p(is_goal), _p(d))
        return is_goal, d
