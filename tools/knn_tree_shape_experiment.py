"""CPU experiment (numpy, no GPU): how many 32-point leaves / 32-child blocks of the SE(3) index a k = 16 query
MUST visit, as a function of (i) the split rule of the build and (ii) the bound stored per node.

For a sample of C5 queries the final k-th distance R_k is taken from an exhaustive scan; a node is "needed" when its
lower bound is <= R_k.  That is the floor of any best-first search over that structure (the real search starts with a
larger threshold and visits somewhat more: the shipped r1 structure measures 321 leaves + 54 inner nodes per query on
the device, its floor below is what the model says for the same structure).

Bounds compared for the rotation part of a node:
  box   AABB over the sign-canonicalised quaternion coefficients, max-corner dot (r1, knn_bvh.cuh childKey)
  cap   centre quaternion c + angular radius rho: max(0, acos|q.c| - rho)  (triangle inequality on RP^3)
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from mpt_b200 import workloads as W  # noqa: E402

W0, W1 = 50.0, 1.0


def canon(q, mode):
    q = q.copy()
    if mode == "w":
        s = np.where(q[:, 3] < 0, -1.0, 1.0)
    else:  # largest-magnitude coefficient positive
        j = np.abs(q[:, :4]).argmax(1)
        s = np.sign(q[np.arange(len(q)), j])
    q[:, :4] *= s[:, None]
    return q


def build(pts, coords, weights, levels=15):
    """binary median splits on the widest weighted coordinate of `coords`; returns the permutation"""
    n = len(pts)
    order = np.arange(n)
    seg = 1
    for lv in range(levels):
        size = n // seg
        c = coords[order].reshape(seg, size, -1)
        ext = (c.max(1) - c.min(1)) * weights[None, :]
        axis = ext.argmax(1)
        key = np.take_along_axis(c, axis[:, None, None].repeat(size, 1), 2)[:, :, 0]
        srt = np.argsort(key, axis=1, kind="stable")
        order = np.take_along_axis(order.reshape(seg, size), srt, 1).reshape(-1)
        seg *= 2
    return order


def node_bounds(p, group):
    """p [n,7] in tree order -> per node of `group` consecutive points: box lo/hi, cap centre / radius"""
    g = p.reshape(-1, group, 7)
    lo, hi = g.min(1), g.max(1)
    q = g[:, :, :4]
    # centre: dominant eigenvector is overkill; sign-align to the first member and average, twice
    ref = q[:, :1, :]
    s = np.sign((q * ref).sum(2, keepdims=True))
    s[s == 0] = 1
    c = (q * s).mean(1)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    s = np.sign((q * c[:, None, :]).sum(2, keepdims=True))
    s[s == 0] = 1
    c = (q * s).mean(1)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    rho = np.arccos(np.minimum(1.0, np.abs((q * c[:, None, :]).sum(2)))).max(1)
    return lo, hi, c, rho


def lb_box(qv, lo, hi):
    v = qv[:4]
    chi = np.where(v >= 0, hi[:, :4], lo[:, :4])
    clo = np.where(v >= 0, lo[:, :4], hi[:, :4])
    ad = np.minimum(1.0, np.maximum((chi * v).sum(1), -(clo * v).sum(1)))
    ad = np.maximum(ad, 0.0)
    return np.arccos(ad)


def lb_cap(qv, c, rho):
    return np.maximum(0.0, np.arccos(np.minimum(1.0, np.abs(c @ qv[:4]))) - rho)


def lb_trans(qv, lo, hi):
    e = np.maximum(np.maximum(lo[:, 4:] - qv[4:], qv[4:] - hi[:, 4:]), 0.0)
    return np.sqrt((e * e).sum(1))


def main():
    n = 1 << 20
    nq = int(sys.argv[1]) if len(sys.argv) > 1 else 192
    pts = W.se3_states(n, W.TREE_SEED).astype(np.float64)
    qs = W.se3_states(65536, W.QUERY_SEED).astype(np.float64)[:: 65536 // nq][:nq]
    t0 = time.time()
    rk = np.empty(nq)
    inball_t = np.empty(nq)
    for i, qv in enumerate(qs):
        d = W0 * np.arccos(np.minimum(1.0, np.abs(pts[:, :4] @ qv[:4]))) + W1 * np.linalg.norm(pts[:, 4:] - qv[4:], axis=1)
        rk[i] = np.partition(d, 15)[15]
        inball_t[i] = (np.linalg.norm(pts[:, 4:] - qv[4:], axis=1) <= rk[i]).sum()
    print(f"scan {time.time() - t0:.1f}s: mean R_k {rk.mean():.2f}, points within R_k of the translation alone {inball_t.mean():.0f}")

    variants = {}
    pw = canon(pts, "w")
    wq = np.array([W0] * 4 + [W1] * 3)
    variants["r1: coeff AABB split (w>=0), weights 50/1"] = (pw, pw, wq)
    for f in (0.5, 0.7, 1.4, 2.0):
        variants[f"coeff split, rotation extents x{f}"] = (pw, pw, wq * np.array([f] * 4 + [1] * 3))
    pm = canon(pts, "max")
    variants["coeff split, canonical = largest coeff positive"] = (pm, pm, wq)
    for name, (p, coords, wts) in variants.items():
        t0 = time.time()
        order = build(p, coords, wts)
        ps = p[order]
        levels = [node_bounds(ps, 32), node_bounds(ps, 1024), node_bounds(ps, 32768)]
        res = {}
        for kind in ("box", "cap"):
            cnt = np.zeros((nq, 3))
            for i, qv in enumerate(qs):
                qc = qv.copy()
                for l, (lo, hi, c, rho) in enumerate(levels):
                    r = lb_box(qc, lo, hi) if kind == "box" else lb_cap(qc, c, rho)
                    lb = W0 * r + W1 * lb_trans(qc, lo, hi)
                    cnt[i, l] = (lb <= rk[i]).sum()
            res[kind] = cnt.mean(0)
        rho0 = levels[0][3]
        ext_t = (levels[0][1][:, 4:] - levels[0][0][:, 4:]).mean()
        print(f"{name}: leaf cap radius mean {rho0.mean():.3f} rad ({W0 * rho0.mean():.1f} weighted), leaf translation extent {ext_t:.1f}")
        for kind, c in res.items():
            print(f"    {kind}: leaves needed {c[0]:.0f}, level-1 nodes {c[1]:.1f}, top nodes {c[2]:.1f}  -> blocks visited {1 + c[2] + c[1]:.1f}")
        print(f"    ({time.time() - t0:.0f}s)", flush=True)


if __name__ == "__main__":
    main()
