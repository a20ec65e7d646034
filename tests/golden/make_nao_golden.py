#!/usr/bin/env python3
"""Generate tests/golden/nao_golden.npz: decisions of the REFERENCE'S OWN Nao-cup code (demo/nao_cup/src/{naocup,collide,
linear}.hpp compiled from /root/reference against oracle/shim/Eigen, oracle/ref_nao.cpp) on seeded inputs, float and double.
Can only run where /root/reference exists; the vectors are committed so that the CPU suite and the GPU box check the oracle
and the CUDA kernels against the reference's decisions.   Run from the repo root:  python tests/golden/make_nao_golden.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from mpt_b200 import workloads as W  # noqa: E402
from tests import reference_binding  # noqa: E402


def main():
    ref = reference_binding.load()
    out = {}
    for scalar, tag, dt in ((8, "f64", np.float64), (4, "f32", np.float32)):
        init, lo, hi, target = ref.nao_configs(scalar)
        out[f"configs_{tag}"] = np.stack([init, target, lo, hi])
        q = W.nao_states(16384, 101, dtype=dt)
        q[0], q[1] = init.astype(dt), target.astype(dt)
        ok, col = ref.nao_clear(q, scalar)
        out[f"q_{tag}"], out[f"clear_{tag}"], out[f"collision_{tag}"] = q, ok, col
        # edges: from CLEAR states (what a planner links), plus a block of long ones across the joint range
        pool = W.nao_states(60000, 202, dtype=dt)
        pool = pool[ref.nao_clear(pool, scalar)[0] == 1][:3072]
        rng = np.random.default_rng(303)
        b = np.clip(pool + rng.normal(0, 0.12, pool.shape) * rng.random((pool.shape[0], 1)), W.NAO_LO, W.NAO_HI).astype(dt)
        la, lb = W.nao_states(256, 404, dtype=dt), W.nao_states(256, 405, dtype=dt)
        a = np.ascontiguousarray(np.concatenate([pool, la]))
        b = np.ascontiguousarray(np.concatenate([b, lb]))
        out[f"a_{tag}"], out[f"b_{tag}"], out[f"link_{tag}"] = a, b, ref.nao_link(a, b, scalar)
        print(tag, "clear", ok.mean(), "collision", col.mean(), "edges", a.shape[0], "link", out[f"link_{tag}"].mean())
    path = Path(__file__).with_name("nao_golden.npz")
    np.savez_compressed(path, **out)
    print(path, path.stat().st_size)


if __name__ == "__main__":
    main()
