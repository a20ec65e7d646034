"""Nao-cup scenario on the device: states/s of valid and edges/s, midpoints/s of link, float and double (device-resident
inputs, CUDA events on the context's stream).  With `once`, a single wave for an ncu capture:
    ncu --set full --import-source on --clock-control none -k regex:bisectLinkKernel -s 1 -c 1 -o gpurun_out/r2_nao python tools/nao_time.py once"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import mpt_b200 as m  # noqa: E402
from mpt_b200 import workloads as W  # noqa: E402

once = len(sys.argv) > 1 and sys.argv[1] == "once"
ctx = m.Context(0)
dev = torch.device("cuda:0")
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
FLOPS_PER_STATE = 6500  # unfused adds / multiplies / compares of one full clear() (nao.cuh), counted from the SASS


def timed(fn, reps):
    ts = []
    for it in range(reps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.sync()
        with torch.cuda.stream(stream):
            e0.record(stream)
            fn()
            e1.record(stream)
        ctx.sync()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


for scalar, dt, tdt in ((m.F32, np.float32, torch.float32), (m.F64, np.float64, torch.float64)):
    sc = m.Scenario.nao_cup(ctx, scalar)
    q = W.nao_states(1 << 20, 3, dtype=dt)
    ok = sc.valid(q)
    clear = q[ok == 1]
    dq = torch.from_numpy(q).to(dev)
    dok = torch.empty(q.shape[0], dtype=torch.uint8, device=dev)
    if not once:
        ms = timed(lambda: sc.valid_dev(dq.data_ptr(), q.shape[0], dok.data_ptr()), 5)
        print(f"{dt.__name__}: valid {q.shape[0]} states (mixed, {ok.mean():.3f} clear) {ms:.3f} ms = {q.shape[0] / ms / 1e3:.1f} M states/s")
        dc = torch.from_numpy(np.ascontiguousarray(np.tile(clear, (q.shape[0] // clear.shape[0] + 1, 1))[: q.shape[0]])).to(dev)
        ms = timed(lambda: sc.valid_dev(dc.data_ptr(), q.shape[0], dok.data_ptr()), 5)
        print(f"{dt.__name__}: valid {q.shape[0]} CLEAR states (every pair test runs) {ms:.3f} ms = {q.shape[0] / ms / 1e3:.1f} M states/s "
              f"~ {q.shape[0] * FLOPS_PER_STATE / ms / 1e9:.1f} TFLOP/s unfused")
    rng = np.random.default_rng(4)
    for reach in (0.3,) if once else (0.1, 0.3, 1.0):
        E = 65536
        a = clear[rng.integers(0, clear.shape[0], E)]
        b = np.clip(a + rng.normal(0, reach / np.sqrt(10), a.shape), W.NAO_LO, W.NAO_HI).astype(dt)
        da, db = torch.from_numpy(np.ascontiguousarray(a)).to(dev), torch.from_numpy(np.ascontiguousarray(b)).to(dev)
        dl = torch.empty(E, dtype=torch.uint8, device=dev)
        ms = timed(lambda: sc.link_dev(da.data_ptr(), db.data_ptr(), E, dl.data_ptr()), 1 if once else 5)
        st = sc.last_stats()
        print(f"{dt.__name__}: link {E} edges reach {reach}: {ms:.3f} ms = {E / ms / 1e3:.2f} M edges/s, {st['prim_tests']} midpoints "
              f"= {st['prim_tests'] / ms / 1e3:.1f} M states/s, valid fraction {dl.float().mean().item():.3f}")
    if once:
        break
