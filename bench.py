#!/usr/bin/env python3
"""bench.py -- throughput of the planning hot path on B200 (BASELINE.json: "batched kNN queries/s and
edge-validity checks/s at 1/2/4/8 B200").

One "step" = one sample wave of the planner's hot path on synthetic input (BASELINE.json configs[4],
the only configuration the metric is quoted on that needs a GPU):
    * batched SE(3) kNN: Q = 65,536 queries, k = 16, against a tree of N = 1,048,576 nodes
      (SO(3) weight 50, L2 weight 1, float32)            -> `value` = queries/s
    * batched edge validity: E = 65,536 SE(3) edges through DiscreteMotionValidator against a synthetic
      rigid-body mesh pair (~1k robot / ~4k environment triangles) -> `edges_per_s`
Multi-GPU (torchrun, one rank per GPU), as north_star splits the path:
    * kNN: the TREE is sharded spatially across the ranks and ONE 65,536-query wave is answered by all of them
      (mptg_knn_query_sharded: root bounds -> home search -> bounded search -> NCCL exchange -> merge); every rank
      returns its slice of the wave.  `value` = 65,536 / time of the slowest rank: "scaling": "strong".
    * edges: independent units -- every rank checks its own 65,536-edge wave against replicated meshes, no collective
      (`edges_per_s` is the aggregate, weak).
    * `replicated_tree`: the zero-collective alternative for trees that fit one HBM (every rank its own wave), for reference.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_TREE = 1_048_576
Q_WAVE = 65_536
E_WAVE = 65_536
K_NN = 16
SO3_W, L2_W = 50.0, 1.0
MESH_LO, MESH_HI = -45.0, 45.0
EDGE_TRANS, EDGE_ANGLE = 12.0, 0.5  # a steered sample: <= 12 units and <= 0.5 rad from the tree node
ALGO_BYTES_KNN = N_TREE * 7 * 4 + Q_WAVE * 7 * 4 + Q_WAVE * K_NN * (4 + 4)  # 39,583,744
F_BV, F_TRI = 82.0, 170.0  # flop per box-pair test (box transform + world-axis stage as implemented; the robot-frame
# stage adds 39 more when it is reached -- not counted) / per 17-axis SAT


def _profile_json(name: str):
    try:
        return json.loads((ROOT / "profiles" / name).read_text())
    except Exception:
        return {}


def measured_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture."""
    try:
        return float(_profile_json("r2_traffic.json")[kernel]["traffic_bytes_per_launch"])
    except Exception:
        return None


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed regions (B200_PROFILING.md recipe): NVML polled every millisecond from
    a thread (the timed region of the default run is ~20 ms, too short for `nvidia-smi -lms 100`, which is the fallback)."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.index = index
        self.rows = []  # (sm_mhz, reasons bitmask)
        self.max_mhz = None
        self.proc = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.source = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self.stop_flag.is_set():
                    try:
                        self.rows.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), int(get_reasons(h))))
                    except Exception:
                        pass
                    time.sleep(0.001)

            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            self.source = "nvml, 1 ms"
            return
        except Exception:
            self.thread = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            self.source = "nvidia-smi -lms 100"
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            try:
                mask = sum(bit for i, (_, bit) in enumerate(self.REASONS) if len(c) > 3 + i and c[3 + i].lower().startswith("active"))
                self.rows.append((float(c[0]), mask))
                self.max_mhz = max(self.max_mhz or 0.0, float(c[1]))
            except (ValueError, IndexError):
                pass

    def stop(self):
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=1)
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm = [r[0] for r in self.rows]
        reasons = [n for n, bit in self.REASONS if any(r[1] & bit for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                "source": self.source, "covers": "the device-timed steps and the two end-to-end passes"}


def workload(rank: int):
    """Synthetic inputs of one rank (host, numpy)."""
    from mpt_b200 import workloads as W

    tree = W.se3_states(N_TREE, W.TREE_SEED)
    queries = W.se3_states(Q_WAVE, W.QUERY_SEED + 1000 * rank)
    robot, env, vmin, vmax = W.alpha_puzzle_like(env_tris_target=4000, robot_tris_target=1000)
    step = W.se3_step_size(vmin, vmax, SO3_W)
    ea, eb = W.se3_edges(E_WAVE, W.EDGE_SEED + 1000 * rank, MESH_LO, MESH_HI, EDGE_TRANS, EDGE_ANGLE)
    return tree, queries, robot, env, step, ea, eb


CONFIG = {
    "workload": "synthetic SE(3) kNN + edge-check sweep (BASELINE.json configs[4])",
    "tree_nodes": N_TREE, "queries_per_wave": Q_WAVE, "k": K_NN, "edges_per_wave": E_WAVE,
    "space": "SE3 (SO3 weight 50, L2 weight 1), float32",
    "mesh_pair": "synthetic bent-tube robot (~1k tris) vs environment (~4k tris), DiscreteMotionValidator resolution 0.01",
    "edge_length": f"<= {EDGE_TRANS} units, <= {EDGE_ANGLE} rad",
    "l2": "256 MiB memset between steps (L2 flush), outside the per-step event intervals",
    "parallelism": "one rank per GPU; kNN: tree sharded spatially, one wave answered by all ranks (NCCL exchange + merge, strong scaling); "
                   "edges: one wave per rank, meshes replicated, no collective",
}


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the oracle (a port of the reference path; the reference itself cannot be built here --
    Eigen/Nigh/FCL absent) on all host threads, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests import oracle_binding

    orc = oracle_binding.load()
    orc.set_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host thread
    tree, queries, robot, env, step, ea, eb = workload(0)
    sp = oracle_binding.se3_space(SO3_W, L2_W)  # descriptor built without loading libmptg.so: nothing of the product runs here
    t0 = time.perf_counter()
    otree = orc.tree(sp, tree)
    build_s = time.perf_counter() - t0
    omesh = orc.mesh_pair(robot, env, sp, step)
    qs, es = args.cpu_queries, args.cpu_edges
    knn_t, edge_t = [], []
    for it in range(args.warmup + args.steps):
        sel = slice((it * qs) % (Q_WAVE - qs + 1), (it * qs) % (Q_WAVE - qs + 1) + qs)
        t0 = time.perf_counter()
        otree.knn(queries[sel], K_NN)
        t1 = time.perf_counter()
        esel = slice((it * es) % (E_WAVE - es + 1), (it * es) % (E_WAVE - es + 1) + es)
        omesh.link(ea[esel], eb[esel])
        t2 = time.perf_counter()
        if it >= args.warmup:
            knn_t.append(t1 - t0)
            edge_t.append(t2 - t1)
    qps = qs / float(np.mean(knn_t))
    eps = es / float(np.mean(edge_t))
    sample = f"{qs} of {Q_WAVE} queries and {es} of {E_WAVE} edges per step (tree build {build_s:.1f}s untimed)"
    # solve time: the reference's own multi-threaded planner classes on the occupancy-grid scenario (same map as the
    # GPU arm's secondary.device_prrt_grid)
    from mpt_b200 import workloads as W

    occ = W.synthetic_grid()
    free = np.argwhere(occ == 0)
    planner = reference_planner_cpu(occ, free[len(free) // 7][::-1].astype(np.float64), free[-len(free) // 9][::-1].astype(np.float64), 12.0, 200.0)
    line = {
        "impl": "reference", "metric": "knn_queries_per_s", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(knn_t) + np.mean(edge_t)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": CONFIG, "edges_per_s": eps,
        "cpu_baseline": {"value": qps, "unit": "queries/s", "edges_per_s": eps, "cores": orc.threads, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if planner:
        line["reference_planner_cpu_grid"] = planner
    emit(line)


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import mpt_b200 as m

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    tree, queries, robot, env, step, ea, eb = workload(rank)
    ctx = m.Context(local)
    sp = m.se3_space(SO3_W, L2_W)
    comm = None
    q_first, q_count = 0, Q_WAVE
    if world == 1:
        nn = m.Nearest(ctx, sp, N_TREE)
        nn.insert(tree)
        nn.build_index()
    else:
        # the tree sharded spatially (fixed halving of the translation bounds), ONE wave for all ranks (rank 0's queries)
        from mpt_b200 import sharding
        from mpt_b200 import workloads as W

        queries = W.se3_states(Q_WAVE, W.QUERY_SEED)
        uid = [m.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = m.Comm(ctx, uid[0], rank, world)
        cells = sharding.spatial_cells(tree, (4, 5, 6), -100.0, 100.0, world)
        ids = np.nonzero(cells == rank)[0].astype(np.uint32)
        nn = m.Nearest(ctx, sp, len(ids) + 64, m.KNN_BVH)
        nn.insert_ids(tree[ids], ids)
        comm.sync(nn)  # index the shard, exchange the shards' top-level boxes
        q_first, q_count = comm.slice(Q_WAVE)
    mesh = m.Scenario.mesh_pair(ctx, robot, env, sp, step)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    # device-resident inputs / outputs for `value`
    d_q = torch.from_numpy(queries).to(dev)
    d_idx = torch.empty((q_count, K_NN), dtype=torch.int32, device=dev)
    d_dist = torch.empty((q_count, K_NN), dtype=torch.float32, device=dev)
    d_cnt = torch.empty(q_count, dtype=torch.int32, device=dev)
    d_ea, d_eb = torch.from_numpy(ea).to(dev), torch.from_numpy(eb).to(dev)
    d_ok = torch.empty(E_WAVE, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # pinned host buffers for `e2e` (the reference-facing host-pointer C-ABI calls)
    h_q = torch.from_numpy(queries).pin_memory()
    h_idx = torch.empty((q_count, K_NN), dtype=torch.int32).pin_memory()
    h_dist = torch.empty((q_count, K_NN), dtype=torch.float32).pin_memory()
    h_cnt = torch.empty(q_count, dtype=torch.int32).pin_memory()
    h_ea, h_eb = torch.from_numpy(ea).pin_memory(), torch.from_numpy(eb).pin_memory()
    h_ok = torch.empty(E_WAVE, dtype=torch.uint8).pin_memory()
    # ... and pageable ones: what a reference-side caller hands over (std::vector memory)
    p_q, p_ea, p_eb = queries.copy(), ea.copy(), eb.copy()
    p_idx, p_dist = np.empty((q_count, K_NN), np.uint32), np.empty((q_count, K_NN), np.float32)
    p_cnt, p_ok = np.empty(q_count, np.uint32), np.empty(E_WAVE, np.uint8)
    torch.cuda.synchronize()

    def knn_dev():
        if comm is None:
            nn.nearest_dev(d_q.data_ptr(), Q_WAVE, K_NN, -1.0, d_idx.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr())
        else:
            comm.nearest_dev(nn, d_q.data_ptr(), Q_WAVE, K_NN, -1.0, d_idx.data_ptr(), d_dist.data_ptr(), d_cnt.data_ptr())

    def knn_host(q_ptr, i_ptr, d_ptr, c_ptr):
        if comm is None:
            nn.nearest_host_into(q_ptr, Q_WAVE, K_NN, -1.0, i_ptr, d_ptr, c_ptr)
        else:
            comm.nearest_host_into(nn, q_ptr, Q_WAVE, K_NN, -1.0, i_ptr, d_ptr, c_ptr)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def step_device(record):
        with torch.cuda.stream(stream):
            flush.zero_()
            e0, e1, e2 = ev(), ev(), ev()
            e0.record(stream)
            knn_dev()
            e1.record(stream)
            mesh.link_dev(d_ea.data_ptr(), d_eb.data_ptr(), E_WAVE, d_ok.data_ptr())
            e2.record(stream)
        if record is not None:
            record.append((e0, e1, e2))

    def step_e2e(record):
        with torch.cuda.stream(stream):
            flush.zero_()
            e0, e1, e2 = ev(), ev(), ev()
            e0.record(stream)
        knn_host(h_q.data_ptr(), h_idx.data_ptr(), h_dist.data_ptr(), h_cnt.data_ptr())
        e1.record(stream)
        mesh.link_host_into(h_ea.data_ptr(), h_eb.data_ptr(), E_WAVE, h_ok.data_ptr())
        e2.record(stream)
        if record is not None:
            record.append((e0, e1, e2))

    def step_e2e_pageable(record):
        with torch.cuda.stream(stream):
            flush.zero_()
            e0, e1, e2 = ev(), ev(), ev()
            e0.record(stream)
        knn_host(p_q.ctypes.data, p_idx.ctypes.data, p_dist.ctypes.data, p_cnt.ctypes.data)
        e1.record(stream)
        mesh.link_host_into(p_ea.ctypes.data, p_eb.ctypes.data, E_WAVE, p_ok.ctypes.data)
        e2.record(stream)
        if record is not None:
            record.append((e0, e1, e2))

    def timed(step_fn):
        for _ in range(args.warmup):
            step_fn(None)
        barrier()
        launches0 = ctx.launches
        rec = []
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_fn(rec)
        ctx.sync()
        barrier()
        wall = time.perf_counter() - t0
        knn_ms = sum(a.elapsed_time(b) for a, b, _ in rec) / len(rec)
        edge_ms = sum(b.elapsed_time(c) for _, b, c in rec) / len(rec)
        t = torch.tensor([knn_ms, edge_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), wall, ctx.launches - launches0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    knn_ms, edge_ms, wall, launches = timed(step_device)
    e2e_knn_ms, e2e_edge_ms, _, _ = timed(step_e2e)
    pg_knn_ms, pg_edge_ms, _, _ = timed(step_e2e_pageable)
    clocks = sampler.stop() if rank == 0 else None

    knn_stats = nn.last_stats()
    mesh_stats = mesh.last_stats()
    # parity spot check of the timed outputs against the host-pointer path (same library, both paths)
    ok_dev = d_ok.cpu().numpy()
    assert np.array_equal(ok_dev, h_ok.numpy()), "device-pointer and host-pointer edge results differ"
    assert np.array_equal(d_idx.cpu().numpy(), h_idx.numpy()), "device-pointer and host-pointer kNN results differ"
    assert np.array_equal(h_idx.numpy().view(np.uint32), p_idx) and np.array_equal(h_ok.numpy(), p_ok), "pinned and pageable results differ"

    extra = {}
    if rank == 0 and world == 1 and not args.no_secondary:  # one-GPU figures: not repeated while the other ranks of a larger run wait
        extra["secondary"] = secondary(ctx, torch, dev, stream)
    if world > 1:
        extra["replicated_tree"] = bench_replicated_tree(args, ctx, sp, tree, dev, stream, world, rank)

    if rank != 0:
        if comm is not None:
            comm.close()
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = peaks()
    # kNN: one structure at N = 1; at N > 1 ONE wave answered by all ranks (strong scaling).  Edges: a wave per rank.
    qps = Q_WAVE / (knn_ms * 1e-3)
    eps = world * E_WAVE / (edge_ms * 1e-3)
    achieved = ALGO_BYTES_KNN / (knn_ms * 1e-3) / 1e9
    launches_before_probe = ctx.launches
    fp32_peak = fp32_probe(ctx)
    assert ctx.launches == launches_before_probe + 3
    sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    issue = _profile_json("r2_knn_issue.json")  # warp instructions per query of the tree kernel, from the committed ncu capture
    issue_slots = ctx.sm_count * 4 * sm_mhz * 1e6  # warp instructions the GPU can issue per second (4 schedulers per SM)
    wipq = issue.get("warp_instructions_per_query")
    # mesh roofline numerator: box-pair / triangle-pair tests of the ORACLE's traversal of the same edge wave
    oc = _profile_json("r2_mesh_oracle_counts.json")
    own_flops = mesh_stats["bv_tests"] * F_BV + mesh_stats["prim_tests"] * F_TRI
    edge_flops = (oc["bv_tests"] * F_BV + oc["tri_tests"] * F_TRI) if oc else own_flops
    line = {
        "metric": "knn_queries_per_s", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": knn_ms + edge_ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
        "knn_ms": knn_ms, "edge_ms": edge_ms, "edges_per_s": eps, "edges_scaling": "weak (one 65,536-edge wave per rank)",
        "edge_states_per_s": world * mesh_stats["states"] / (edge_ms * 1e-3),
        "roofline": {
            "kernel": "knnSe3Kernel<1>", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved / hbm_peak, "traffic": measured_traffic("knnSe3Kernel<1>"), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": ALGO_BYTES_KNN,
            "note": "exact kNN in a 6-dimensional space with 10 points per dimension is bound by node and point tests (instruction "
                    "issue), not by compulsory HBM bytes (DESIGN.md 4.2); the issue-slot roofline below is the one that moves",
            "distance_evals_per_query": knn_stats["distance_evals"] / Q_WAVE,
            "nodes_visited_per_query": knn_stats["nodes_visited"] / Q_WAVE,
            # issue-slot roofline: warp instructions per query (ncu smsp__inst_executed of the same kernel on the same wave,
            # profiles/r2_knn_issue.json) x Q / (SMs x 4 schedulers x SM clock during this run) = time at one instruction
            # per scheduler per cycle; frac = that time / measured time
            "issue": {"warp_instructions_per_query": wipq, "issue_slots_per_s": issue_slots, "sm_mhz": sm_mhz,
                      "floor_ms": (wipq * Q_WAVE / issue_slots * 1e3) if wipq else None,
                      "frac": (wipq * Q_WAVE / issue_slots / (knn_ms * 1e-3)) if (wipq and world == 1) else None,
                      "source": issue.get("source"), "r1": {"warp_instructions_per_query": 31600, "ms": 2.29}},
            # SURVEY.md 8(d) honesty check: which roofline binds.  F_pair = 21 flop per SE(3) distance evaluation.
            "fp32": {"achieved": knn_stats["distance_evals"] * 21.0 / (knn_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": (knn_stats["distance_evals"] * 21.0 / (knn_ms * 1e-3) / 1e12 / fp32_peak) if fp32_peak else None,
                     "flop_per_distance_eval": 21.0},
        },
        "roofline_edges": {
            "kernel": "meshFlatKernel", "bound": "fp32", "achieved": edge_flops / (edge_ms * 1e-3) / 1e12, "peak": fp32_peak,
            "unit": "TFLOP/s", "frac": edge_flops / (edge_ms * 1e-3) / 1e12 / fp32_peak if fp32_peak else None,
            "peak_source": "FFMA microbenchmark of this library on this GPU, same run (mptg_probe_fp32_tflops)",
            "algorithmic_flops_per_launch": edge_flops,
            "numerator": "box-pair and triangle-pair tests of the ORACLE's traversal of the same 65,536 edges (profiles/r2_mesh_oracle_counts.json, "
                         "tools/mesh_oracle_counts.py), at the flop cost of the tests as implemented",
            "oracle_bv_tests": oc.get("bv_tests"), "oracle_tri_tests": oc.get("tri_tests"), "oracle_states": oc.get("states"),
            "kernel_bv_tests": mesh_stats["bv_tests"], "kernel_tri_tests": mesh_stats["prim_tests"], "kernel_states": mesh_stats["states"],
            "work_inflation": {"bv": mesh_stats["bv_tests"] / oc["bv_tests"], "tri": mesh_stats["prim_tests"] / oc["tri_tests"],
                               "flops": own_flops / edge_flops} if oc else None,
            "flop_per_bv_test": F_BV, "flop_per_tri_test": F_TRI,
            "traffic": measured_traffic("meshFlatKernel"),
        },
        "e2e": {
            "value": Q_WAVE / (e2e_knn_ms * 1e-3), "unit": "queries/s", "edges_per_s": world * E_WAVE / (e2e_edge_ms * 1e-3),
            "ms_per_step": e2e_knn_ms + e2e_edge_ms,
            "h2d_bytes_per_step": int(h_q.numel() * 4 + h_ea.numel() * 4 + h_eb.numel() * 4),
            "d2h_bytes_per_step": int(h_idx.numel() * 4 + h_dist.numel() * 4 + h_cnt.numel() * 4 + h_ok.numel()),
            "api": ("mptg_knn_query" if world == 1 else "mptg_knn_query_sharded") + " + mptg_link_batch with pinned host buffers",
            "pageable": {"value": Q_WAVE / (pg_knn_ms * 1e-3), "edges_per_s": world * E_WAVE / (pg_edge_ms * 1e-3),
                         "ms_per_step": pg_knn_ms + pg_edge_ms, "note": "the same calls with ordinary (pageable) host memory, as a "
                         "reference-side caller's std::vector buffers would be"},
        },
        "gpu_launches": launches, "clocks": clocks, "wall_s_timed_region": wall, "edge_valid_fraction": float(ok_dev.mean()),
        **extra,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, sp, tree, queries, robot, env, step, ea, eb)
    emit(line)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def secondary(ctx, torch, dev, stream):
    """The other back-ends of the path (C1, C2, C4), device-resident inputs, rank 0 only,
    outside the main timed region.  Reported for coverage; `value` stays the C5 kNN figure."""
    import time

    import mpt_b200 as m
    from mpt_b200 import workloads as W

    out = {}

    def time_link(sc, a, b, reps=5):
        da, db = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        ok = torch.empty(a.shape[0], dtype=torch.uint8, device=dev)
        ts = []
        for it in range(reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                sc.link_dev(da.data_ptr(), db.data_ptr(), a.shape[0], ok.data_ptr())
                e1.record(stream)
            ctx.sync()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        st = sc.last_stats()
        ms = float(np.mean(ts))
        return {"edges_per_s": a.shape[0] / (ms * 1e-3), "ms": ms, "probes_per_s": st["prim_tests"] / (ms * 1e-3) if st["prim_tests"] else None,
                "valid_fraction": float(ok.float().mean().item())}

    occ = W.synthetic_grid()  # 3976 x 2603, the shipped PNG's size
    grid = m.Scenario.grid(ctx, occ, m.F64)
    for name, max_len in (("grid_edges_range64", 64.0), ("grid_edges_unbounded", None)):
        a, b = W.grid_edges(E_WAVE, occ.shape[1], occ.shape[0], 31, max_len)
        out[name] = time_link(grid, a, b)
    for n_links in (8, 16, 32):
        lengths, radius, circles = W.link_arm_scene(n_links)
        arm = m.Scenario.link_arm(ctx, lengths, radius, circles, m.F64)
        a, b = W.arm_edges(E_WAVE, n_links, 41, 0.5)
        out[f"link_arm_{n_links}_edges"] = time_link(arm, a, b)
    # the fifth scenario of the reference: Nao with cup and ball, 10 joints; edges between CLEAR
    # configurations a few tenths of a radian apart (what a planner with a range links), one flat list of midpoints
    issue_nao = _profile_json("r2_nao_issue.json")
    for scalar, tag, dt in ((m.F32, "f32", np.float32), (m.F64, "f64", np.float64)):
        nao = m.Scenario.nao_cup(ctx, scalar)
        pool = W.nao_states(1 << 20, 3, dtype=dt)
        pool = pool[nao.valid(pool) == 1]
        rng = np.random.default_rng(4)
        a = np.ascontiguousarray(pool[rng.integers(0, pool.shape[0], E_WAVE)])
        b = np.ascontiguousarray(np.clip(a + rng.normal(0, 0.3 / np.sqrt(10), a.shape), W.NAO_LO, W.NAO_HI).astype(dt))
        r = time_link(nao, a, b)
        r["midpoints_per_s"] = r.pop("probes_per_s")
        tipp = issue_nao.get(f"thread_instructions_per_midpoint_{tag}")
        if tipp and r["midpoints_per_s"]:
            lanes = ctx.sm_count * 4 * 32 * 1.965e9  # thread instructions per second at one warp instruction per scheduler per clock
            r["issue"] = {"thread_instructions_per_midpoint": tipp, "frac_of_issue_slots": r["midpoints_per_s"] * tipp / lanes,
                          "fp32_instructions_per_midpoint": issue_nao.get(f"fp32_instructions_per_midpoint_{tag}"),
                          "source": issue_nao.get("source"),
                          "note": "every test of this scenario decides, so its arithmetic is unfused (one flop per FP32 instruction): "
                                  "the FFMA yardstick's flop rate is out of reach by construction, its instruction rate is the bound"}
        if scalar == m.F64:  # kept for the cpu_baseline leg: the CPU restatement of the same edges on the host threads
            _NAO_SAMPLE["a"], _NAO_SAMPLE["b"], _NAO_SAMPLE["ok"] = a[:8192].copy(), b[:8192].copy(), nao.link(a[:8192], b[:8192])
        out[f"nao_cup_edges_{tag}"] = r
        nao.close()

    def time_knn(sp, pts, q, strat, np_dtype=np.float32):
        nn = m.Nearest(ctx, sp, pts.shape[0], strat)
        nn.insert(pts)
        dq = torch.from_numpy(q).to(dev)
        di = torch.empty((q.shape[0], K_NN), dtype=torch.int32, device=dev)
        dd = torch.empty((q.shape[0], K_NN), dtype=torch.float32 if np_dtype == np.float32 else torch.float64, device=dev)
        ts = []
        for it in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                nn.nearest_dev(dq.data_ptr(), q.shape[0], K_NN, -1.0, di.data_ptr(), dd.data_ptr())
                e1.record(stream)
            ctx.sync()
            if it >= 2:
                ts.append(e0.elapsed_time(e1))
        nn.close()
        return {"queries_per_s": q.shape[0] / (float(np.mean(ts)) * 1e-3), "ms": float(np.mean(ts))}

    # planar L2 kNN at a planner-realistic size (tiled scan) and at 1M (tree)
    for n_pts, label, strat in ((1 << 14, "knn_l2_2d_16k_brute", m.KNN_BRUTE), (1 << 14, "knn_l2_2d_16k_auto", m.KNN_AUTO),
                                (1 << 20, "knn_l2_2d_1m_tree", m.KNN_AUTO)):
        pts = W.box_states(n_pts, 2, 51, 0.0, [3976, 2603], np.float32)
        q = W.box_states(Q_WAVE, 2, 52, 0.0, [3976, 2603], np.float32)
        out[label] = time_knn(m.lp_space(2, 2, m.F32), pts, q, strat)
    # SURVEY.md 8(d) sweep: the C5 space at planner-realistic set sizes (2^20 is the headline line), strategy AUTO
    q = W.se3_states(Q_WAVE, W.QUERY_SEED)
    for lg in (10, 14, 17):
        out[f"knn_se3_2^{lg}"] = time_knn(m.se3_space(SO3_W, L2_W), W.se3_states(1 << lg, W.TREE_SEED), q, m.KNN_AUTO)
    # ... and the N-link arm's space (L1 over [-pi, pi)^N, float64 as the reference's main builds it), 2^17 points
    for n_links in (8, 16, 32):
        pts = W.box_states(1 << 17, n_links, 61, -np.pi, np.pi)
        q = W.box_states(Q_WAVE, n_links, 62, -np.pi, np.pi)
        out[f"knn_l1_{n_links}d_128k"] = time_knn(m.lp_space(n_links, 1, m.F64), pts, q, m.KNN_AUTO, np.float64)
    # mesh edges steered to a range of 5 % / 20 % of the scenario's diagonal (|max - min| + 50 pi / 2, the quantity the
    # reference's step-size rule uses, se3_rigid_body_scenario.hpp:333), half of it spent on translation, half on rotation
    robot, env, vmin, vmax = W.alpha_puzzle_like(env_tris_target=4000, robot_tris_target=1000)
    sp3 = m.se3_space(SO3_W, L2_W)
    mesh = m.Scenario.mesh_pair(ctx, robot, env, sp3, W.se3_step_size(vmin, vmax, SO3_W))
    diag = float(np.linalg.norm(np.asarray(vmax, dtype=np.float64) - np.asarray(vmin, dtype=np.float64))) + SO3_W * np.pi / 2
    for pct in (5, 20):
        rho = diag * pct / 100.0
        a, b = W.se3_edges(E_WAVE, W.EDGE_SEED + pct, MESH_LO, MESH_HI, rho / 2, rho / 2 / SO3_W)
        r = time_link(mesh, a, b)
        r["states_per_s"] = mesh.last_stats()["states"] / (r["ms"] * 1e-3)
        r["range"] = rho
        out[f"mesh_edges_range{pct}pct"] = r
    mesh.close()
    # device-resident PRRT: tree, sampling and every stage of the loop on the GPU; wall clock
    # around whole waves, two words read back per wave
    import time

    free = np.argwhere(occ == 0)
    start = free[len(free) // 7][::-1].astype(np.float64)
    # wall clock on a shared box: the faster of two identical runs (the tree is the same both times)
    for attempt in range(2):
        pl = m.DevicePRRT(grid, m.lp_space(2, 2, m.F64), [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1], range=200.0, seed=17,
                          capacity=1 << 20, max_wave=16384)
        pl.add_start(start)
        pl.wave(16384)
        ctx.sync()
        t0, n0 = time.perf_counter(), pl.size
        while pl.size < 500_000:
            pl.wave(16384)
        dt = time.perf_counter() - t0
        if attempt == 0 or (pl.size - n0) / dt > out["device_prrt_grid"]["nodes_per_s"]:
            out["device_prrt_grid"] = {"nodes_per_s": (pl.size - n0) / dt, "samples_per_s": (pl.samples_drawn - 16384) / dt, "nodes": pl.size, "s": dt,
                                       "timing": "wall clock, faster of two identical runs"}
        pl.close()
    # device-resident PRRT* (BASELINE configs[1]: PRRT* on the occupancy grid): same map, start and range
    goal = free[-len(free) // 9][::-1].astype(np.float64)
    warm = m.DevicePRRTStar(grid, m.lp_space(2, 2, m.F64), [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1], range=200.0, goal=goal, goal_radius=12.0,
                            seed=3, capacity=1 << 16, max_wave=8192)  # untimed: loads every kernel of the wave (wall-clock timing below)
    warm.add_start(start)
    for _ in range(6):
        warm.wave(8192)
    warm.close()
    for attempt in range(2):
        ps = m.DevicePRRTStar(grid, m.lp_space(2, 2, m.F64), [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1], range=200.0, goal=goal, goal_radius=12.0,
                              seed=17, capacity=1 << 20, max_wave=8192)
        ps.add_start(start)
        ctx.sync()
        t0 = time.perf_counter()
        first_solution = None
        while ps.size < 200_000:
            ps.wave(8192)
            if first_solution is None and ps.solved():
                first_solution = (time.perf_counter() - t0, ps.size)
        dt = time.perf_counter() - t0
        if attempt == 0 or ps.size / dt > out["device_prrtstar_grid"]["nodes_per_s"]:
            out["device_prrtstar_grid"] = {"nodes_per_s": ps.size / dt, "nodes": ps.size, "s": dt, "rewires": ps.rewires, "solved": ps.solved(),
                                           "first_solution_s": first_solution[0] if first_solution else None,
                                           "first_solution_nodes": first_solution[1] if first_solution else None,
                                           "solution_cost": ps.solution_cost() if ps.solved() else None,
                                           "timing": "wall clock, faster of two identical runs"}
        ps.close()
    # time to the FIRST solution (13.4 ms / 52,804 nodes with 8,192-sample waves throughout against 1.3 ms / 816 nodes
    # for the reference's PRRT* on 16 host threads): a tree grows by at most one range per wave, so early waves are kept small
    # and held at 512 samples until a solution exists -- the wave ramp of the C++ DevicePRRT / DevicePRRTStar (planner.hpp impl::WaveRamp)
    for name, cls in (("device_prrt_grid_first_solution", m.DevicePRRT), ("device_prrtstar_grid_first_solution", m.DevicePRRTStar)):
        best = None
        for attempt in range(3):
            pl = cls(grid, m.lp_space(2, 2, m.F64), [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1], range=200.0, goal=goal, goal_radius=12.0,
                     goal_bias=0.01, seed=17, capacity=1 << 18, max_wave=8192)
            pl.add_start(start)
            ctx.sync()
            t0, w, waves = time.perf_counter(), 64, 0
            while not pl.solved() and pl.size < 200_000:  # impl::WaveRamp of planner.hpp: 64, 128, 256, then 512 held for 16 waves, then doubled every fourth
                pl.wave(w)
                waves += 1
                if w < 512 or (waves >= 16 and (waves - 16) % 4 == 3):
                    w = min(2 * w, 8192)
            dt = time.perf_counter() - t0
            if best is None or dt < best["first_solution_s"]:
                best = {"first_solution_s": dt, "first_solution_nodes": pl.size, "waves": waves, "solved": pl.solved(), "ramp": "64, 128, 256 samples, then 512 per wave until the first solution (doubled every fourth wave after 16)",
                        "timing": "wall clock, fastest of three identical runs"}
            pl.close()
        out[name] = best
    # device-resident PPRM (BASELINE configs[3]: PPRM for the N-link arm): roadmap, components and every stage on the GPU
    for n_links in (8, 16):
        lengths, radius, circles = W.link_arm_scene(n_links)
        spn = m.lp_space(n_links, 1, m.F64)
        arm = m.Scenario.link_arm(ctx, lengths, radius, circles, m.F64)
        cand = W.box_states(512, n_links, 3, -np.pi, np.pi)
        ok = arm.valid(cand) != 0
        for attempt in range(2):  # wall clock on a shared box: the faster of two identical runs (the roadmap is the same both times)
            pp = m.DevicePPRM(arm, spn, -np.pi, np.pi, seed=23, capacity=1 << 18, max_wave=4096)
            pp.add_start(cand[ok][0])
            pp.add_goal(cand[ok][1])
            pp.wave(4096)
            ctx.sync()
            t0, n0 = time.perf_counter(), pp.size
            while pp.size < 150_000:
                pp.wave(4096)
            dt = time.perf_counter() - t0
            if attempt == 0 or (pp.size - n0) / dt > out[f"device_pprm_arm{n_links}"]["nodes_per_s"]:
                ei = pp.graph(n0, pp.size - n0)[1]
                out[f"device_pprm_arm{n_links}"] = {"nodes_per_s": (pp.size - n0) / dt, "edges_checked_per_s": float((pp.size - n0) * pp.row_stride) / dt,
                                                    "roadmap_edges": int((ei != m.NO_INDEX).sum()), "nodes": pp.size, "solved": pp.solved(), "s": dt,
                                                    "timing": "wall clock, faster of two identical runs"}
            pp.close()
        ref_arm = reference_planner_cpu_arm(lengths, radius, circles, cand[ok][0], cand[ok][1])
        if ref_arm:
            out[f"reference_planner_cpu_arm{n_links}"] = ref_arm
        if n_links == 8:  # PPRM-IRS: the same waves with the spanner's bounded search per new node on the device.
            # In 8 dimensions a ball of 5 x the neighbour distance holds most of the roadmap, so a search that KEEPS an edge
            # labels most nodes (for the reference as for us): small roadmap, short waves
            pp = m.DevicePPRM(arm, spn, -np.pi, np.pi, seed=23, capacity=1 << 15, max_wave=512, spanner_stretch=5.0)
            pp.add_start(cand[ok][0])
            pp.add_goal(cand[ok][1])
            pp.wave(512)
            ctx.sync()
            t0, n0 = time.perf_counter(), pp.size
            while pp.size < 8_000 and time.perf_counter() - t0 < 20.0:
                pp.wave(512)
            dt = time.perf_counter() - t0
            ei = pp.graph(n0, pp.size - n0)[1]
            out["device_pprm_irs_arm8"] = {"nodes_per_s": (pp.size - n0) / dt, "nodes": pp.size, "sparse_edges": int((ei != m.NO_INDEX).sum()), "solved": pp.solved(), "s": dt,
                                           "stretch": 5.0, "timing": "wall clock, one run"}
            pp.close()
            ref_irs = reference_planner_cpu_arm(lengths, radius, circles, cand[ok][0], cand[ok][1], nodes=8_000, time_ms=20000, algo="pprmirs")
            if ref_irs:
                out["reference_planner_cpu_irs_arm8"] = ref_irs
    # PPRM-IRS where the spanner's searches are local: the PNG-size occupancy grid (planar L2)
    pp = m.DevicePPRM(grid, m.lp_space(2, 2, m.F64), [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1], seed=23, capacity=1 << 18, max_wave=4096, spanner_stretch=5.0)
    pp.add_start(start)
    pp.add_goal(goal)
    pp.wave(4096)
    ctx.sync()
    t0, n0 = time.perf_counter(), pp.size
    while pp.size < 100_000:
        pp.wave(4096)
    dt = time.perf_counter() - t0
    ei = pp.graph(n0, pp.size - n0)[1]
    out["device_pprm_irs_grid"] = {"nodes_per_s": (pp.size - n0) / dt, "nodes": pp.size, "sparse_edges_per_node": float((ei != m.NO_INDEX).sum()) / (pp.size - n0),
                                   "solved": pp.solved(), "s": dt, "stretch": 5.0, "timing": "wall clock, one run"}
    pp.close()
    pp = m.DevicePPRM(grid, m.lp_space(2, 2, m.F64), [0, 0], [occ.shape[1] - 1, occ.shape[0] - 1], seed=23, capacity=1 << 18, max_wave=4096)
    pp.add_start(start)
    pp.add_goal(goal)
    pp.wave(4096)
    ctx.sync()
    t0, n0 = time.perf_counter(), pp.size
    while pp.size < 100_000:
        pp.wave(4096)
    dt = time.perf_counter() - t0
    ei = pp.graph(n0, pp.size - n0)[1]
    out["device_pprm_grid"] = {"nodes_per_s": (pp.size - n0) / dt, "nodes": pp.size, "edges_per_node": float((ei != m.NO_INDEX).sum()) / (pp.size - n0), "solved": pp.solved(),
                               "s": dt, "timing": "wall clock, one run"}
    pp.close()
    # BASELINE configs[3] as a PLANNING problem (the scene above connects start and goal almost directly): rings of
    # circles with narrow gaps, W.link_arm_passage_scene -- time and roadmap size to the first solution, ours and the reference's
    for n_links in (8, 16):
        lengths, radius, circles, a_start, a_goal = W.link_arm_passage_scene(n_links)
        arm = m.Scenario.link_arm(ctx, lengths, radius, circles, m.F64)
        ea, eb = W.arm_edges(E_WAVE, n_links, 41, 0.5)
        edges = time_link(arm, ea, eb)
        best = None
        for attempt in range(2 if n_links == 8 else 1):
            pp = m.DevicePPRM(arm, m.lp_space(n_links, 1, m.F64), -np.pi, np.pi, seed=23, capacity=1 << 19, max_wave=4096)
            pp.add_start(a_start)
            pp.add_goal(a_goal)
            ctx.sync()
            t0, w = time.perf_counter(), 256
            while not pp.solved() and pp.size < 400_000 and time.perf_counter() - t0 < 10.0:
                pp.wave(w)
                w = min(2 * w, 4096)
            dt = time.perf_counter() - t0
            path_ok = None
            if pp.solved():  # the roadmap's own answer: every edge of the shortest path passes the edge check again
                path = pp.solution()
                path_ok = bool(len(path) >= 2 and (arm.link(path[:-1], path[1:]) == 1).all() and np.array_equal(path[0], a_start) and np.array_equal(path[-1], a_goal))
            if best is None or dt < best["first_solution_s"]:
                best = {"solved": pp.solved(), "first_solution_s": dt, "first_solution_nodes": pp.size, "nodes_per_s": pp.size / dt, "path_revalidated": path_ok,
                        "edge_wave": {"edges_per_s": edges["edges_per_s"], "valid_fraction": edges["valid_fraction"]},
                        "timing": "wall clock, faster of two identical runs; waves of 256 samples doubled up to 4,096"}
            pp.close()
        out[f"device_pprm_arm{n_links}_passage"] = best
        ref_arm = reference_planner_cpu_arm(lengths, radius, circles, a_start, a_goal, nodes=200_000, time_ms=10000)
        if ref_arm:
            out[f"reference_planner_cpu_arm{n_links}_passage"] = ref_arm
            # the reference's roadmap connects with FEWER nodes the fewer workers build it (measured: 1 thread 160 - 900 nodes,
            # 8 threads ~10 K, 16 threads none within 72 K): its single-threaded run is the fair time-to-solution figure
            one = reference_planner_cpu_arm(lengths, radius, circles, a_start, a_goal, nodes=200_000, time_ms=10000, threads=1)
            if one:
                out[f"reference_planner_cpu_arm{n_links}_passage_1thread"] = one
        arm.close()
    # same map, same start, same range: the reference's own multi-threaded PRRT / PRRT* on the host cores
    ref = reference_planner_cpu(occ, start, goal, 12.0, 200.0)
    if ref:
        out["reference_planner_cpu_grid"] = ref
    return out


def reference_planner_cpu(occ, start, goal, goal_radius, prrt_range, nodes=200_000, time_ms=8000):
    """The reference's OWN multi-threaded planner classes (src/mpt PRRT / PRRT*, OpenMP worker pool, one worker per host
    thread) on the occupancy-grid scenario, timed on this box's host cores: oracle/_ref/ref_planner_bench, compiled where
    /root/reference exists (oracle/Makefile, target ref) and shipped with the snapshot.  Nigh is absent, so the nearest-
    neighbour structure is a stand-in concurrent kd-tree (oracle/shim/nigh/nigh_kdtree.hpp); everything else -- worker
    loop, sampling, steering, rewiring, PNG2dScenario::valid / link -- is the reference's code.  Returns {} when the
    program was not shipped."""
    import json
    import subprocess
    import tempfile

    prog = ROOT / "oracle" / "_ref" / "ref_planner_bench"
    if not prog.exists():
        return {}
    out = {}
    with tempfile.NamedTemporaryFile(suffix=".pgm") as f:
        f.write(b"P5\n%d %d\n255\n" % (occ.shape[1], occ.shape[0]) + (np.asarray(occ) != 0).astype(np.uint8).tobytes())
        f.flush()
        env = dict(os.environ)
        env["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1
        for algo in ("prrt", "prrtstar"):
            cmd = [str(prog), "--map", f.name, "--start", str(start[0]), str(start[1]), "--goal", str(goal[0]), str(goal[1]), "--goal-radius",
                   str(goal_radius), "--range", str(prrt_range), "--algo", algo, "--nodes", str(nodes), "--time-ms", str(time_ms), "--seed", "17"]
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=time_ms / 1e3 + 60, env=env)
                out[algo] = json.loads(r.stdout.strip().splitlines()[-1])
            except Exception as e:  # a baseline, never fatal
                out[algo] = {"error": str(e)[:200]}
    return out


def reference_planner_cpu_arm(lengths, radius, circles, start, goal, nodes=30_000, time_ms=6000, threads=None, algo="pprm"):
    """The reference's own multi-threaded PPRM (BASELINE configs[3]) on the reference's LinkManipulatorScenario<double, N>
    (demo/link_manipulator_scenario.hpp), N = 8 or 16, on this box's host cores.  Same program and caveats as
    reference_planner_cpu."""
    import json
    import subprocess
    import tempfile

    prog = ROOT / "oracle" / "_ref" / "ref_planner_bench"
    if not prog.exists() or len(lengths) not in (8, 16):
        return {}
    with tempfile.NamedTemporaryFile("w", suffix=".txt") as f:
        f.write(f"{len(lengths)} {radius!r}\n" + " ".join(repr(float(x)) for x in lengths) + f"\n{len(circles)}\n")
        for c in circles:
            f.write(" ".join(repr(float(x)) for x in c) + "\n")
        f.write(" ".join(repr(float(x)) for x in start) + "\n" + " ".join(repr(float(x)) for x in goal) + "\n")
        f.flush()
        env = dict(os.environ)
        env["OMP_NUM_THREADS"] = str(threads or os.cpu_count() or 1)
        try:
            r = subprocess.run([str(prog), "--arm", f.name, "--algo", algo, "--nodes", str(nodes), "--time-ms", str(time_ms), "--seed", "23"]
                               + (["--threads", str(threads)] if threads else []),
                               capture_output=True, text=True, timeout=time_ms / 1e3 + 60, env=env)
            return json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:
            return {"error": str(e)[:200]}


def fp32_probe(ctx):
    """FP32 CUDA-core yardstick: the library's register-only FFMA kernel (mptg_probe_fp32_tflops), as SURVEY.md 8(d) asks."""
    return ctx.probe_fp32_tflops()


_NAO_SAMPLE = {}  # secondary() -> cpu_baseline(): 8,192 Nao-cup edges (float64) with the device's decisions


def cpu_baseline(args, sp, tree, queries, robot, env, step, ea, eb):
    from tests import oracle_binding

    orc = oracle_binding.load()
    orc.set_threads(os.cpu_count() or 1)
    otree = orc.tree(sp, tree)
    omesh = orc.mesh_pair(robot, env, sp, step)
    qs, es = args.cpu_queries, args.cpu_edges
    otree.knn(queries[:256], K_NN)
    t0 = time.perf_counter()
    otree.knn(queries[:qs], K_NN)
    t1 = time.perf_counter()
    omesh.link(ea[:es], eb[:es])
    t2 = time.perf_counter()
    out = {"value": qs / (t1 - t0), "unit": "queries/s", "edges_per_s": es / (t2 - t1), "cores": orc.threads, "kind": "port",
           "sample": f"first {qs} of {Q_WAVE} queries (k={K_NN}, N={N_TREE}, box-tree search) and first {es} of {E_WAVE} edges, "
                     f"OpenMP over all host threads"}
    if _NAO_SAMPLE:  # the Nao-cup edges of secondary.nao_cup_edges_f64 on the oracle (oracle/oracle_nao.hpp), same threads
        on = orc.nao_cup(8)
        t3 = time.perf_counter()
        want = on.link(_NAO_SAMPLE["a"], _NAO_SAMPLE["b"])
        t4 = time.perf_counter()
        out["nao_cup_edges"] = {"edges_per_s": _NAO_SAMPLE["a"].shape[0] / (t4 - t3), "sample": "8,192 of the 65,536 float64 edges",
                                "decisions_equal_device": bool(np.array_equal(want, _NAO_SAMPLE["ok"]))}
    return out


def bench_replicated_tree(args, ctx, sp, tree, dev, stream, world, rank):
    """The zero-collective alternative when the tree fits one HBM (28 MB here): every rank holds the whole tree and
    answers its OWN wave.  Trivially linear; reported next to the sharded figure, which is the north_star's split."""
    import torch
    import torch.distributed as dist

    import mpt_b200 as m
    from mpt_b200 import workloads as W

    q = W.se3_states(Q_WAVE, W.QUERY_SEED + 1000 * rank)
    full = m.Nearest(ctx, sp, N_TREE)
    full.insert(tree)
    full.build_index()
    d_q = torch.from_numpy(q).to(dev)
    loc_i = torch.empty((Q_WAVE, K_NN), dtype=torch.int32, device=dev)
    loc_d = torch.empty((Q_WAVE, K_NN), dtype=torch.float32, device=dev)
    times = []
    for it in range(args.warmup + args.steps):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            full.nearest_dev(d_q.data_ptr(), Q_WAVE, K_NN, -1.0, loc_i.data_ptr(), loc_d.data_ptr())
            e1.record(stream)
        torch.cuda.synchronize()
        if it >= args.warmup:
            times.append(e0.elapsed_time(e1))
    full.close()
    t = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    return {"queries_per_s": world * Q_WAVE / (ms * 1e-3), "ms_per_wave": ms, "scaling": "weak", "collective": "none",
            "tree_nodes_per_gpu": N_TREE}


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    # Libraries (NCCL prints its version banner) must not pollute stdout: everything but the JSON line goes to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-queries", type=int, default=8192, help="CPU-baseline sample: queries per step")
    ap.add_argument("--cpu-edges", type=int, default=4096, help="CPU-baseline sample: edges per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the grid / link-arm / planar-kNN side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
