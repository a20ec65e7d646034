"""N>1 path on CPU: world_size-2 gloo processes exercise the sharding protocol of mpt_b200/sharding.py
(unit slices, round-robin tree dealing with global indices, all-gather + merge by (distance, index)),
with the CPU oracle standing in for the per-rank device search.  The GPU equivalent is
tests/test_gpu_parity.py::test_knn_merge_sharded_equals_single and bench.py --gpus N."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from mpt_b200 import sharding  # noqa: E402


def test_unit_slices_cover_exactly():
    for n in (0, 1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(n, dtype=int)
            for r in range(world):
                seen[sharding.unit_slice(n, world, r)] += 1
            assert (seen == 1).all()
            sizes = [sharding.unit_slice(n, world, r).stop - sharding.unit_slice(n, world, r).start for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_round_robin_dealing_and_index_map():
    pts = np.arange(23 * 2, dtype=np.float32).reshape(23, 2)
    for world in (1, 2, 4, 8):
        for r in range(world):
            shard = sharding.tree_shard(pts, world, r)
            mul, add = sharding.index_map(world, r)
            for local in range(shard.shape[0]):
                g = local * mul + add
                assert np.array_equal(pts[g], shard[local]) and sharding.owner_of(g, world) == (r, local)


def test_merge_topk_host_ties_by_index():
    idx = np.array([[[5, 9, 0xFFFFFFFF]], [[2, 7, 11]]], dtype=np.uint32)  # [G=2, Q=1, k=3]
    d = np.array([[[1.0, 2.0, np.inf]], [[1.0, 2.0, 3.0]]], dtype=np.float32)
    oi, od, cnt = sharding.merge_topk_host(idx, d, 4)
    assert list(oi[0]) == [2, 5, 7, 9] and list(od[0]) == [1.0, 1.0, 2.0, 2.0] and cnt[0] == 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mpt_b200 as m
    from mpt_b200 import workloads as W
    from tests import oracle_binding

    orc = oracle_binding.load()
    orc.set_threads(2)
    sp = m.se3_space(50, 1)
    pts = W.se3_states(6000, 1)   # identical on every rank (same seed)
    q = W.se3_states(96, 2)
    k = 16
    shard = sharding.tree_shard(pts, world, rank)
    mul, add = sharding.index_map(world, rank)

    def local_topk(queries, kk):
        idx, d, _ = orc.knn(sp, shard, queries, kk)
        gidx = np.where(idx == m.NO_INDEX, idx, idx * np.uint32(mul) + np.uint32(add)).astype(np.uint32)
        return gidx, d

    def all_gather(x):
        t = torch.from_numpy(np.ascontiguousarray(x).view(np.int32 if x.dtype == np.uint32 else x.dtype))
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return np.stack([p.numpy().view(x.dtype) for p in parts])

    idx, d, cnt = sharding.sharded_knn(local_topk, sharding.merge_topk_host, all_gather, q, k)
    wi, wd, wc = orc.knn(sp, pts, q, k)
    ok_tree = np.array_equal(idx, wi) and np.array_equal(d, wd) and np.array_equal(cnt, wc)

    # spatially sharded tree, the protocol of csrc/comm.cu (root bounds -> home search -> bounded search -> exchange -> merge)
    cells = sharding.spatial_cells(pts, (4, 5, 6), -100.0, 100.0, world)
    mine_ids = np.nonzero(cells == rank)[0].astype(np.uint32)
    mine_pts = pts[mine_ids]
    tlo, thi = mine_pts[:, 4:].min(0).astype(np.float64), mine_pts[:, 4:].max(0).astype(np.float64)

    def root_bound(queries):  # translation distance to the shard's bounding box: a lower bound of the SE(3) distance
        e = np.maximum(np.maximum(tlo - queries[:, 4:], queries[:, 4:] - thi), 0.0)
        return np.sqrt((e.astype(np.float64) ** 2).sum(1)) * (1 - 1e-6)

    searched = [0]

    def local_capped(queries, kk, cap):
        idx = np.full((queries.shape[0], kk), m.NO_INDEX, dtype=np.uint32)
        d = np.full((queries.shape[0], kk), np.inf, dtype=np.float32)
        for r in np.unique(cap[cap >= 0]):
            sel = np.nonzero(cap == r)[0]
            li, ld, _ = orc.knn(sp, mine_pts, queries[sel], kk, float(r) if np.isfinite(r) else -1.0)
            idx[sel] = np.where(li == m.NO_INDEX, li, mine_ids[np.minimum(li, len(mine_ids) - 1)])
            d[sel] = ld
        searched[0] += int((cap >= 0).sum())
        return idx, d

    def all_reduce_min(x):
        t = torch.from_numpy(np.ascontiguousarray(x))
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return t.numpy()

    def exchange(idx, d):
        sl_own = sharding.unit_slice(q.shape[0], world, rank)
        gi, gd = all_gather(idx), all_gather(d)      # gloo stand-in for the grouped send/recv: keep this rank's slice
        return gi[:, sl_own], gd[:, sl_own]

    si, sd, sc = sharding.sharded_knn_spatial(root_bound, local_capped, sharding.merge_topk_host, all_gather, all_reduce_min, exchange, q, k, rank, world)
    sl_own = sharding.unit_slice(q.shape[0], world, rank)
    ok_tree = ok_tree and np.array_equal(si, wi[sl_own]) and np.array_equal(sd, wd[sl_own]) and np.array_equal(sc, wc[sl_own])
    ok_tree = ok_tree and searched[0] < 2 * q.shape[0]   # far shards are not searched: fewer than (home + all) searches

    # unit-sharded validity: every rank checks its slice, rank 0 gathers the bytes
    occ = W.synthetic_grid(300, 200, seed=3)
    a, b = W.grid_edges(1000, 300, 200, 5, 30.0)
    sl = sharding.unit_slice(1000, world, rank)
    mine = orc.grid(occ).link(a[sl], b[sl])
    sizes = [sharding.unit_slice(1000, world, r).stop - sharding.unit_slice(1000, world, r).start for r in range(world)]
    parts = [torch.empty(s, dtype=torch.uint8) for s in sizes]
    dist.all_gather(parts, torch.from_numpy(mine)) if len(set(sizes)) == 1 else None
    ok_edges = True
    if len(set(sizes)) == 1:
        ok_edges = np.array_equal(np.concatenate([p.numpy() for p in parts]), orc.grid(occ).link(a, b))
    Path(out_dir, f"rank{rank}.txt").write_text(f"{int(ok_tree)} {int(ok_edges)}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_tree_and_edges_gloo(tmp_path, world):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"rank{r}.txt").read_text() == "1 1"
