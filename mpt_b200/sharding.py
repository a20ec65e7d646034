"""Multi-GPU layout of the hot path: one process per GPU, torch.distributed for
the plumbing.  Pure host logic -- the same functions are driven by CUDA callables in bench.py and by
CPU callables in the gloo tests.

  * state / edge validity and query waves shard by independent units: rank r owns a contiguous slice,
    geometry and (when it fits) the tree are replicated, no data-path collective.
  * a tree too large for one HBM is dealt round-robin: global node g lives on rank g % G at local
    slot g // G (balanced under append-only inserts); every rank answers the whole query wave on
    its shard with GLOBAL indices (index map mul = G, add = r), the [Q, k] candidate lists are
    all-gathered and merged by the kNN total order (distance, global index).
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np


def unit_slice(n: int, world: int, rank: int) -> slice:
    """Contiguous slice of n independent units (queries, edges, states) owned by `rank`."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def tree_shard(points: np.ndarray, world: int, rank: int) -> np.ndarray:
    """Rows of `points` stored on `rank` under round-robin dealing."""
    return np.ascontiguousarray(points[rank::world])


def index_map(world: int, rank: int) -> Tuple[int, int]:
    """(mul, add) such that global index = local index * mul + add."""
    return world, rank


def owner_of(global_index: int, world: int) -> Tuple[int, int]:
    """(rank, local slot) of a global node index."""
    return global_index % world, global_index // world


def spatial_cells(states: np.ndarray, coords, lo, hi, world: int) -> np.ndarray:
    """Rank of every state under a FIXED spatial partition of the box [lo, hi) over the given coordinates: the box is
    halved along coords[0], coords[1], ... cyclically until there are `world` cells (world a power of two; otherwise
    the last split is uneven in cell count, never in correctness).  Fixed by the bounds, not by the data, so a planner
    can route a new node to its rank without knowing the other nodes.  For SE(3) states use the translation
    coordinates (4, 5, 6): the rotation part of a uniform sample gives nothing to prune on at the top level."""
    s = np.asarray(states)
    cell = np.zeros(s.shape[0], dtype=np.int64)
    lo = np.broadcast_to(np.asarray(lo, dtype=np.float64), (len(coords),)).copy()
    hi = np.broadcast_to(np.asarray(hi, dtype=np.float64), (len(coords),)).copy()
    cells, depth = 1, 0
    # bit d of the cell index = which half along coords[d % len] at the (d // len)-th halving of that coordinate
    while cells < world:
        j = depth % len(coords)
        level = depth // len(coords)
        width = (hi[j] - lo[j]) / (2 ** level)
        rel = (s[:, coords[j]].astype(np.float64) - lo[j]) / width
        bit = (np.floor(rel * 2).astype(np.int64)) & 1
        cell |= np.clip(bit, 0, 1) << depth
        cells *= 2
        depth += 1
    return (cell % world).astype(np.int32)


def sharded_knn_spatial(root_bound: Callable, local_topk: Callable, merge: Callable, all_gather: Callable, all_reduce_min: Callable,
                        exchange: Callable, queries, k: int, rank: int, world: int):
    """Host mirror of the protocol in csrc/comm.cu (same five steps), driven by callables so that the gloo tests can run
    it on CPU:
      root_bound(queries) -> [Q] lower bound of every query to this rank's shard
      local_topk(queries, k, cap) -> (idx [Q,k], dist [Q,k]) with global indices; rows with cap < 0 are skipped
         (NO_INDEX / inf), others are searched within radius cap
      all_gather(x) -> [G, ...]; all_reduce_min(x) -> elementwise min over ranks
      exchange(idx, dist) -> (idx_parts [G, n_own, k], dist_parts [G, n_own, k]) rows of this rank's slice from every rank
      merge(idx_parts, dist_parts, k) -> (idx, dist, count) of the slice."""
    lb_mine = np.asarray(root_bound(queries), dtype=np.float64)
    lb_all = np.asarray(all_gather(lb_mine))                      # 1
    home = lb_all.argmin(axis=0)                                  # lowest rank on ties
    cap = np.where(home == rank, np.inf, -1.0)
    idx, dist = local_topk(queries, k, cap)                       # 2
    bound = np.asarray(all_reduce_min(dist[:, k - 1].astype(np.float64)))
    cap2 = np.where((home != rank) & (lb_mine <= bound), bound, -1.0)
    idx2, dist2 = local_topk(queries, k, cap2)                    # 3
    sel = cap2 >= 0
    idx[sel], dist[sel] = idx2[sel], dist2[sel]
    ip, dp = exchange(idx, dist)                                  # 4
    return merge(ip, dp, k)                                       # 5


def sharded_knn(local_topk: Callable, merge: Callable, all_gather: Callable, queries, k: int):
    """Tree-sharded kNN step.
    local_topk(queries, k) -> (idx [Q,k] global indices, dist [Q,k]) on this rank's shard
    all_gather(x) -> stacked [G, ...] over ranks
    merge(idx_parts [G,Q,k], dist_parts [G,Q,k], k) -> (idx [Q,k], dist [Q,k], count [Q])
    """
    idx, dist = local_topk(queries, k)
    return merge(all_gather(idx), all_gather(dist), k)


def merge_topk_host(idx_parts: np.ndarray, dist_parts: np.ndarray, k: int, no_index: int = 0xFFFFFFFF):
    """Reference semantics of the merge on the host (numpy): k best by (distance, index) per query.
    The product merges on the device (mptg_knn_merge_dev); this is what the gloo tests and the
    sharding unit tests compare against."""
    G, Q, kk = idx_parts.shape
    idx = np.transpose(idx_parts, (1, 0, 2)).reshape(Q, G * kk)
    dist = np.transpose(dist_parts, (1, 0, 2)).reshape(Q, G * kk)
    out_i = np.full((Q, k), no_index, dtype=np.uint32)
    out_d = np.full((Q, k), np.inf, dtype=dist.dtype)
    cnt = np.zeros(Q, dtype=np.uint32)
    for q in range(Q):
        real = idx[q] != no_index
        order = np.lexsort((idx[q][real], dist[q][real]))[:k]
        n = order.size
        out_i[q, :n] = idx[q][real][order]
        out_d[q, :n] = dist[q][real][order]
        cnt[q] = n
    return out_i, out_d, cnt
