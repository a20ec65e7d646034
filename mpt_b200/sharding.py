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
