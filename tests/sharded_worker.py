"""One rank of the tree-sharded kNN test / demo: python tests/sharded_worker.py <rank> <world> <dir> [n] [Q] [k]
The NCCL unique id travels through a file in <dir> (rank 0 writes it); results of this rank's slice are compared with
the CPU oracle over ALL points and the verdict is written to <dir>/rank<r>.txt."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    rank, world, d = int(sys.argv[1]), int(sys.argv[2]), Path(sys.argv[3])
    n = int(sys.argv[4]) if len(sys.argv) > 4 else 60_000
    Q = int(sys.argv[5]) if len(sys.argv) > 5 else 1000
    k = int(sys.argv[6]) if len(sys.argv) > 6 else 16
    import mpt_b200 as m
    from mpt_b200 import sharding
    from mpt_b200 import workloads as W
    from tests import oracle_binding

    idf = d / "nccl_id.bin"
    if rank == 0:
        uid = m.Comm.unique_id()
        tmp = d / "nccl_id.tmp"
        tmp.write_bytes(uid)
        tmp.rename(idf)
    else:
        t0 = time.time()
        while not idf.exists():
            if time.time() - t0 > 120:
                raise RuntimeError("no NCCL id from rank 0")
            time.sleep(0.05)
        uid = idf.read_bytes()
    ctx = m.Context(rank)  # one GPU per rank
    comm = m.Comm(ctx, uid, rank, world)
    ok = True
    msgs = []
    orc = oracle_binding.load()
    for scalar, dt in ((m.F32, np.float32), (m.F64, np.float64)):
        sp = m.se3_space(50, 1, scalar)
        pts = W.se3_states(n, 1, dtype=dt)
        pts[n // 2] = pts[7]                      # duplicates across shards: ties broken by global index
        pts[n // 2 + 1, 4:] = -pts[n // 2 + 1, 4:]
        q = W.se3_states(Q, 2, dtype=dt)
        q[0] = pts[7]
        cells = sharding.spatial_cells(pts, (4, 5, 6), -100.0, 100.0, world)
        ids = np.nonzero(cells == rank)[0].astype(np.uint32)
        shard = m.Nearest(ctx, sp, max(len(ids), 64), m.KNN_BVH)
        # two batches: the second one arrives after a search (tail of the shard: re-indexed by the sharded search)
        half = len(ids) // 2
        shard.insert_ids(pts[ids[:half]], ids[:half])
        comm.sync(shard)
        comm.nearest(shard, q[:64], 4)
        shard.insert_ids(pts[ids[half:]], ids[half:])
        try:
            comm.nearest(shard, q[:64], 4)   # not synchronised since the insert: refused (no collective has started)
            ok = False
        except m.MptgError as e:
            ok = ok and e.code == -1
        comm.sync(shard)
        first, count = comm.slice(Q)
        for kk, radius in ((k, -1.0), (1, -1.0), (k, 30.0), (48, -1.0)):
            gi, gd, gc = comm.nearest(shard, q, kk, radius)
            wi, wd, wc = orc.knn(sp, pts, q, kk, radius)
            same = np.array_equal(gi, wi[first:first + count]) and np.array_equal(gd, wd[first:first + count]) and np.array_equal(gc, wc[first:first + count])
            ok = ok and same
            msgs.append(f"{'f32' if scalar == m.F32 else 'f64'} k={kk} r={radius}: {'ok' if same else 'MISMATCH'}")
        shard.close()
    comm.close()
    ctx.close()
    (d / f"rank{rank}.txt").write_text(("1 " if ok else "0 ") + "; ".join(msgs))


if __name__ == "__main__":
    main()
