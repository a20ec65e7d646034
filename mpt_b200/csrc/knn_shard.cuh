// knn_shard.cuh -- what the sharded search (comm.cu) needs from a kNN store (knn.cu).
#pragma once

#include "common.cuh"

namespace mptg {
mptg_ctx* knnShardCtx(mptg_knn* knn);
int knnShardScalar(const mptg_knn* knn);
int knnShardScalars(const mptg_knn* knn);
const mptg_space_desc* knnShardSpace(const mptg_knn* knn);
uint32_t knnShardSize(const mptg_knn* knn);
uint64_t knnShardBuilds(const mptg_knn* knn);
// index everything stored; boxes of the top node's children (device, [2 D][32] scalars: lo rows, hi rows) and their number
int knnShardIndexAll(mptg_knn* knn, const void** topBoxDev, uint32_t* nTop);
// one box per shard [world][2 D] from the synchronised top boxes peerBox [world][2 D][32] / peerN [world]
int knnShardUnion(mptg_knn* knn, const void* peerBox, const uint32_t* peerN, int world, void* shardBox);
// lbMine[q] = lower bound of query q to this rank's shard, home[q] = the shard with the smallest bound over the shard
// boxes, cap[q] = +inf where home == rank, -1 elsewhere
int knnShardRootAll(mptg_knn* knn, const void* shardBox, const uint32_t* peerN, int world, int rank, const void* queriesDev, uint32_t Q, float* lbMine,
                    uint8_t* home, void* cap);
// search with a per-query radius cap (< 0: skip the query, its output row is left alone; large waves are ordered over the
// queries that are searched only).
int knnShardQuery(mptg_knn* knn, const void* queriesDev, uint32_t Q, uint32_t k, double radius, const void* qcapDev, uint32_t* idxOut,
                  void* distOut);
}  // namespace mptg
