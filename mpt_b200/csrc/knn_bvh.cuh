// knn_bvh.cuh -- wide bounding-box tree for large point sets (interface; see knn_bvh_impl below).
#pragma once

#include "common.cuh"
#include "space.cuh"

namespace mptg {

struct KnnIndex {
    uint32_t count = 0;  // number of leading points covered by the index (0 = none)
    void* mem = nullptr;
    uint64_t* devStats = nullptr;
};

inline void knnIndexFree(KnnIndex& ix) {
    if (ix.mem) cudaFree(ix.mem);
    if (ix.devStats) cudaFree(ix.devStats);
    ix = KnnIndex();
}

inline int knnAutoStrategy(uint32_t /*size*/, uint32_t /*Q*/, const KnnIndex& /*ix*/) { return MPTG_KNN_BRUTE; }

template <typename S>
int knnBuildIndex(mptg_ctx* ctx, KnnIndex&, const mptg_space_desc&, const S*, uint32_t, uint32_t) {
    return fail(ctx, MPTG_ERR_UNSUPPORTED, "kNN spatial index not built into this library");
}
template <typename S>
int knnEnsureIndex(mptg_ctx* ctx, KnnIndex& ix, const mptg_space_desc& sp, const S* pts, uint32_t stride, uint32_t n) {
    return knnBuildIndex<S>(ctx, ix, sp, pts, stride, n);
}
template <typename S>
int knnBvhQuery(mptg_ctx* ctx, KnnIndex&, const mptg_space_desc&, const S*, uint32_t, uint32_t, double, uint32_t,
                uint32_t, uint32_t*, S*, uint32_t*, uint64_t*) {
    return fail(ctx, MPTG_ERR_UNSUPPORTED, "kNN spatial index not built into this library");
}
inline int knnIndexReadStats(mptg_ctx*, KnnIndex&, uint64_t*) { return MPTG_OK; }

}  // namespace mptg
