// knn_index.cuh -- the device image of the kNN index (shared by the search kernels, the host build and
// the device build).  Layouts are described in knn_bvh.cuh.
#pragma once

#include "common.cuh"

namespace mptg {

constexpr int BVH_MAXL = 5;  // 32^5 leaves -> up to 2^30 points
constexpr uint32_t BVH_DEAD = 0xFFFFFFFFu;

struct KnnIndex {
    uint32_t count = 0;    // points covered (a prefix of the store)
    uint32_t nPad = 0;     // leaves * 32
    int top = 0;           // top level (0 = leaves)
    uint32_t nNodes[BVH_MAXL] = {0, 0, 0, 0, 0};
    void* mem = nullptr;   // one block: leaf points, perm, boxes
    size_t memBytes = 0;
    void* leafPts = nullptr;  // [leaf][D][32]
    uint32_t* perm = nullptr;
    void* box[BVH_MAXL] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // [block][2D][32]
    // SE(3)/f32 only: half-precision copies that feed the conservative prefilters (half the L2 traffic)
    uint32_t* leafH = nullptr;                                                 // [leaf][4][32] half2: (qx,qy) (qz,qw) (tx,ty) (tz,0)
    uint32_t* boxH[BVH_MAXL] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // [block][7][32] half2 (lo rounded down, hi rounded up)
    float errQ = 0.f, errT = 0.f;  // max |half(v) - v| over the stored quaternion / translation coordinates
    unsigned long long* devStats = nullptr;  // [0] leaves visited, [1] inner nodes visited
    uint64_t builds = 0;
};

inline void knnIndexFree(KnnIndex& ix) {
    if (ix.mem) cudaFree(ix.mem);
    if (ix.devStats) cudaFree(ix.devStats);
    ix = KnnIndex();
}


// Device build of the index for float32 spaces (knn_build.cu).  Fills `ix` like the host build does.
int knnBuildIndexGpu(mptg_ctx* ctx, KnnIndex& ix, const mptg_space_desc& space, const float* ptsDev, uint32_t stride, uint32_t n);
int knnBuildIndexGpu(mptg_ctx* ctx, KnnIndex& ix, const mptg_space_desc& space, const double* ptsDev, uint32_t stride, uint32_t n);

}  // namespace mptg
