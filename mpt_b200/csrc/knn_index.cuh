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
    // SE(3)/f32 only (device build): what the conservative prefilters read, one 128-bit load per lane
    //   leafH   [leaf][32] uint4 = half2 (qx,qy) (qz,qw) (tx,ty) (tz,0) of the lane's point, translations times tScale
    //   cap[l]  [block][3][32] float4 per child: (c0 c1 c2 c3) (cos rho, sin rho, tlo.x, thi.x) (tlo.y, thi.y, tlo.z, thi.z)
    //           rotation part of a node = geodesic cap on RP^3: centre quaternion c and angular radius rho with
    //           acos(|p.c| / (|p||c|)) <= rho for every member p; translation part = its axis-aligned box
    uint32_t* leafH = nullptr;
    void* cap[BVH_MAXL] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float errQ = 0.f, errT = 0.f;  // max |half(v) - v| over the stored quaternion / translation coordinates
    float normMax = 1.f;           // max |quaternion| over the indexed points (rounded up)
    float tScale = 1.f;            // power of two: leafH translations are stored times tScale, max |t| tScale in (1024, 2048]
    unsigned long long* devStats = nullptr;  // [0] leaves visited, [1] inner nodes visited
    uint64_t builds = 0;
    uint32_t capacityHint = 0;  // capacity of the store: the image and the build's work space are sized for it once, so that
                                // rebuilding a growing set does not go through cudaFree / cudaMalloc (they stall for
                                // milliseconds to hundreds of milliseconds on a busy host)
};

// Indexed part of the tail (points inserted since the last build of the tree): each batch of points that has arrived
// between two searches is sorted along a Morton curve and cut into 32-point leaves with boxes -- a flat, one-level
// index that costs one small sort to extend.  Planner waves insert thousands of points per wave and rebuild the tree
// only every few waves; without this the exhaustive scan of the tail was half of a device-resident planner wave.
constexpr uint32_t TAIL_REBUILD_LIMIT = 131072;                   // the tree is rebuilt before the tail outgrows this (knnEnsureIndex)
constexpr uint32_t TAIL_MAX_POINTS = TAIL_REBUILD_LIMIT + 8192;  // room for the 32-point padding of every chunk
constexpr uint32_t TAIL_MIN_CHUNK = 1024;           // fewer new points than this are scanned exhaustively
struct KnnTail {
    uint32_t base = 0;      // first point covered (== index.count when valid)
    uint32_t covered = 0;   // points [base, base + covered) are in the leaves below
    uint32_t nLeaves = 0;   // 32-point leaves (each chunk padded to a multiple of 32, perm = MPTG_NO_INDEX for padding)
    uint64_t forBuild = ~0ull;  // KnnIndex::builds this tail belongs to
    void* mem = nullptr;
    void* leafPts = nullptr;   // [leaf][D][32]
    uint32_t* perm = nullptr;  // [leaf * 32] insertion index
    void* box = nullptr;       // [block][2D][32], block = leaf / 32
};

inline void knnIndexFree(KnnIndex& ix) {
    if (ix.mem) cudaFree(ix.mem);
    if (ix.devStats) cudaFree(ix.devStats);
    ix = KnnIndex();
}


// Device build of the index for float32 spaces (knn_build.cu).  Fills `ix` like the host build does.
int knnBuildIndexGpu(mptg_ctx* ctx, KnnIndex& ix, const mptg_space_desc& space, const float* ptsDev, uint32_t stride, uint32_t n);
int knnBuildIndexGpu(mptg_ctx* ctx, KnnIndex& ix, const mptg_space_desc& space, const double* ptsDev, uint32_t stride, uint32_t n);
// Append points [first, first + count) of the store to the tail's leaves (knn_build.cu).
int knnTailAppend(mptg_ctx* ctx, KnnTail& tail, const mptg_space_desc& space, const float* ptsDev, uint32_t stride, uint32_t first, uint32_t count);
int knnTailAppend(mptg_ctx* ctx, KnnTail& tail, const mptg_space_desc& space, const double* ptsDev, uint32_t stride, uint32_t first, uint32_t count);

}  // namespace mptg
