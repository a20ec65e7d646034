// knn_bvh.cuh -- exact kNN over a 32-ary bounding-box hierarchy (the "GPU kd-tree" engine of knn.cu).
//
// Structure.  The indexed points are copied in a spatially sorted order (kd-style median splits on
// the widest weighted coordinate, split positions aligned to powers of 32) into their own SoA block.
// 32 consecutive points form a leaf, 32 consecutive leaves a level-1 node, and so on until at most
// 32 nodes remain (the top level).  Every node stores an axis-aligned box over the state scalars
// (for SO(3) parts: over the sign-canonicalised quaternion coefficients, w >= 0 -- q and -q are the
// same rotation and the metric only sees |a.b|, which is exact under negation).  Boxes are SoA per
// level, so the 32 lanes of a warp test the 32 children of a node with coalesced loads.
//
// Search.  One warp per query, no stack in memory: at each level the lower bounds of the current
// node's 32 children live one per lane; the warp repeatedly takes the smallest remaining bound
// (redux.min) and descends, so nearer subtrees are visited first and the k-th best distance shrinks
// early.  A subtree is skipped only when bound > current k-th distance (or > radius).
//
// Exactness.  bound(box, q) <= distance(p, q) for every p in the box, in floating point: each part
// of the bound is computed by the same operation sequence as distance() applied to the box corner
// that minimises it, and every operation in that sequence is monotone (IEEE rounding is monotone).
// The acos polynomial is not proven monotone in its last ulp, so its result is shaved by 8 eps.
// Results are therefore identical to the brute-force scan, including the (distance, index) ties.
#pragma once

#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <numeric>

#include "common.cuh"
#include "knn_index.cuh"
#include "../../include/mptg/mptg_space.h"
#include "topk.cuh"

namespace mptg {

// AUTO policy.  The tiled scan evaluates Q*N pairs at full machine width whatever the queries are; the tree visits
// far fewer points but each query is a chain of dependent node fetches, and the tree has to be rebuilt as the set
// grows.  Measured on B200 (tools/knn_crossover.py: static sets, k = 16 and 1; tools/planner_wave_profile.py and the
// demos: planner waves):
//   planar L2, double: the tree wins from 1,024 points on for every wave size (16,384 points, 2,048 queries: 0.120 ms
//                   scanned, 0.030 ms tree); PRRT* waves of 256 / 1,024 / 4,096 samples at 60K nodes: 0.60 / 0.90 /
//                   1.89 ms scanned, 0.41 / 0.40 / 0.49 ms through the tree                 -> scan up to 2^23 pairs
//   SE(3), float:   the tree wins for waves of >= 2,048 queries at every size, ties at 256 queries (65,536 points,
//                   2,048 queries: 0.70 ms scanned, 0.24 ms tree); device-resident PRRT / PRRT* to 200K nodes:
//                   3.4 / 1.3 M nodes/s with the old 2^29-pair rule, 5.1 / 1.9 M with 2^23..2^26 -> scan up to 2^24 pairs
//   8 scalars and more (N-link arms): boxes prune little; 4,096-sample PPRM waves at 100K nodes: 8-D 7.1 ms scanned,
//                   7.3 ms tree; 16-D 19 ms scanned, 28 ms tree                             -> scan up to 2^36 pairs
// Below 16,384 points the scan stays: such sets are rebuilt every wave or two while a planner grows them.
inline int knnAutoStrategy(uint32_t size, uint32_t Q, int shape, int scalars) {
    if (size < 16384u) return MPTG_KNN_BRUTE;
    static const int forced = [] {  // MPTG_KNN_AUTO_LOG2_PAIRS: tuning experiments
        const char* e = getenv("MPTG_KNN_AUTO_LOG2_PAIRS");
        const int v = e ? atoi(e) : -1;
        return v > 62 ? 62 : v;
    }();
    int log2Pairs = 26;
    if (shape == SHAPE_L2_2 || shape == SHAPE_L2_3) log2Pairs = 23;
    else if (shape == SHAPE_SE3) log2Pairs = 24;
    else if ((shape == SHAPE_GENERIC || shape == SHAPE_L1) && scalars >= 8) log2Pairs = 36;
    if (forced >= 0) log2Pairs = forced;
    return (unsigned long long)size * Q <= (1ull << log2Pairs) ? MPTG_KNN_BRUTE : MPTG_KNN_BVH;
}

// ------------------------------------------------------------------ lower bounds
namespace dev {

// generic: lo/hi are callables c -> scalar (absolute scalar index), q likewise
template <typename S, typename LO, typename HI, typename QF>
MPTG_HD S boxLowerBound(const DevSpace<S>& sp, LO lo, HI hi, QF q) {
    S total = S(0);
    for (int i = 0; i < sp.nParts; ++i) {
        const int off = sp.off[i];
        S d;
        if (sp.kind[i] == MPTG_PART_SO3) {
            S dotHi = S(0), dotLo = S(0);
            for (int j = 0; j < 4; ++j) {
                const S v = q(off + j);
                const S l = lo(off + j), h = hi(off + j);
                const S cHi = v >= S(0) ? h : l;
                const S cLo = v >= S(0) ? l : h;
                if (j == 0) {
                    dotHi = cHi * v;
                    dotLo = cLo * v;
                } else {
                    dotHi = fp::fma_(cHi, v, dotHi);
                    dotLo = fp::fma_(cLo, v, dotLo);
                }
            }
            S ad = dotHi > -dotLo ? dotHi : -dotLo;
            if (ad > S(1)) ad = S(1);
            if (!(ad > S(0))) ad = S(0);
            d = fp::acos01(ad) * (S(1) - S(8) * fp::consts<S>::eps());
        } else {
            const S pi = fp::consts<S>::pi();
            S acc = S(0);
            for (int j = 0; j < sp.dim[i]; ++j) {
                const S v = q(off + j);
                const S l = lo(off + j), h = hi(off + j);
                S e;
                if (sp.kind[i] == MPTG_PART_SO2) {
                    if (v >= l && v <= h) {
                        e = S(0);
                    } else {
                        S d0 = fp::abs_(l - v), d1 = fp::abs_(h - v);
                        if (d0 > pi) d0 = S(2) * pi - d0;
                        if (d1 > pi) d1 = S(2) * pi - d1;
                        e = d0 < d1 ? d0 : d1;
                        if (!(e > S(0))) e = S(0);
                    }
                } else {
                    e = v < l ? l - v : (v > h ? h - v : S(0));
                }
                const S ae = fp::abs_(e);
                if (sp.p[i] == 2) acc = (j == 0) ? e * e : fp::fma_(e, e, acc);
                else if (sp.p[i] == 1) acc = (j == 0) ? ae : acc + ae;
                else acc = (j == 0) ? ae : (ae > acc ? ae : acc);
            }
            d = sp.p[i] == 2 ? fp::sqrt_(acc) : acc;
        }
        if (sp.weighted[i]) d = d * sp.weight[i];
        total = (i == 0) ? d : total + d;
    }
    return total;
}

}  // namespace dev

// ------------------------------------------------------------------ search kernel
// Device layouts are blocked so that one address computation serves a whole node visit:
//   leaves  leafPts[leaf][c][32]            (scalar c of the leaf's 32 points)
//   boxes   box[l][block][r][32], r < D: lo of scalar r, r >= D: hi of scalar r - D; block = node / 32
template <typename S>
struct BvhArgs {
    const S* leafPts;
    const uint32_t* perm;
    const S* box[BVH_MAXL];
    const uint32_t* leafH;  // nullptr when the compressed copies / caps are not available
    const float4* cap[BVH_MAXL];
    float errQ, errT, normMax, tScale, tScaleInv;  // cap image (knn_se3.cuh)
    uint32_t nNodes[BVH_MAXL];
    int top;
    const S* queries;
    const uint32_t* order;  // optional: processing order of the queries (spatially sorted), or nullptr
    uint32_t* orderKeys;    // key pass only: leaf reached by greedy descent, per query
    uint32_t* orderHist;    // key pass only: histogram over (leaf >> orderShift)
    uint32_t orderShift;
    uint32_t Q, k;
    S radius;
    uint32_t idxMul, idxAdd;
    uint32_t* idxOut;
    S* distOut;
    uint32_t* countOut;
    unsigned long long* stats;
    uint32_t* cursor;  // next position of the wave (persistent warps), zeroed before the launch
    const uint32_t* nActive;  // ordered waves with a per-query cap: the order lists only the queries whose cap is >= 0, this many
    const uint32_t* gid;  // optional: reported index of stored point i is gid[i] (spatially sharded sets) instead of i * idxMul + idxAdd
    const S* qcap;     // optional per-query radius cap (sharded search): < 0 skips the query and leaves its output row alone
    float* rootLb;     // root-bound pass only: lower bound of every query to the whole indexed set
    DevSpace<S> sp;
};

template <typename S>
__device__ __forceinline__ uint32_t boundKey(S lb);
template <>
__device__ __forceinline__ uint32_t boundKey<float>(float lb) {
    return __float_as_uint(lb + 0.0f);  // non-negative floats order like their bit patterns (+0.0f: never -0)
}
template <>
__device__ __forceinline__ uint32_t boundKey<double>(double lb) {
    return __float_as_uint(__double2float_rd(lb) + 0.0f);  // rounded down: still a lower bound
}
template <typename S>
__device__ __forceinline__ float thrAsFloat(S thr);
template <>
__device__ __forceinline__ float thrAsFloat<float>(float thr) {
    return thr;
}
template <>
__device__ __forceinline__ float thrAsFloat<double>(double thr) {
    return __double2float_ru(thr);
}

__device__ __forceinline__ float sqrtApprox(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Cheap SE(3) lower bound from |dot| (or its box maximum) and the squared translation distance
// (or its box minimum).  Uses acos(x) >= sqrt(2 - 2x) on [0,1], approximate square roots and fused
// arithmetic; the slack terms (|dot| + 1e-6, result * (1 - 1e-5)) dominate every rounding error of
// this expression and of the exact distance (a handful of 6e-8 relative errors), so
// bound <= fl-distance for every point concerned.
__device__ __forceinline__ float se3CheapBound(float ad, float s, float w0, float w1) {
    const float t = fmaxf(0.0f, __fmaf_rn(-2.0f, ad + 1e-6f, 2.0f));
    return __fmaf_rn(w0, sqrtApprox(t), w1 * sqrtApprox(s)) * (1.0f - 1e-5f);
}

#ifndef MPTG_BVH_WARPS
#define MPTG_BVH_WARPS 8
#endif
constexpr int BVH_WARPS = MPTG_BVH_WARPS;  // warps (queries) per CTA; 4 and 16 measured no faster than 8

template <typename S, int SHAPE, int KPL>
struct BvhWalk {
    const BvhArgs<S>& a;
    const S* myq;  // query in shared memory (generic path)
    S qr[7];       // query in registers (SE(3) path)
    float w0, w1;
    WarpTopK<S, KPL> top;
    int lane;
    uint32_t leaves = 0, inner = 0;

    __device__ __forceinline__ BvhWalk(const BvhArgs<S>& args, const S* q, int ln) : a(args), myq(q), lane(ln) {}

    // Pruning threshold min(k-th distance, radius), cached as a float (rounded up) and refreshed after
    // every offer.
    float thrF;
    S rad;  // search radius of this query: a.radius, or tighter (tail search bounded by the tree's k-th distance)
    __device__ __forceinline__ void refreshThr() {
        const S t = top.kthD < rad ? top.kthD : rad;
        thrF = thrAsFloat<S>(t);
    }
    __device__ __forceinline__ void start() {
        top.init(a.k);
        rad = a.radius;
        refreshThr();
    }
    __device__ __forceinline__ float threshold() const { return thrF; }
    __device__ __forceinline__ uint32_t reported(uint32_t orig) const {
        return a.gid ? (orig != MPTG_NO_INDEX ? __ldg(a.gid + orig) : MPTG_NO_INDEX) : orig * a.idxMul + a.idxAdd;
    }

    // key (float bits of a lower bound) of child `lane` of `block` at level L
    template <int L>
    __device__ __forceinline__ uint32_t childKey(uint32_t block) const {
        const uint32_t node = block * 32u + (uint32_t)lane;
        if (node >= a.nNodes[L]) return BVH_DEAD;
        const int D = a.sp.D;
        const S* b = a.box[L] + ((size_t)block * (size_t)(2 * D)) * 32u + lane;
        if (SHAPE == SHAPE_SE3 && sizeof(S) == 4) {
            float dotHi, dotLo, acc = 0.0f;
            {
                {
                    const float l = __ldg((const float*)b), h = __ldg((const float*)b + 7 * 32);
                    const float v = (float)qr[0];
                    dotHi = (v >= 0.0f ? h : l) * v;
                    dotLo = (v >= 0.0f ? l : h) * v;
                }
#pragma unroll
                for (int j = 1; j < 4; ++j) {
                    const float l = __ldg((const float*)b + j * 32), h = __ldg((const float*)b + (7 + j) * 32);
                    const float v = (float)qr[j];
                    dotHi = __fmaf_rn(v >= 0.0f ? h : l, v, dotHi);
                    dotLo = __fmaf_rn(v >= 0.0f ? l : h, v, dotLo);
                }
#pragma unroll
                for (int j = 4; j < 7; ++j) {
                    const float l = __ldg((const float*)b + j * 32), h = __ldg((const float*)b + (7 + j) * 32);
                    const float v = (float)qr[j];
                    const float e = fmaxf(fmaxf(l - v, v - h), 0.0f);
                    acc = __fmaf_rn(e, e, acc);
                }
            }
            const float ad = fminf(1.0f, fmaxf(dotHi, -dotLo));
            return __float_as_uint(se3CheapBound(ad, acc, w0, w1) + 0.0f);
        } else {
            const S lb = dev::boxLowerBound<S>(
                a.sp, [&](int c) { return __ldg(b + c * 32); }, [&](int c) { return __ldg(b + (D + c) * 32); },
                [&](int c) { return myq[c]; });
            return boundKey<S>(lb);
        }
    }

    __device__ __forceinline__ void leaf(uint32_t node) {
        ++leaves;
        const int D = a.sp.D;
        const uint32_t p = node * 32u + (uint32_t)lane;
        const uint32_t orig = __ldg(a.perm + p);
        const bool have = orig != MPTG_NO_INDEX;
        const S* pt = a.leafPts + ((size_t)node * (size_t)D) * 32u + lane;
        if (SHAPE == SHAPE_SE3 && sizeof(S) == 4) {
            float pv[7];
#pragma unroll
            for (int c = 0; c < 7; ++c) pv[c] = __ldg((const float*)pt + c * 32);
            // exact-order pieces shared by the prefilter and the distance (mptg_space.h)
            float dot = pv[0] * (float)qr[0];
            dot = __fmaf_rn(pv[1], (float)qr[1], dot);
            dot = __fmaf_rn(pv[2], (float)qr[2], dot);
            dot = __fmaf_rn(pv[3], (float)qr[3], dot);
            const float d0 = pv[4] - (float)qr[4], d1 = pv[5] - (float)qr[5], d2 = pv[6] - (float)qr[6];
            float s2 = d0 * d0;
            s2 = __fmaf_rn(d1, d1, s2);
            s2 = __fmaf_rn(d2, d2, s2);
            const float ad = fminf(1.0f, fabsf(dot));
            const bool maybe = have && se3CheapBound(ad, s2, w0, w1) <= threshold();
            if (!__any_sync(FULL_MASK, maybe)) return;
            float dr = fp::acos01(ad);
            if (a.sp.weighted[0]) dr = dr * (float)a.sp.weight[0];
            float dt = fp::sqrt_(s2);
            if (a.sp.weighted[1]) dt = dt * (float)a.sp.weight[1];
            top.offer(maybe, (S)(dr + dt), reported(orig), rad, lane);
        } else {
            const S dist = dev::distance<S>(
                a.sp, [&](int c) { return __ldg(pt + c * 32); }, [&](int c) { return myq[c]; });
            top.offer(have, dist, reported(orig), rad, lane);
        }
        refreshThr();
    }

    // take the child with the smallest remaining bound; false when none is left within the threshold
    __device__ __forceinline__ bool pick(uint32_t& key, uint32_t block, uint32_t& node, uint32_t& best) const {
        best = __reduce_min_sync(FULL_MASK, key);
        if (best == BVH_DEAD || __uint_as_float(best) > thrF) return false;
        const int src = __ffs(__ballot_sync(FULL_MASK, key == best)) - 1;
        if (lane == src) key = BVH_DEAD;
        node = block * 32u + (uint32_t)src;
        return true;
    }
    __device__ __forceinline__ bool pick(uint32_t& key, uint32_t block, uint32_t& node) const {
        uint32_t best;
        return pick(key, block, node, best);
    }

    // visit the children (level L nodes) of `block`, nearest bound first
    template <int L>
    __device__ __forceinline__ void descend(uint32_t block) {
        uint32_t key = childKey<L>(block);
        uint32_t node;
        {
            while (pick(key, block, node)) {
                if constexpr (L == 0) {
                    leaf(node);
                } else {
                    ++inner;
                    descend<L - 1>(node);
                }
            }
        }
    }
};

// ---- query ordering: queries that end up in the same region of the tree are processed by
// neighbouring warps, so the node and leaf lines they touch are shared through L1.
// Key = the leaf reached by always following the smallest child bound (three node visits at 1M).
template <typename S, int SHAPE>
__global__ void __launch_bounds__(BVH_WARPS * 32) knnBvhKeyKernel(const BvhArgs<S> a) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    S* qsm = reinterpret_cast<S*>(smemRaw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = a.sp.D;
    const uint32_t q = blockIdx.x * BVH_WARPS + warp;
    if (q >= a.Q) return;
    if (a.qcap && a.qcap[q] < S(0)) return;  // not searched in this pass: not in the order
    S* myq = qsm + warp * D;
    for (int c = lane; c < D; c += 32) myq[c] = a.queries[(size_t)q * D + c];
    __syncwarp();
    BvhWalk<S, SHAPE, 1> w(a, myq, lane);
    if (SHAPE == SHAPE_SE3) {
#pragma unroll
        for (int c = 0; c < 7; ++c) w.qr[c] = myq[c];
        w.w0 = a.sp.weighted[0] ? (float)a.sp.weight[0] : 1.0f;
        w.w1 = a.sp.weighted[1] ? (float)a.sp.weight[1] : 1.0f;
    }
    uint32_t node = 0;
    auto step = [&](uint32_t key) {
        const uint32_t best = __reduce_min_sync(FULL_MASK, key);
        const int src = __ffs(__ballot_sync(FULL_MASK, key == best)) - 1;
        node = node * 32u + (uint32_t)(src < 0 ? 0 : src);
    };
    switch (a.top) {  // fall through from the top level down to the leaves
        case 4: step(w.template childKey<4>(node));
        case 3: step(w.template childKey<3>(node));
        case 2: step(w.template childKey<2>(node));
        case 1: step(w.template childKey<1>(node));
        default: step(w.template childKey<0>(node));
    }
    if (lane == 0) {
        a.orderKeys[q] = node;
        atomicAdd(a.orderHist + (node >> a.orderShift), 1u);
    }
}

// Lower bound of every query to the whole indexed set: the smallest child bound of the top node (sharded search: which
// shards a query has to visit at all).  Rounded down to float.
template <typename S, int SHAPE>
__global__ void __launch_bounds__(BVH_WARPS * 32) knnBvhRootBoundKernel(const BvhArgs<S> a) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    S* qsm = reinterpret_cast<S*>(smemRaw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = a.sp.D;
    const uint32_t q = blockIdx.x * BVH_WARPS + warp;
    if (q >= a.Q) return;
    S* myq = qsm + warp * D;
    for (int c = lane; c < D; c += 32) myq[c] = a.queries[(size_t)q * D + c];
    __syncwarp();
    BvhWalk<S, SHAPE, 1> w(a, myq, lane);
    if (SHAPE == SHAPE_SE3) {
#pragma unroll
        for (int c = 0; c < 7; ++c) w.qr[c] = myq[c];
        w.w0 = a.sp.weighted[0] ? (float)a.sp.weight[0] : 1.0f;
        w.w1 = a.sp.weighted[1] ? (float)a.sp.weight[1] : 1.0f;
    }
    uint32_t key;
    switch (a.top) {
        case 4: key = w.template childKey<4>(0); break;
        case 3: key = w.template childKey<3>(0); break;
        case 2: key = w.template childKey<2>(0); break;
        case 1: key = w.template childKey<1>(0); break;
        default: key = w.template childKey<0>(0); break;
    }
    const uint32_t best = __reduce_min_sync(FULL_MASK, key);
    if (lane == 0) a.rootLb[q] = best == BVH_DEAD ? INFINITY : __uint_as_float(best);
}

// exclusive scan of up to 65536 bins by one CTA
__global__ void __launch_bounds__(1024) knnOrderScanKernel(uint32_t* hist, uint32_t bins, uint32_t* total) {
    __shared__ uint32_t part[1024];
    const uint32_t per = (bins + 1023u) / 1024u;
    const uint32_t b0 = threadIdx.x * per;
    uint32_t sum = 0;
    for (uint32_t i = b0; i < b0 + per && i < bins; ++i) sum += hist[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (uint32_t o = 1; o < 1024; o <<= 1) {
        const uint32_t v = threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    if (threadIdx.x == 1023) *total = part[1023];
    uint32_t run = part[threadIdx.x] - sum;
    for (uint32_t i = b0; i < b0 + per && i < bins; ++i) {
        const uint32_t c = hist[i];
        hist[i] = run;
        run += c;
    }
}

template <typename S>
__global__ void knnOrderScatterKernel(const uint32_t* keys, uint32_t* cursor, uint32_t shift, uint32_t Q, const S* __restrict__ qcap, uint32_t* order) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    if (qcap && qcap[q] < S(0)) return;
    order[atomicAdd(cursor + (keys[q] >> shift), 1u)] = q;
}

#ifndef MPTG_BVH_MIN_CTAS
#define MPTG_BVH_MIN_CTAS 1
#endif
// Persistent warps: the grid is sized to the machine and every warp draws the next position of the (spatially sorted)
// wave from a counter until the wave is exhausted.  A search takes 0.5x to 3x the mean, and with one query per warp of
// an 8-warp CTA a slot stayed occupied until its slowest warp was done: ncu showed 23 of the 32 resident warps active
// and the kernel waiting on loads (long scoreboard 3.3 per issue) rather than issuing.
template <typename S, int SHAPE, int KPL>
__global__ void __launch_bounds__(BVH_WARPS * 32, MPTG_BVH_MIN_CTAS) knnBvhKernel(const BvhArgs<S> a) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    S* qsm = reinterpret_cast<S*>(smemRaw);  // [BVH_WARPS][D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = a.sp.D;
    S* myq = qsm + warp * D;
    uint32_t leaves = 0, inner = 0;
    const uint32_t nSlots = a.nActive ? __ldg(a.nActive) : a.Q;
    for (;;) {
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(a.cursor, 1u);
        slot = __shfl_sync(FULL_MASK, slot, 0);
        if (slot >= nSlots) break;  // warp-uniform; no block-wide barriers in this kernel
        const uint32_t q = a.order ? __ldg(a.order + slot) : slot;
        __syncwarp();
        for (int c = lane; c < D; c += 32) myq[c] = a.queries[(size_t)q * D + c];
        __syncwarp();
        BvhWalk<S, SHAPE, KPL> w(a, myq, lane);
        if (SHAPE == SHAPE_SE3) {
#pragma unroll
            for (int c = 0; c < 7; ++c) w.qr[c] = myq[c];
            w.w0 = a.sp.weighted[0] ? (float)a.sp.weight[0] : 1.0f;
            w.w1 = a.sp.weighted[1] ? (float)a.sp.weight[1] : 1.0f;
        }
        w.start();
        if (a.qcap) {
            const S c = a.qcap[q];
            if (c < S(0)) continue;
            if (c < w.rad) w.rad = c;
            w.refreshThr();
        }
        switch (a.top) {
            case 0: w.template descend<0>(0); break;
            case 1: w.template descend<1>(0); break;
            case 2: w.template descend<2>(0); break;
            case 3: w.template descend<3>(0); break;
            default: w.template descend<4>(0); break;
        }
        const uint32_t count = w.top.store(a.k, a.idxOut + (size_t)q * a.k, a.distOut + (size_t)q * a.k, lane);
        if (a.countOut && lane == 0) a.countOut[q] = count;
        leaves += w.leaves;
        inner += w.inner;
    }
    if (lane == 0 && a.stats) {
        atomicAdd(a.stats + 0, (unsigned long long)leaves);
        atomicAdd(a.stats + 1, (unsigned long long)inner);
    }
}

// ---- search of the tail's Morton-sorted leaves (KnnTail, knn_index.cuh) with the same walker: args.box[0] / leafPts /
// perm describe the flat list of 32-point leaves (args.nNodes[0] of them, no half-precision copies).  One warp per query
// tests the leaf boxes 32 at a time and visits the leaves whose bound is within the threshold.  `cap` ([Q][k] distances
// of the tree search of the same queries) bounds the search from the start: a tail point farther than the tree's k-th
// neighbour cannot be among the k nearest of the union.
template <typename S, int SHAPE, int KPL>
__global__ void __launch_bounds__(BVH_WARPS * 32) knnTailKernel(const BvhArgs<S> a, const S* __restrict__ cap) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    S* qsm = reinterpret_cast<S*>(smemRaw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = a.sp.D;
    const uint32_t q = blockIdx.x * BVH_WARPS + warp;
    if (q >= a.Q) return;
    S* myq = qsm + warp * D;
    for (int c = lane; c < D; c += 32) myq[c] = a.queries[(size_t)q * D + c];
    __syncwarp();
    BvhWalk<S, SHAPE, KPL> w(a, myq, lane);
    if (SHAPE == SHAPE_SE3) {
#pragma unroll
        for (int c = 0; c < 7; ++c) w.qr[c] = myq[c];
        w.w0 = a.sp.weighted[0] ? (float)a.sp.weight[0] : 1.0f;
        w.w1 = a.sp.weighted[1] ? (float)a.sp.weight[1] : 1.0f;
    }
    w.start();
    if (cap) {
        const S c = cap[(size_t)q * a.k + (a.k - 1)];  // +inf when the tree returned fewer than k
        if (c < w.rad) w.rad = c;
        w.refreshThr();
    }
    const uint32_t nBlocks = (a.nNodes[0] + 31u) / 32u;
    for (uint32_t b = 0; b < nBlocks; ++b) {
        const uint32_t key = w.template childKey<0>(b);
        unsigned m = __ballot_sync(FULL_MASK, __uint_as_float(key) <= w.thrF);  // BVH_DEAD is a NaN pattern
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1u;
            const uint32_t lkey = __shfl_sync(FULL_MASK, key, src);
            if (__uint_as_float(lkey) > w.thrF) continue;  // the threshold has shrunk since the vote
            w.leaf(b * 32u + (uint32_t)src);
        }
    }
    w.top.store(a.k, a.idxOut + (size_t)q * a.k, a.distOut + (size_t)q * a.k, lane);
}

}  // namespace mptg
#include "knn_se3.cuh"
namespace mptg {

// ------------------------------------------------------------------ host: build
namespace {

template <typename S>
struct HostBuild {
    int D;
    std::vector<S> pts;        // AoS, canonicalised
    std::vector<S> w;          // per scalar weight for the split heuristic
    std::vector<uint32_t> order;

    void split(uint32_t b, uint32_t e) {
        const uint32_t len = e - b;
        if (len <= 32) return;
        uint32_t blk = 32;
        while ((uint64_t)blk * 32 < len) blk *= 32;
        const uint32_t nblk = (len + blk - 1) / blk;
        const uint32_t mid = b + ((nblk + 1) / 2) * blk;
        // widest weighted coordinate
        int axis = 0;
        S best = S(-1);
        for (int c = 0; c < D; ++c) {
            S mn = pts[(size_t)order[b] * D + c], mx = mn;
            for (uint32_t i = b + 1; i < e; ++i) {
                const S v = pts[(size_t)order[i] * D + c];
                mn = v < mn ? v : mn;
                mx = v > mx ? v : mx;
            }
            const S ext = (mx - mn) * w[c];
            if (ext > best) best = ext, axis = c;
        }
        std::nth_element(order.begin() + b, order.begin() + mid, order.begin() + e, [&](uint32_t x, uint32_t y) {
            const S vx = pts[(size_t)x * D + axis], vy = pts[(size_t)y * D + axis];
            return vx < vy || (vx == vy && x < y);
        });
        if (len > 65536) {
#pragma omp task
            split(b, mid);
#pragma omp task
            split(mid, e);
#pragma omp taskwait
        } else {
            split(b, mid);
            split(mid, e);
        }
    }
};

}  // namespace

template <typename S>
int knnBuildIndex(mptg_ctx* ctx, KnnIndex& ix, const mptg_space_desc& space, const S* ptsDev, uint32_t stride, uint32_t n) {
    if (n == 0) {
        ix.count = 0;
        return MPTG_OK;
    }
    // the index is built on the device (knn_build.cu); MPTG_KNN_HOST_BUILD=1 keeps the host path below for comparison
    {
        const char* env = getenv("MPTG_KNN_HOST_BUILD");
        if (!(env && env[0] == '1')) return knnBuildIndexGpu(ctx, ix, space, ptsDev, stride, n);
    }
    const DevSpace<S> sp = makeDevSpace<S>(space);
    const int D = sp.D;
    // 1. fetch the points (SoA rows -> AoS on the host), canonicalise rotations
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<S> soa((size_t)D * n);
    MPTG_CUDA(ctx, cudaMemcpy2DAsync(soa.data(), (size_t)n * sizeof(S), ptsDev, (size_t)stride * sizeof(S), (size_t)n * sizeof(S), D,
                                     cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    HostBuild<S> hb;
    hb.D = D;
    hb.pts.resize((size_t)n * D);
    for (int c = 0; c < D; ++c)
        for (uint32_t i = 0; i < n; ++i) hb.pts[(size_t)i * D + c] = soa[(size_t)c * n + i];
    hb.w.assign(D, S(1));
    for (int i = 0; i < sp.nParts; ++i) {
        for (int j = 0; j < sp.dim[i]; ++j) hb.w[sp.off[i] + j] = sp.weight[i];
        if (sp.kind[i] == MPTG_PART_SO3)
            for (uint32_t p = 0; p < n; ++p) {
                S* qv = &hb.pts[(size_t)p * D + sp.off[i]];
                if (qv[3] < S(0)) qv[0] = -qv[0], qv[1] = -qv[1], qv[2] = -qv[2], qv[3] = -qv[3];
            }
    }
    // 2. spatial order
    hb.order.resize(n);
    std::iota(hb.order.begin(), hb.order.end(), 0u);
#pragma omp parallel
#pragma omp single
    hb.split(0, n);
    // 3. levels: SoA boxes per level on the host first
    KnnIndex nx;
    nx.count = n;
    nx.nNodes[0] = (n + 31) / 32;
    nx.nPad = nx.nNodes[0] * 32;
    nx.top = 0;
    while (nx.nNodes[nx.top] > 32) {
        if (nx.top + 1 >= BVH_MAXL) return fail(ctx, MPTG_ERR_CAPACITY, "kNN index: too many points for %d levels", BVH_MAXL);
        nx.nNodes[nx.top + 1] = (nx.nNodes[nx.top] + 31) / 32;
        ++nx.top;
    }
    const S inf = fp::consts<S>::inf();
    std::vector<std::vector<S>> lo(nx.top + 1), hi(nx.top + 1);
    for (int l = 0; l <= nx.top; ++l) {
        const uint32_t nn = nx.nNodes[l];
        lo[l].assign((size_t)D * nn, inf);
        hi[l].assign((size_t)D * nn, -inf);
        if (l == 0) {
#pragma omp parallel for schedule(static)
            for (int64_t j = 0; j < (int64_t)nn; ++j)
                for (uint32_t i = (uint32_t)j * 32; i < (uint32_t)j * 32 + 32 && i < n; ++i) {
                    const S* pt = &hb.pts[(size_t)hb.order[i] * D];
                    for (int c = 0; c < D; ++c) {
                        lo[0][(size_t)c * nn + j] = pt[c] < lo[0][(size_t)c * nn + j] ? pt[c] : lo[0][(size_t)c * nn + j];
                        hi[0][(size_t)c * nn + j] = pt[c] > hi[0][(size_t)c * nn + j] ? pt[c] : hi[0][(size_t)c * nn + j];
                    }
                }
        } else {
            const uint32_t cn = nx.nNodes[l - 1];
            for (uint32_t j = 0; j < nn; ++j)
                for (int c = 0; c < D; ++c) {
                    S mn = inf, mx = -inf;
                    for (uint32_t i = j * 32; i < j * 32 + 32 && i < cn; ++i) {
                        mn = lo[l - 1][(size_t)c * cn + i] < mn ? lo[l - 1][(size_t)c * cn + i] : mn;
                        mx = hi[l - 1][(size_t)c * cn + i] > mx ? hi[l - 1][(size_t)c * cn + i] : mx;
                    }
                    lo[l][(size_t)c * nn + j] = mn;
                    hi[l][(size_t)c * nn + j] = mx;
                }
        }
    }
    // blocked device image
    size_t bytes = 0;
    auto take = [&](size_t b) {
        const size_t o = bytes;
        bytes += (b + 255) & ~(size_t)255;
        return o;
    };
    const size_t oPts = take((size_t)D * nx.nPad * sizeof(S));
    const size_t oPerm = take((size_t)nx.nPad * sizeof(uint32_t));
    size_t oBox[BVH_MAXL] = {0, 0, 0, 0, 0};
    uint32_t nBlocks[BVH_MAXL] = {0, 0, 0, 0, 0};
    for (int l = 0; l <= nx.top; ++l) {
        nBlocks[l] = (nx.nNodes[l] + 31) / 32;
        oBox[l] = take((size_t)nBlocks[l] * 2 * D * 32 * sizeof(S));
    }
    std::vector<unsigned char> host(bytes, 0);
    S* hp = reinterpret_cast<S*>(host.data() + oPts);
    uint32_t* hperm = reinterpret_cast<uint32_t*>(host.data() + oPerm);
    for (uint32_t i = 0; i < nx.nPad; ++i) {
        const uint32_t src = hb.order[i < n ? i : n - 1];  // padding repeats the last point (perm marks it unused)
        hperm[i] = i < n ? src : MPTG_NO_INDEX;
        const uint32_t leaf = i >> 5, ln = i & 31;
        for (int c = 0; c < D; ++c) hp[((size_t)leaf * D + c) * 32 + ln] = hb.pts[(size_t)src * D + c];
    }
    for (int l = 0; l <= nx.top; ++l) {
        S* bx = reinterpret_cast<S*>(host.data() + oBox[l]);
        const uint32_t nn = nx.nNodes[l];
        for (uint32_t b = 0; b < nBlocks[l]; ++b)
            for (uint32_t ln = 0; ln < 32; ++ln) {
                const uint32_t j = b * 32 + ln;
                for (int c = 0; c < D; ++c) {
                    bx[((size_t)b * 2 * D + c) * 32 + ln] = j < nn ? lo[l][(size_t)c * nn + j] : inf;
                    bx[((size_t)b * 2 * D + D + c) * 32 + ln] = j < nn ? hi[l][(size_t)c * nn + j] : -inf;
                }
            }
    }
    // 4. upload (reuse the block when it is large enough)
    void* mem = ix.mem;
    size_t memBytes = ix.memBytes;
    if (memBytes < bytes) {
        if (mem) MPTG_CUDA(ctx, cudaFree(mem));
        mem = nullptr;
        memBytes = bytes + bytes / 2;
        MPTG_CUDA(ctx, cudaMalloc(&mem, memBytes));
    }
    if (int rc = uploadSync(ctx, mem, host.data(), bytes)) return rc;
    unsigned long long* stats = ix.devStats;
    if (!stats) {
        MPTG_CUDA(ctx, cudaMalloc(&stats, 8 * sizeof(unsigned long long)));
        if (int rc = memsetSync(ctx, stats, 0, 8 * sizeof(unsigned long long))) return rc;
    }
    nx.mem = mem;
    nx.memBytes = memBytes;
    nx.devStats = stats;
    nx.builds = ix.builds + 1;
    nx.capacityHint = ix.capacityHint;
    nx.leafPts = (char*)mem + oPts;
    nx.perm = (uint32_t*)((char*)mem + oPerm);
    for (int l = 0; l <= nx.top; ++l) nx.box[l] = (char*)mem + oBox[l];
    ix = nx;
    return MPTG_OK;
}

// (Re)build when there is no index yet or the tail (points inserted since the build) has grown past
// min(count/2, 131072) points.  In between the tail is searched through its Morton-sorted leaves (KnnTail) and, for the
// newest points, by brute force (knn.cu), and merged.  Measured on planner waves (8,192 samples, 150K-node planar
// roadmap): a wave costs 0.6 ms, a rebuild ~6 ms (sorts, allocation growth, one synchronisation) -- with the earlier
// limit of min(count/4, 65536) a rebuild every seven waves cost more than the searches themselves, while the tail
// search grows by only ~2 us per 1,000 tail points.
template <typename S>
int knnEnsureIndex(mptg_ctx* ctx, KnnIndex& ix, const mptg_space_desc& sp, const S* pts, uint32_t stride, uint32_t n) {
    if (ix.count != 0 && ix.count <= n) {
        const uint32_t tail = n - ix.count;
        uint32_t limit = ix.count / 2;
        if (limit > TAIL_REBUILD_LIMIT) limit = TAIL_REBUILD_LIMIT;
        if (tail <= limit) return MPTG_OK;
    }
    return knnBuildIndex<S>(ctx, ix, sp, pts, stride, n);
}

// persistent grid: as many CTAs as the machine holds at once (occupancy of the instantiation), at most one warp per query
template <typename K, typename A>
int launchPersistent(mptg_ctx* ctx, K kernel, const A& a, uint32_t Q, size_t smem) {
    int perSm = 0;
    MPTG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kernel, BVH_WARPS * 32, smem));
    uint32_t ctas = (uint32_t)(perSm > 0 ? perSm : 1) * (uint32_t)ctx->smCount;
    const uint32_t need = (Q + BVH_WARPS - 1) / BVH_WARPS;
    if (ctas > need) ctas = need;
    kernel<<<ctas, BVH_WARPS * 32, smem, ctx->stream>>>(a);
    MPTG_LAUNCHED(ctx);
    return MPTG_OK;
}

// spatial processing order for large waves (skipped for small ones: three extra launches): queries that end up in the
// same region of the tree are processed by neighbouring warps, so the node and leaf lines they touch are shared in L1
template <typename S, typename KEYK>
int orderWave(mptg_ctx* ctx, BvhArgs<S>& a, KEYK keyKernel, size_t smem) {
    if (!(a.Q >= 4096 && a.top >= 1)) return MPTG_OK;
    const dim3 grid((a.Q + BVH_WARPS - 1) / BVH_WARPS), block(BVH_WARPS * 32);
    // bins of a few neighbouring leaves: the unit of locality is the level-0 block of 32 leaves, and the one-CTA scan of
    // the histogram is on the critical path of every wave (32,769 bins: 28 us; 4,097: a few)
    static const uint32_t maxBins = getenv("MPTG_ORDER_BINS") ? (uint32_t)atoi(getenv("MPTG_ORDER_BINS")) : 4096u;
    uint32_t shift = 0;
    while ((a.nNodes[0] >> shift) > maxBins) ++shift;
    const uint32_t bins = (a.nNodes[0] >> shift) + 1u;
    void* buf;
    int rc = scratch(ctx, 6, ((size_t)2 * a.Q + bins + 1) * sizeof(uint32_t), &buf);
    if (rc) return rc;
    uint32_t* order = (uint32_t*)buf;
    a.orderKeys = order + a.Q;
    a.orderHist = order + 2 * (size_t)a.Q;
    a.orderShift = shift;
    uint32_t* total = a.orderHist + bins;
    // With a per-query cap (sharded search) the order lists only the queries this pass searches -- at 8 GPUs one in
    // eight -- so the persistent warps do not draw, look up and skip the others one counter increment at a time.
    MPTG_CUDA(ctx, cudaMemsetAsync(a.orderHist, 0, bins * sizeof(uint32_t), ctx->stream));
    keyKernel<<<grid, block, smem, ctx->stream>>>(a);
    MPTG_LAUNCHED(ctx);
    knnOrderScanKernel<<<1, 1024, 0, ctx->stream>>>(a.orderHist, bins, total);
    MPTG_LAUNCHED(ctx);
    knnOrderScatterKernel<S><<<(a.Q + 255) / 256, 256, 0, ctx->stream>>>(a.orderKeys, a.orderHist, shift, a.Q, a.qcap, order);
    MPTG_LAUNCHED(ctx);
    a.order = order;
    if (a.qcap) a.nActive = total;
    return MPTG_OK;
}

template <typename S, int SHAPE>
int launchBvhShape(mptg_ctx* ctx, BvhArgs<S>& a) {
    const size_t smem = (size_t)BVH_WARPS * a.sp.D * sizeof(S);
    if (int rc = orderWave<S>(ctx, a, knnBvhKeyKernel<S, SHAPE>, smem)) return rc;
    if (a.k <= 32) return launchPersistent(ctx, knnBvhKernel<S, SHAPE, 1>, a, a.Q, smem);
    if (a.k <= 64) return launchPersistent(ctx, knnBvhKernel<S, SHAPE, 2>, a, a.Q, smem);
    return launchPersistent(ctx, knnBvhKernel<S, SHAPE, 4>, a, a.Q, smem);
}

// SE(3)/float32 over the cap image (knn_se3.cuh)
inline int launchSe3(mptg_ctx* ctx, BvhArgs<float>& a) {
    if (int rc = orderWave<float>(ctx, a, knnSe3KeyKernel, 0)) return rc;
    if (a.k <= 32) return launchPersistent(ctx, knnSe3Kernel<1>, a, a.Q, 0);
    if (a.k <= 64) return launchPersistent(ctx, knnSe3Kernel<2>, a, a.Q, 0);
    return launchPersistent(ctx, knnSe3Kernel<4>, a, a.Q, 0);
}
inline int launchSe3(mptg_ctx* ctx, BvhArgs<double>&) { return fail(ctx, MPTG_ERR_UNSUPPORTED, "cap image on a double-precision set"); }
inline void launchSe3RootBound(mptg_ctx* ctx, BvhArgs<float>& a, dim3 grid, dim3 block) { knnSe3RootBoundKernel<<<grid, block, 0, ctx->stream>>>(a); }
inline void launchSe3RootBound(mptg_ctx*, BvhArgs<double>&, dim3, dim3) {}

template <typename S>
int knnBvhQuery(mptg_ctx* ctx, KnnIndex& ix, const mptg_space_desc& space, const S* queries, uint32_t Q, uint32_t k,
                double radius, uint32_t idxMul, uint32_t idxAdd, uint32_t* idxOut, S* distOut, uint32_t* countOut,
                uint64_t* /*hostStats*/, const uint32_t* gid = nullptr, const S* qcap = nullptr, float* rootLbOut = nullptr) {
    BvhArgs<S> a{};
    a.gid = gid;
    a.qcap = qcap;
    a.rootLb = rootLbOut;
    a.leafPts = (const S*)ix.leafPts;
    a.perm = ix.perm;
    a.leafH = ix.leafH;
    a.errQ = ix.errQ;
    a.errT = ix.errT;
    for (int l = 0; l < BVH_MAXL; ++l) a.cap[l] = (const float4*)ix.cap[l];
    a.normMax = ix.normMax;
    a.tScale = ix.tScale;
    a.tScaleInv = 1.0f / ix.tScale;
    for (int l = 0; l < BVH_MAXL; ++l) {
        a.box[l] = (const S*)ix.box[l];
        a.nNodes[l] = ix.nNodes[l];
    }
    a.top = ix.top;
    a.queries = queries;
    a.Q = Q;
    a.k = k;
    a.radius = (radius >= 0 && radius == radius) ? (S)radius : fp::consts<S>::inf();
    a.idxMul = idxMul;
    a.idxAdd = idxAdd;
    a.idxOut = idxOut;
    a.distOut = distOut;
    a.countOut = countOut;
    a.stats = ix.devStats;
    a.cursor = reinterpret_cast<uint32_t*>(ix.devStats + 4);
    a.sp = makeDevSpace<S>(space);
    if (rootLbOut) {  // root-bound pass only
        const dim3 grid((Q + BVH_WARPS - 1) / BVH_WARPS), block(BVH_WARPS * 32);
        const size_t smem = (size_t)BVH_WARPS * a.sp.D * sizeof(S);
        if (classifySpace(space) == SHAPE_SE3) {
            if (sizeof(S) == 4 && ix.leafH) launchSe3RootBound(ctx, a, grid, block);
            else knnBvhRootBoundKernel<S, SHAPE_SE3><<<grid, block, smem, ctx->stream>>>(a);
        } else {
            knnBvhRootBoundKernel<S, SHAPE_GENERIC><<<grid, block, smem, ctx->stream>>>(a);
        }
        MPTG_LAUNCHED(ctx);
        return MPTG_OK;
    }
    MPTG_CUDA(ctx, cudaMemsetAsync(ix.devStats, 0, 8 * sizeof(unsigned long long), ctx->stream));
    if (classifySpace(space) == SHAPE_SE3) return (sizeof(S) == 4 && ix.leafH) ? launchSe3(ctx, a) : launchBvhShape<S, SHAPE_SE3>(ctx, a);
    return launchBvhShape<S, SHAPE_GENERIC>(ctx, a);
}

// fold the device counters of the last tree search into stats[0] (distance evaluations) and
// stats[1] (nodes visited)
inline int knnIndexReadStats(mptg_ctx* ctx, KnnIndex& ix, uint64_t* stats) {
    if (!ix.devStats || ix.count == 0 || stats[3] != MPTG_KNN_BVH) return MPTG_OK;
    unsigned long long h[4];
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MPTG_CUDA(ctx, cudaMemcpyAsync(h, ix.devStats, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
#ifdef MPTG_KNN_PROBE
    fprintf(stderr, "[knn probe] leaf visits %llu, inner %llu, leaf visits with a candidate %llu, candidate lanes %llu\n", h[0], h[1], h[2], h[3]);
#endif
    stats[0] += h[0] * 32ull;
    stats[1] = h[0] + h[1];
    MPTG_CUDA(ctx, cudaMemsetAsync(ix.devStats, 0, sizeof h, ctx->stream));
    return MPTG_OK;
}

}  // namespace mptg
