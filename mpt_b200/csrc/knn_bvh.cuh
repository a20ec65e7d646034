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

#include <algorithm>
#include <numeric>

#include "common.cuh"
#include "../../include/mptg/mptg_space.h"
#include "topk.cuh"

namespace mptg {

constexpr int BVH_MAXL = 5;  // 32^5 leaves -> up to 2^30 points
constexpr uint32_t BVH_DEAD = 0xFFFFFFFFu;

struct KnnIndex {
    uint32_t count = 0;    // points covered (a prefix of the store)
    uint32_t nPad = 0;     // leaves * 32
    int top = 0;           // top level (0 = leaves)
    uint32_t nNodes[BVH_MAXL] = {0, 0, 0, 0, 0};
    uint32_t nStride[BVH_MAXL] = {0, 0, 0, 0, 0};  // nodes padded to a multiple of 32
    void* mem = nullptr;   // one block: sorted points, perm, boxes
    size_t memBytes = 0;
    void* spts = nullptr;  // [D][nPad]
    uint32_t* perm = nullptr;
    void* lo[BVH_MAXL] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // [D][nStride[l]]
    void* hi[BVH_MAXL] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned long long* devStats = nullptr;  // [0] leaves visited, [1] inner nodes visited
    uint64_t builds = 0;
};

inline void knnIndexFree(KnnIndex& ix) {
    if (ix.mem) cudaFree(ix.mem);
    if (ix.devStats) cudaFree(ix.devStats);
    ix = KnnIndex();
}

// AUTO policy: the tree pays off once a scan of the whole set costs more than a handful of node
// visits per query; below that the tiled scan wins and needs no build.
inline int knnAutoStrategy(uint32_t size, uint32_t /*Q*/, const KnnIndex& /*ix*/) {
    return size >= 16384 ? MPTG_KNN_BVH : MPTG_KNN_BRUTE;
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
template <typename S>
struct BvhArgs {
    const S* spts;
    uint32_t nPad;
    const uint32_t* perm;
    const S* lo[BVH_MAXL];
    const S* hi[BVH_MAXL];
    uint32_t nNodes[BVH_MAXL];
    uint32_t nStride[BVH_MAXL];
    int top;
    const S* queries;
    uint32_t Q, k;
    S radius;
    uint32_t idxMul, idxAdd;
    uint32_t* idxOut;
    S* distOut;
    uint32_t* countOut;
    unsigned long long* stats;
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

constexpr int BVH_WARPS = 8;

template <typename S, int SHAPE, int KPL>
__global__ void __launch_bounds__(BVH_WARPS * 32) knnBvhKernel(const BvhArgs<S> a) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    S* qsm = reinterpret_cast<S*>(smemRaw);  // [BVH_WARPS][D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = a.sp.D;
    const uint32_t q = blockIdx.x * BVH_WARPS + warp;
    if (q >= a.Q) return;  // warp-uniform; no block-wide barriers below

    S* myq = qsm + warp * D;
    for (int c = lane; c < D; c += 32) myq[c] = a.queries[(size_t)q * D + c];
    __syncwarp();
    S qr[7];
    if (SHAPE == SHAPE_SE3) {
#pragma unroll
        for (int c = 0; c < 7; ++c) qr[c] = myq[c];
    }

    WarpTopK<S, KPL> top;
    top.init(a.k);
    unsigned long long leaves = 0, inner = 0;

    // lower bound of the distance to node `node` of level `lvl`
    auto bound = [&](int lvl, uint32_t node) -> S {
        const S* lo = a.lo[lvl];
        const S* hi = a.hi[lvl];
        const uint32_t st = a.nStride[lvl];
        if (SHAPE == SHAPE_SE3) {
            S dotHi, dotLo;
            {
                const S l = __ldg(lo + node), h = __ldg(hi + node);
                const S v = qr[0];
                dotHi = (v >= S(0) ? h : l) * v;
                dotLo = (v >= S(0) ? l : h) * v;
            }
#pragma unroll
            for (int j = 1; j < 4; ++j) {
                const S l = __ldg(lo + (size_t)j * st + node), h = __ldg(hi + (size_t)j * st + node);
                const S v = qr[j];
                dotHi = fp::fma_(v >= S(0) ? h : l, v, dotHi);
                dotLo = fp::fma_(v >= S(0) ? l : h, v, dotLo);
            }
            S ad = dotHi > -dotLo ? dotHi : -dotLo;
            if (ad > S(1)) ad = S(1);
            if (!(ad > S(0))) ad = S(0);
            S dr = fp::acos01(ad) * (S(1) - S(8) * fp::consts<S>::eps());
            if (a.sp.weighted[0]) dr = dr * a.sp.weight[0];
            S acc = S(0);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const S l = __ldg(lo + (size_t)(4 + j) * st + node), h = __ldg(hi + (size_t)(4 + j) * st + node);
                const S v = qr[4 + j];
                const S e = v < l ? l - v : (v > h ? h - v : S(0));
                acc = (j == 0) ? e * e : fp::fma_(e, e, acc);
            }
            S dt = fp::sqrt_(acc);
            if (a.sp.weighted[1]) dt = dt * a.sp.weight[1];
            return dr + dt;
        } else {
            return dev::boxLowerBound<S>(
                a.sp, [&](int c) { return __ldg(lo + (size_t)c * st + node); },
                [&](int c) { return __ldg(hi + (size_t)c * st + node); }, [&](int c) { return myq[c]; });
        }
    };

    uint32_t key[BVH_MAXL];   // per lane: bound of "my" child at each level, BVH_DEAD when consumed/pruned
    uint32_t base[BVH_MAXL];  // first child index at each level (warp-uniform)
#pragma unroll
    for (int l = 0; l < BVH_MAXL; ++l) {
        key[l] = BVH_DEAD;
        base[l] = 0;
    }
    int cur = a.top;
    {
        const uint32_t node = (uint32_t)lane;
        uint32_t kx = BVH_DEAD;
        if (node < a.nNodes[cur]) kx = boundKey<S>(bound(cur, node));
#pragma unroll
        for (int l = 0; l < BVH_MAXL; ++l)
            if (l == cur) key[l] = kx;
    }
    for (;;) {
        uint32_t mine = BVH_DEAD;
#pragma unroll
        for (int l = 0; l < BVH_MAXL; ++l)
            if (l == cur) mine = key[l];
        const uint32_t best = __reduce_min_sync(FULL_MASK, mine);
        const S thrS = top.kthD < a.radius ? top.kthD : a.radius;
        const bool exhausted = best == BVH_DEAD || __uint_as_float(best) > thrAsFloat<S>(thrS);
        if (exhausted) {
            if (cur == a.top) break;
            ++cur;
            continue;
        }
        const int src = __ffs(__ballot_sync(FULL_MASK, mine == best)) - 1;
        uint32_t b = 0;
#pragma unroll
        for (int l = 0; l < BVH_MAXL; ++l)
            if (l == cur) {
                if (lane == src) key[l] = BVH_DEAD;
                b = base[l];
            }
        const uint32_t node = b + (uint32_t)src;
        if (cur == 0) {
            // leaf: 32 points, one per lane
            ++leaves;
            const uint32_t p = node * 32u + (uint32_t)lane;
            const uint32_t orig = __ldg(a.perm + p);
            const bool have = orig != MPTG_NO_INDEX;
            S dist;
            if (SHAPE == SHAPE_SE3) {
                S pv[7];
#pragma unroll
                for (int c = 0; c < 7; ++c) pv[c] = __ldg(a.spts + (size_t)c * a.nPad + p);
                dist = dev::se3Distance<S>(a.sp.weight[0], a.sp.weighted[0] != 0, a.sp.weight[1], a.sp.weighted[1] != 0, pv, qr);
            } else {
                dist = dev::distance<S>(
                    a.sp, [&](int c) { return __ldg(a.spts + (size_t)c * a.nPad + p); }, [&](int c) { return myq[c]; });
            }
            top.offer(have, dist, orig * a.idxMul + a.idxAdd, a.radius, lane);
        } else {
            ++inner;
            const int child = cur - 1;
            const uint32_t cb = node * 32u;
            const uint32_t cn = cb + (uint32_t)lane;
            uint32_t kx = BVH_DEAD;
            if (cn < a.nNodes[child]) kx = boundKey<S>(bound(child, cn));
#pragma unroll
            for (int l = 0; l < BVH_MAXL; ++l)
                if (l == child) {
                    key[l] = kx;
                    base[l] = cb;
                }
            cur = child;
        }
    }
    const uint32_t count = top.store(a.k, a.idxOut + (size_t)q * a.k, a.distOut + (size_t)q * a.k, lane);
    if (a.countOut && lane == 0) a.countOut[q] = count;
    if (lane == 0 && a.stats) {
        atomicAdd(a.stats + 0, leaves);
        atomicAdd(a.stats + 1, inner);
    }
}

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
    // 3. levels
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
    size_t bytes = 0;
    auto take = [&](size_t b) {
        const size_t o = bytes;
        bytes += (b + 255) & ~(size_t)255;
        return o;
    };
    const size_t oPts = take((size_t)D * nx.nPad * sizeof(S));
    const size_t oPerm = take((size_t)nx.nPad * sizeof(uint32_t));
    size_t oLo[BVH_MAXL], oHi[BVH_MAXL];
    for (int l = 0; l <= nx.top; ++l) {
        nx.nStride[l] = ((nx.nNodes[l] + 31) / 32) * 32;
        oLo[l] = take((size_t)D * nx.nStride[l] * sizeof(S));
        oHi[l] = take((size_t)D * nx.nStride[l] * sizeof(S));
    }
    std::vector<unsigned char> host(bytes, 0);
    S* hp = reinterpret_cast<S*>(host.data() + oPts);
    uint32_t* hperm = reinterpret_cast<uint32_t*>(host.data() + oPerm);
    for (uint32_t i = 0; i < nx.nPad; ++i) {
        const uint32_t src = hb.order[i < n ? i : n - 1];  // padding repeats the last point (perm marks it unused)
        hperm[i] = i < n ? src : MPTG_NO_INDEX;
        for (int c = 0; c < D; ++c) hp[(size_t)c * nx.nPad + i] = hb.pts[(size_t)src * D + c];
    }
    const S inf = fp::consts<S>::inf();
    for (int l = 0; l <= nx.top; ++l) {
        S* lo = reinterpret_cast<S*>(host.data() + oLo[l]);
        S* hi = reinterpret_cast<S*>(host.data() + oHi[l]);
        const uint32_t st = nx.nStride[l];
        for (int c = 0; c < D; ++c)
            for (uint32_t j = 0; j < st; ++j) lo[(size_t)c * st + j] = inf, hi[(size_t)c * st + j] = -inf;
        if (l == 0) {
#pragma omp parallel for schedule(static)
            for (int64_t j = 0; j < (int64_t)nx.nNodes[0]; ++j)
                for (int c = 0; c < D; ++c) {
                    S mn = inf, mx = -inf;
                    for (uint32_t i = (uint32_t)j * 32; i < (uint32_t)j * 32 + 32; ++i) {
                        const S v = hp[(size_t)c * nx.nPad + i];
                        mn = v < mn ? v : mn;
                        mx = v > mx ? v : mx;
                    }
                    lo[(size_t)c * st + j] = mn;
                    hi[(size_t)c * st + j] = mx;
                }
        } else {
            const S* clo = reinterpret_cast<const S*>(host.data() + oLo[l - 1]);
            const S* chi = reinterpret_cast<const S*>(host.data() + oHi[l - 1]);
            const uint32_t cst = nx.nStride[l - 1];
            for (uint32_t j = 0; j < nx.nNodes[l]; ++j)
                for (int c = 0; c < D; ++c) {
                    S mn = inf, mx = -inf;
                    for (uint32_t i = j * 32; i < j * 32 + 32 && i < nx.nNodes[l - 1]; ++i) {
                        mn = clo[(size_t)c * cst + i] < mn ? clo[(size_t)c * cst + i] : mn;
                        mx = chi[(size_t)c * cst + i] > mx ? chi[(size_t)c * cst + i] : mx;
                    }
                    lo[(size_t)c * st + j] = mn;
                    hi[(size_t)c * st + j] = mx;
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
        MPTG_CUDA(ctx, cudaMalloc(&stats, 4 * sizeof(unsigned long long)));
        if (int rc = memsetSync(ctx, stats, 0, 4 * sizeof(unsigned long long))) return rc;
    }
    nx.mem = mem;
    nx.memBytes = memBytes;
    nx.devStats = stats;
    nx.builds = ix.builds + 1;
    nx.spts = (char*)mem + oPts;
    nx.perm = (uint32_t*)((char*)mem + oPerm);
    for (int l = 0; l <= nx.top; ++l) {
        nx.lo[l] = (char*)mem + oLo[l];
        nx.hi[l] = (char*)mem + oHi[l];
    }
    ix = nx;
    return MPTG_OK;
}

// (Re)build when there is no index yet or the unindexed tail has grown past a quarter of the
// indexed prefix; the tail is scanned by brute force in between (knn.cu).
template <typename S>
int knnEnsureIndex(mptg_ctx* ctx, KnnIndex& ix, const mptg_space_desc& sp, const S* pts, uint32_t stride, uint32_t n) {
    if (ix.count != 0 && ix.count <= n && (n - ix.count) <= ix.count / 4) return MPTG_OK;
    return knnBuildIndex<S>(ctx, ix, sp, pts, stride, n);
}

template <typename S, int SHAPE>
int launchBvhShape(mptg_ctx* ctx, const BvhArgs<S>& a) {
    const dim3 grid((a.Q + BVH_WARPS - 1) / BVH_WARPS), block(BVH_WARPS * 32);
    const size_t smem = (size_t)BVH_WARPS * a.sp.D * sizeof(S);
    if (a.k <= 32) knnBvhKernel<S, SHAPE, 1><<<grid, block, smem, ctx->stream>>>(a);
    else if (a.k <= 64) knnBvhKernel<S, SHAPE, 2><<<grid, block, smem, ctx->stream>>>(a);
    else knnBvhKernel<S, SHAPE, 4><<<grid, block, smem, ctx->stream>>>(a);
    MPTG_LAUNCHED(ctx);
    return MPTG_OK;
}

template <typename S>
int knnBvhQuery(mptg_ctx* ctx, KnnIndex& ix, const mptg_space_desc& space, const S* queries, uint32_t Q, uint32_t k,
                double radius, uint32_t idxMul, uint32_t idxAdd, uint32_t* idxOut, S* distOut, uint32_t* countOut,
                uint64_t* /*hostStats*/) {
    BvhArgs<S> a{};
    a.spts = (const S*)ix.spts;
    a.nPad = ix.nPad;
    a.perm = ix.perm;
    for (int l = 0; l < BVH_MAXL; ++l) {
        a.lo[l] = (const S*)ix.lo[l];
        a.hi[l] = (const S*)ix.hi[l];
        a.nNodes[l] = ix.nNodes[l];
        a.nStride[l] = ix.nStride[l];
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
    a.sp = makeDevSpace<S>(space);
    MPTG_CUDA(ctx, cudaMemsetAsync(ix.devStats, 0, 4 * sizeof(unsigned long long), ctx->stream));
    if (classifySpace(space) == SHAPE_SE3) return launchBvhShape<S, SHAPE_SE3>(ctx, a);
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
    stats[0] += h[0] * 32ull;
    stats[1] = h[0] + h[1];
    MPTG_CUDA(ctx, cudaMemsetAsync(ix.devStats, 0, sizeof h, ctx->stream));
    return MPTG_OK;
}

}  // namespace mptg
