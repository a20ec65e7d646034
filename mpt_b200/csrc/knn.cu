// knn.cu -- device-resident batched nearest-neighbour structure.
//
// Replaces nigh::Nigh<Node*,Space,NodeKey,Concurrency,Strategy> as used by the planners
// (src/mpt/impl/prrt/prrt.hpp:121-122,186,406-409,447; impl/prrt_star/prrt_star.hpp:182-183,278,
// 505-508,559-562,619; impl/pprm/pprm.hpp:80-81,304,337; impl/rrg_rewire_neighbors.hpp:65-67,125-128).
//
// Layout in HBM: structure-of-arrays, row c holds scalar c of every stored state
// (pts[c*stride + i]); stride is the capacity rounded to 256 so every row is 1 KB aligned and a warp
// reads 32 consecutive points of a row with one 128 B transaction.
//
// Two exact search engines share one distance function and one (distance, index) total order:
//   * tiled brute force (this file): one warp per query, 8 queries per CTA share point tiles staged
//     in shared memory; per-warp sorted top-k in registers (topk.cuh).  Used for small trees and as
//     the scan of the not-yet-indexed tail.
//   * wide bounding-box tree (knn_bvh.cuh): 32-ary hierarchy over a spatially sorted copy,
//     warp-cooperative traversal.  Used for large trees.
#include "common.cuh"
#include "knn_bvh.cuh"
#include "../../include/mptg/mptg_space.h"
#include "topk.cuh"

namespace mptg {

// ---------------------------------------------------------------------------------------------
// brute force
// ---------------------------------------------------------------------------------------------
template <typename S>
struct BruteArgs {
    const S* pts;
    uint32_t stride;
    uint32_t begin, end;  // scan points [begin, end)
    const S* queries;     // AoS
    uint32_t Q, k;
    S radius;
    uint32_t chunk;  // points per split (multiple of tile)
    uint32_t tile;   // points per shared-memory tile (multiple of 32)
    uint32_t idxMul, idxAdd;
    uint32_t* idxOut;  // [split][Q][k]
    S* distOut;
    uint32_t* countOut;  // [Q], only written when gridDim.y == 1
    unsigned long long* bound;  // [Q] or null: k-th distances published by the splits of a query (knnBruteL1Kernel)
    DevSpace<S> sp;
};

constexpr int BRUTE_WARPS = 8;

template <typename S, int SHAPE, int KPL>
__global__ void __launch_bounds__(BRUTE_WARPS * 32) knnBruteKernel(const BruteArgs<S> a) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    S* tile = reinterpret_cast<S*>(smemRaw);                    // [D][a.tile]
    S* qsm = tile + (size_t)a.sp.D * a.tile;                    // [BRUTE_WARPS][D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = a.sp.D;
    const uint32_t q = blockIdx.x * BRUTE_WARPS + warp;
    const bool active = q < a.Q;

    // query -> shared (generic path reads it from there) and registers (fast paths)
    S* myq = qsm + warp * D;
    if (active)
        for (int c = lane; c < D; c += 32) myq[c] = a.queries[(size_t)q * D + c];
    __syncwarp();
    S qr[7];
    if (SHAPE == SHAPE_SE3) {
#pragma unroll
        for (int c = 0; c < 7; ++c) qr[c] = active ? myq[c] : S(0);
    } else if (SHAPE == SHAPE_L2_2 || SHAPE == SHAPE_L2_3) {
#pragma unroll
        for (int c = 0; c < (SHAPE == SHAPE_L2_2 ? 2 : 3); ++c) qr[c] = active ? myq[c] : S(0);
    }

    WarpTopK<S, KPL> top;
    top.init(a.k);

    // prefilter threshold on the squared translation distance (fast paths): skip when surely d > thr
    S thr2 = fp::consts<S>::inf(), thr1 = fp::consts<S>::inf();
    auto refreshThr = [&]() {
        S thr = top.kthD < a.radius ? top.kthD : a.radius;
        thr1 = thr;
        if (SHAPE == SHAPE_SE3 && a.sp.weighted[1]) thr = fp::div_(thr, a.sp.weight[1]) * (S(1) + S(8) * fp::consts<S>::eps());
        S t2 = thr * thr;
        thr2 = t2 + t2 * (S(16) * fp::consts<S>::eps());  // generous slack: sqrt/round can only shrink d by < 1 ulp
    };
    refreshThr();

    const uint32_t splitBegin = a.begin + blockIdx.y * a.chunk;
    const uint32_t splitEnd = min(a.end, splitBegin + a.chunk);

    for (uint32_t t0 = splitBegin; t0 < splitEnd; t0 += a.tile) {
        const uint32_t cnt = min(a.tile, splitEnd - t0);
        __syncthreads();  // previous tile fully consumed
        for (int c = 0; c < D; ++c) {
            const S* row = a.pts + (size_t)c * a.stride + t0;
            for (uint32_t p = threadIdx.x; p < cnt; p += blockDim.x) tile[(size_t)c * a.tile + p] = __ldg(row + p);
        }
        __syncthreads();
        if (!active) continue;
        for (uint32_t base = 0; base < cnt; base += 32) {
            const uint32_t p = base + lane;
            const bool have = p < cnt;
            const uint32_t pp = have ? p : 0;
            const uint32_t gi = (t0 + pp) * a.idxMul + a.idxAdd;
            S dist;
            bool cand = have;
            if (SHAPE == SHAPE_SE3) {
                const S d0 = tile[4 * a.tile + pp] - qr[4], d1 = tile[5 * a.tile + pp] - qr[5],
                        d2 = tile[6 * a.tile + pp] - qr[6];
                S s = d0 * d0;
                s = fp::fma_(d1, d1, s);
                s = fp::fma_(d2, d2, s);
                cand = cand && s <= thr2;
                if (!__any_sync(FULL_MASK, cand)) continue;
                S dt = fp::sqrt_(s);
                if (a.sp.weighted[1]) dt = dt * a.sp.weight[1];
                S dr = dev::so3Dist<S>(tile[0 * a.tile + pp], tile[1 * a.tile + pp], tile[2 * a.tile + pp],
                                       tile[3 * a.tile + pp], qr[0], qr[1], qr[2], qr[3]);
                if (a.sp.weighted[0]) dr = dr * a.sp.weight[0];
                dist = dr + dt;
            } else if (SHAPE == SHAPE_L2_2 || SHAPE == SHAPE_L2_3) {
                const S d0 = tile[0 * a.tile + pp] - qr[0], d1 = tile[1 * a.tile + pp] - qr[1];
                S s = d0 * d0;
                s = fp::fma_(d1, d1, s);
                if (SHAPE == SHAPE_L2_3) {
                    const S d2 = tile[2 * a.tile + pp] - qr[2];
                    s = fp::fma_(d2, d2, s);
                }
                cand = cand && s <= thr2;
                if (!__any_sync(FULL_MASK, cand)) continue;
                dist = fp::sqrt_(s);
            } else if (SHAPE == SHAPE_L1) {
                // one L1 part of any dimension (the N-link arm's space): the sum in the order of mptg_space.h
                // (partDistance, p == 1), and the sum itself is the prefilter
                S acc = fp::abs_(tile[pp] - myq[0]);
#pragma unroll 4
                for (int c = 1; c < D; ++c) acc = acc + fp::abs_(tile[(size_t)c * a.tile + pp] - myq[c]);
                cand = cand && acc <= thr1;
                if (!__any_sync(FULL_MASK, cand)) continue;
                dist = acc;
            } else {
                dist = dev::distance<S>(
                    a.sp, [&](int c) { return tile[(size_t)c * a.tile + pp]; }, [&](int c) { return myq[c]; });
            }
            const S oldKth = top.kthD;
            top.offer(cand, dist, gi, a.radius, lane);
            if (SHAPE != SHAPE_GENERIC && top.kthD != oldKth) refreshThr();
        }
    }
    if (active) {
        const size_t row = ((size_t)blockIdx.y * a.Q + q) * a.k;
        const uint32_t count = top.store(a.k, a.idxOut + row, a.distOut + row, lane);
        if (gridDim.y == 1 && a.countOut && lane == 0) a.countOut[q] = count;
    }
}

// ---------------------------------------------------------------------------------------------
// scan, one unweighted L1 part (the N-link arm's space): R queries x P points per lane
// ---------------------------------------------------------------------------------------------
// knnBruteKernel<S, SHAPE_L1> is bound by shared-memory wavefronts, not arithmetic: per coordinate and 32 pairs it
// issues one tile load and one broadcast load of the query's coordinate for two adds, and a broadcast load costs as
// many wavefronts as a strided one of the same width (measured: doubles, one query per warp 65 clk per SM
// sub-partition for 4 x 32 coordinate pairs; four queries per warp 40; model 4 wavefronts per 128-bit load: 64 / 40).
// Here a lane owns P consecutive-by-vector points and a warp owns R queries: per coordinate, P / V 128-bit tile
// loads (V = 16 / sizeof(S) points each) and R / V 128-bit broadcast loads of the [D][R] query block feed 2 R P adds
// per lane, which brings the load wavefronts down to the cost of the adds.  Sums run in the order of mptg_space.h
// (partDistance, p == 1), so distances are the same bits as everywhere else; candidates may be offered in any
// order because the list order is the total order (distance, index).
template <typename S, int N>
struct alignas(sizeof(S) * N > 16 ? 16 : sizeof(S) * N) ScalarVec {
    S v[N];
};

template <int BYTES>
__device__ __forceinline__ void cpAsync(void* smemDst, const void* src, int srcBytes) {  // the rest of BYTES is zero-filled
    const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(d), "l"(src), "n"(BYTES), "r"(srcBytes) : "memory");
}

template <typename S> __device__ __forceinline__ S loadBound(const unsigned long long* p);
template <> __device__ __forceinline__ double loadBound<double>(const unsigned long long* p) {
    return __longlong_as_double((long long)*reinterpret_cast<const volatile unsigned long long*>(p));  // all ones (unset) = NaN: compares false
}
template <> __device__ __forceinline__ float loadBound<float>(const unsigned long long* p) {
    return __uint_as_float(*reinterpret_cast<const volatile unsigned int*>(p));
}
template <typename S> __device__ __forceinline__ void publishBound(unsigned long long* p, S d);
template <> __device__ __forceinline__ void publishBound<double>(unsigned long long* p, double d) {
    atomicMin(p, (unsigned long long)__double_as_longlong(d));
}
template <> __device__ __forceinline__ void publishBound<float>(unsigned long long* p, float d) {
    atomicMin(reinterpret_cast<unsigned int*>(p), __float_as_uint(d));
}

constexpr int L1_R = 4;    // queries per warp
constexpr int L1_PTS = 4;  // points per lane and step: a step covers 128 points

template <typename S, int KPL>
__global__ void __launch_bounds__(BRUTE_WARPS * 32) knnBruteL1Kernel(const BruteArgs<S> a) {
    constexpr int R = L1_R, P = L1_PTS;
    constexpr int V = 16 / (int)sizeof(S) < P ? 16 / (int)sizeof(S) : P;  // points per 128-bit load
    constexpr int G = P / V;                                              // loads per coordinate
    constexpr int STEP = 32 * P;
    extern __shared__ __align__(16) unsigned char smemRaw[];
    S* tiles = reinterpret_cast<S*>(smemRaw);          // [2][D][a.tile], a.tile a multiple of STEP
    S* qsm = tiles + 2 * (size_t)a.sp.D * a.tile;      // [BRUTE_WARPS][D][R]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = a.sp.D;
    const uint32_t q0 = (blockIdx.x * BRUTE_WARPS + warp) * R;
    const bool active = q0 < a.Q;

    S* myq = qsm + (size_t)warp * D * R;
    for (int e = lane; e < D * R; e += 32) {
        const int c = e / R, r = e % R;
        myq[e] = q0 + r < a.Q ? a.queries[(size_t)(q0 + r) * D + c] : S(0);
    }
    __syncwarp();
    const ScalarVec<S, R>* qv = reinterpret_cast<const ScalarVec<S, R>*>(myq);

    WarpTopK<S, KPL> top[R];
    S thr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        top[r].init(a.k);
        thr[r] = a.radius;
    }

    const uint32_t splitBegin = a.begin + blockIdx.y * a.chunk;
    const uint32_t splitEnd = min(a.end, splitBegin + a.chunk);

    // Tiles are filled with cp.async into two buffers, the next one while the current one is scanned (filled
    // synchronously the fill was 40-50 % of the kernel; the one-query-per-warp kernel above was tried with the same
    // scheme and gained nothing on SE(3) / planar sets -- its fills are 2-7 rows against a full distance per pair).  16-byte copies when every row of the tile starts on a
    // 16-byte boundary, element copies otherwise; the rest of the last 128-point step is zero-filled and masked below.
    constexpr int E = 16 / (int)sizeof(S);
    const bool rows16 = (a.stride % E) == 0 && (splitBegin % E) == 0 && (reinterpret_cast<uintptr_t>(a.pts) & 15) == 0;
    auto fill = [&](S* buf, uint32_t t0) {
        const uint32_t cnt = min(a.tile, splitEnd - t0);
        const uint32_t cntPad = (cnt + STEP - 1) / STEP * STEP;
        if (rows16) {
            const uint32_t chunks = cntPad / E, total = chunks * (uint32_t)D;
            for (uint32_t x = threadIdx.x; x < total; x += blockDim.x) {
                const uint32_t c = x / chunks, j = (x - c * chunks) * E;
                const int bytes = j + E <= cnt ? 16 : j < cnt ? (int)((cnt - j) * sizeof(S)) : 0;
                cpAsync<16>(buf + (size_t)c * a.tile + j, bytes ? a.pts + (size_t)c * a.stride + t0 + j : a.pts, bytes);
            }
        } else {
            const uint32_t total = cntPad * (uint32_t)D;
            for (uint32_t x = threadIdx.x; x < total; x += blockDim.x) {
                const uint32_t c = x / cntPad, j = x - c * cntPad;
                cpAsync<(int)sizeof(S)>(buf + (size_t)c * a.tile + j, j < cnt ? a.pts + (size_t)c * a.stride + t0 + j : a.pts,
                                        j < cnt ? (int)sizeof(S) : 0);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };

    if (splitBegin < splitEnd) fill(tiles, splitBegin);
    uint32_t which = 0;
    for (uint32_t t0 = splitBegin; t0 < splitEnd; t0 += a.tile, which ^= 1) {
        const uint32_t cnt = min(a.tile, splitEnd - t0);
        const S* tile = tiles + (size_t)which * D * a.tile;
        if (t0 + a.tile < splitEnd) {  // the other buffer was released by the barrier that ended the previous iteration
            fill(tiles + (size_t)(which ^ 1) * D * a.tile, t0 + a.tile);
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        }
        __syncthreads();  // every thread's copies of this tile have landed
        if (active && a.bound) {
            // split scans: the k-th distance of any split's full list bounds the k-th distance of the union, so the
            // splits of a query publish theirs (bit patterns of non-negative numbers order like unsigned integers) and
            // prune with the smallest one seen; "<=" keeps ties, which the index order decides in the merge
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (q0 + r < a.Q) {
                    const S g = loadBound<S>(a.bound + q0 + r);
                    thr[r] = g < thr[r] ? g : thr[r];
                }
        }
        if (active)
        for (uint32_t base = 0; base < cnt; base += STEP) {
            S acc[R][P];
            {
                const ScalarVec<S, R> qc = qv[0];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const ScalarVec<S, V> t = *reinterpret_cast<const ScalarVec<S, V>*>(tile + base + g * 32 * V + lane * V);
#pragma unroll
                    for (int j = 0; j < V; ++j)
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r][g * V + j] = fp::abs_(t.v[j] - qc.v[r]);
                }
            }
#pragma unroll 2
            for (int c = 1; c < D; ++c) {
                const ScalarVec<S, R> qc = qv[c];
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const ScalarVec<S, V> t =
                        *reinterpret_cast<const ScalarVec<S, V>*>(tile + (size_t)c * a.tile + base + g * 32 * V + lane * V);
#pragma unroll
                    for (int j = 0; j < V; ++j)
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r][g * V + j] = acc[r][g * V + j] + fp::abs_(t.v[j] - qc.v[r]);
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool qok = q0 + r < a.Q;
                // no candidate in the whole step: the common case once the list is full
                bool some = false;
#pragma unroll
                for (int x = 0; x < P; ++x) some = some || acc[r][x] <= thr[r];
                if (!__any_sync(FULL_MASK, qok && some)) continue;
#pragma unroll
                for (int g = 0; g < G; ++g)
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        const uint32_t p = base + g * 32 * V + lane * V + j;
                        const bool cand = qok && p < cnt && acc[r][g * V + j] <= thr[r];
                        if (__any_sync(FULL_MASK, cand)) {
                            top[r].offer(cand, acc[r][g * V + j], (t0 + p) * a.idxMul + a.idxAdd, a.radius, lane);
                            const S mine = top[r].kthD < a.radius ? top[r].kthD : a.radius;
                            if (a.bound && lane == 0 && top[r].kthD < fp::consts<S>::inf()) publishBound<S>(a.bound + q0 + r, top[r].kthD);
                            thr[r] = mine < thr[r] ? mine : thr[r];
                        }
                    }
            }
        }
        __syncthreads();  // this tile fully consumed: its buffer is refilled in the next iteration
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t q = q0 + r;
        if (q < a.Q) {
            const size_t row = ((size_t)blockIdx.y * a.Q + q) * a.k;
            const uint32_t count = top[r].store(a.k, a.idxOut + row, a.distOut + row, lane);
            if (gridDim.y == 1 && a.countOut && lane == 0) a.countOut[q] = count;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// merge of `parts` sorted candidate lists per query (split-N brute force, tail + index, multi-GPU)
// ---------------------------------------------------------------------------------------------
template <typename S, int KPL>
__global__ void __launch_bounds__(256) knnMergeKernel(uint32_t parts, uint32_t Q, uint32_t k, const uint32_t* idxIn,
                                                      const S* distIn, uint32_t* idxOut, S* distOut,
                                                      uint32_t* countOut) {
    const int lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Q) return;
    WarpTopK<S, KPL> top;
    top.init(k);
    for (uint32_t part = 0; part < parts; ++part) {
        const size_t row = ((size_t)part * Q + q) * k;
        for (uint32_t base = 0; base < k; base += 32) {
            const uint32_t j = base + lane;
            const bool have = j < k;
            const S d = have ? distIn[row + j] : fp::consts<S>::inf();
            const uint32_t i = have ? idxIn[row + j] : MPTG_NO_INDEX;
            top.offer(have && i != MPTG_NO_INDEX, d, i, fp::consts<S>::inf(), lane);
        }
    }
    const uint32_t count = top.store(k, idxOut + (size_t)q * k, distOut + (size_t)q * k, lane);
    if (countOut && lane == 0) countOut[q] = count;
}

// The same merge for up to 64 lists, using that every list is ASCENDING in (distance, index) with its empty slots at the
// end (the contract of mptg_knn_merge_dev and what every search stores): lane p holds the head of list p (and of list
// p + 32), one step picks the smallest head of the warp -- three or two REDUX over the key words, the bit patterns of
// non-negative distances order like unsigned integers -- and only the winner advances, its next entry already loaded.
// k steps of ~100 cycles, where inserting parts x k candidates one by one into a WarpTopK costs ~500 cycles each
// (16 lists of 36: 63 us -> a few; profiles/r2_launches_ramp_prrtstar_queued.txt).
__device__ __forceinline__ int warpArgMin(float d, uint32_t i, bool& none) {
    const uint32_t ub = __float_as_uint(d + 0.0f);
    const uint32_t m = __reduce_min_sync(FULL_MASK, ub);
    const uint32_t mi = __reduce_min_sync(FULL_MASK, ub == m ? i : MPTG_NO_INDEX);
    none = mi == MPTG_NO_INDEX;
    return __ffs(__ballot_sync(FULL_MASK, ub == m && i == mi)) - 1;
}
__device__ __forceinline__ int warpArgMin(double d, uint32_t i, bool& none) {
    const unsigned long long ub = (unsigned long long)__double_as_longlong(d + 0.0);
    const uint32_t hi = (uint32_t)(ub >> 32), lo = (uint32_t)ub;
    const uint32_t mh = __reduce_min_sync(FULL_MASK, hi);
    const uint32_t ml = __reduce_min_sync(FULL_MASK, hi == mh ? lo : 0xFFFFFFFFu);
    const bool at = hi == mh && lo == ml;
    const uint32_t mi = __reduce_min_sync(FULL_MASK, at ? i : MPTG_NO_INDEX);
    none = mi == MPTG_NO_INDEX;
    return __ffs(__ballot_sync(FULL_MASK, at && i == mi)) - 1;
}
template <typename S>
__global__ void __launch_bounds__(256) knnMergeHeadsKernel(uint32_t parts, uint32_t Q, uint32_t k, const uint32_t* __restrict__ idxIn,
                                                           const S* __restrict__ distIn, uint32_t* __restrict__ idxOut, S* __restrict__ distOut,
                                                           uint32_t* __restrict__ countOut) {
    const int lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= Q) return;
    const S inf = fp::consts<S>::inf();
    const bool liveA = (uint32_t)lane < parts, liveB = (uint32_t)lane + 32u < parts;
    const size_t rowA = ((size_t)(liveA ? lane : 0) * Q + q) * k, rowB = ((size_t)(liveB ? lane + 32 : 0) * Q + q) * k;
    auto load = [&](bool live, size_t row, uint32_t pos, S& d, uint32_t& i) {
        const bool have = live && pos < k;
        d = have ? distIn[row + pos] : inf;
        i = have ? idxIn[row + pos] : MPTG_NO_INDEX;
        if (i == MPTG_NO_INDEX) d = inf;  // an empty slot: nothing follows it in its list
    };
    S dA, dB, ndA, ndB;
    uint32_t iA, iB, niA, niB, posA = 2, posB = 2;
    load(liveA, rowA, 0, dA, iA), load(liveA, rowA, 1, ndA, niA);
    load(liveB, rowB, 0, dB, iB), load(liveB, rowB, 1, ndB, niB);
    S outD = inf;
    uint32_t outI = MPTG_NO_INDEX;
    uint32_t t = 0;
    for (; t < k; ++t) {
        const bool useB = dB < dA || (dB == dA && iB < iA);
        const S d = useB ? dB : dA;
        const uint32_t i = useB ? iB : iA;
        bool none;
        const int src = warpArgMin(d, i, none);
        if (none) break;  // every head is an empty slot
        const S wd = __shfl_sync(FULL_MASK, d, src);
        const uint32_t wi = __shfl_sync(FULL_MASK, i, src);
        if ((t & 31u) == (uint32_t)lane) outD = wd, outI = wi;
        if (lane == src) {
            if (useB) {
                dB = ndB, iB = niB;
                load(liveB, rowB, posB++, ndB, niB);
            } else {
                dA = ndA, iA = niA;
                load(liveA, rowA, posA++, ndA, niA);
            }
        }
        if ((t & 31u) == 31u) {  // 32 results: one coalesced store
            idxOut[(size_t)q * k + (t - 31u) + lane] = outI;
            distOut[(size_t)q * k + (t - 31u) + lane] = outD;
            outD = inf, outI = MPTG_NO_INDEX;
        }
    }
    // the results of the last, partial group of 32 and the empty slots behind them
    const uint32_t base = t & ~31u;
    for (uint32_t j = base + lane; j < k; j += 32) {
        const bool mine = j < t;  // only the first round of this loop can hold results
        idxOut[(size_t)q * k + j] = mine ? outI : MPTG_NO_INDEX;
        distOut[(size_t)q * k + j] = mine ? outD : inf;
    }
    if (countOut && lane == 0) countOut[q] = t;
}

// AoS -> SoA append
template <typename S>
__global__ void knnScatterKernel(const S* aos, uint32_t count, int D, S* pts, uint32_t stride, uint32_t first) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // count * D exceeds 32 bits from 2^26 points of 64 scalars on
    if (t >= (size_t)count * D) return;
    const uint32_t i = (uint32_t)(t / D), c = (uint32_t)(t % D);
    pts[(size_t)c * stride + first + i] = aos[t];
}
template <typename S>
__global__ void knnGatherKernel(const S* pts, uint32_t stride, uint32_t first, uint32_t count, int D, S* aos) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)count * D) return;
    const uint32_t i = (uint32_t)(t / D), c = (uint32_t)(t % D);
    aos[t] = pts[(size_t)c * stride + first + i];
}

}  // namespace mptg

using namespace mptg;

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
struct mptg_knn {
    mptg_ctx* ctx = nullptr;
    mptg_space_desc space{};
    int D = 0;
    int scalar = MPTG_F32;
    SpaceShape shape = SHAPE_GENERIC;
    uint32_t capacity = 0, stride = 0, size = 0;
    void* pts = nullptr;  // SoA [D][stride]
    int strategy = MPTG_KNN_AUTO;
    uint32_t idxMul = 1, idxAdd = 0;
    uint32_t* gid = nullptr;  // optional [capacity]: reported (global) index per stored point (mptg_knn_insert_ids)
    uint64_t stats[4] = {0, 0, 0, 0};
    KnnIndex index;  // knn_bvh.cuh
    KnnTail tail;    // knn_index.cuh: Morton-sorted leaves over the points inserted since the tree was built
    bool sawKnn = false;  // a search with k > 1 has been made on this set (PRRT*, PPRM: every wave will make another)
};

namespace {

template <typename S>
int launchMerge(mptg_ctx* ctx, uint32_t parts, uint32_t Q, uint32_t k, const uint32_t* idxIn, const S* distIn,
                uint32_t* idxOut, S* distOut, uint32_t* countOut) {
    if (Q == 0) return MPTG_OK;
    const dim3 grid((Q + 7) / 8), block(256);
    static const bool insertionMerge = getenv("MPTG_KNN_INSERTION_MERGE") != nullptr;  // the earlier kernel, for comparisons
    if (parts <= 64 && (parts >= 4 || k > 32) && !insertionMerge) {  // (two or three short lists: the insertions are as quick)
        knnMergeHeadsKernel<S><<<grid, block, 0, ctx->stream>>>(parts, Q, k, idxIn, distIn, idxOut, distOut, countOut);
        MPTG_LAUNCHED(ctx);
        return MPTG_OK;
    }
    if (k <= 32) knnMergeKernel<S, 1><<<grid, block, 0, ctx->stream>>>(parts, Q, k, idxIn, distIn, idxOut, distOut, countOut);
    else if (k <= 64) knnMergeKernel<S, 2><<<grid, block, 0, ctx->stream>>>(parts, Q, k, idxIn, distIn, idxOut, distOut, countOut);
    else knnMergeKernel<S, 4><<<grid, block, 0, ctx->stream>>>(parts, Q, k, idxIn, distIn, idxOut, distOut, countOut);
    MPTG_LAUNCHED(ctx);
    return MPTG_OK;
}

template <typename S, int SHAPE>
int launchBruteShape(mptg_ctx* ctx, const BruteArgs<S>& a, dim3 grid, size_t smem) {
    const dim3 block(BRUTE_WARPS * 32);
#define MPTG_BRUTE(KPL)                                                                                              \
    do {                                                                                                             \
        if (smem > 48 * 1024)                                                                                        \
            MPTG_CUDA(ctx, cudaFuncSetAttribute(knnBruteKernel<S, SHAPE, KPL>,                                       \
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
        knnBruteKernel<S, SHAPE, KPL><<<grid, block, smem, ctx->stream>>>(a);                                        \
    } while (0)
    if (a.k <= 32) MPTG_BRUTE(1);
    else if (a.k <= 64) MPTG_BRUTE(2);
    else MPTG_BRUTE(4);
#undef MPTG_BRUTE
    MPTG_LAUNCHED(ctx);
    return MPTG_OK;
}

template <typename S>
int launchBruteL1(mptg_ctx* ctx, const BruteArgs<S>& a, dim3 grid, size_t smem) {
    const dim3 block(BRUTE_WARPS * 32);
#define MPTG_BRUTE_L1(KPL)                                                                                           \
    do {                                                                                                             \
        if (smem > 48 * 1024)                                                                                        \
            MPTG_CUDA(ctx, cudaFuncSetAttribute(knnBruteL1Kernel<S, KPL>,                                      \
                                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
        knnBruteL1Kernel<S, KPL><<<grid, block, smem, ctx->stream>>>(a);                                       \
    } while (0)
    if (a.k <= 32) MPTG_BRUTE_L1(1);
    else MPTG_BRUTE_L1(2);  // k <= 64 (the caller keeps larger k on the one-query kernel: four lists of 128 slots spill)
#undef MPTG_BRUTE_L1
    MPTG_LAUNCHED(ctx);
    return MPTG_OK;
}

// Scan points [begin,end) for all queries; results (k per query) to idxOut/distOut/countOut (device).
template <typename S>
int bruteScan(mptg_knn* knn, uint32_t begin, uint32_t end, const S* queries, uint32_t Q, uint32_t k, double radius,
              uint32_t* idxOut, S* distOut, uint32_t* countOut) {
    mptg_ctx* ctx = knn->ctx;
    const int D = knn->D;
    // tile: <= 32 KB of shared memory, multiple of 32 points, <= 1024
    uint32_t tile = (uint32_t)((32 * 1024) / (D * sizeof(S)));
    tile = tile > 1024 ? 1024 : (tile / 32) * 32;
    if (tile < 32) tile = 32;
    const uint32_t n = end - begin;
    // the L1 kernel with several queries per warp and points per lane, once a wave is large enough to fill the machine
    // with it; its tiles are multiples of a 128-point step
    constexpr uint32_t l1Step = 32 * L1_PTS;
    const bool multiQuery = knn->shape == SHAPE_L1 && k <= 64 && Q >= (uint32_t)(BRUTE_WARPS * L1_R) * 16 &&
                            2 * (size_t)D * l1Step * sizeof(S) <= 160 * 1024;  // two tile buffers
    if (multiQuery) tile = tile < l1Step ? l1Step : (tile / l1Step) * l1Step;
    const uint32_t queriesPerCta = multiQuery ? BRUTE_WARPS * L1_R : BRUTE_WARPS;
    const uint32_t qBlocks = (Q + queriesPerCta - 1) / queriesPerCta;
    // split the scan so that the grid fills the machine (~4 CTAs per SM) when there are few queries
    uint32_t splits = 1;
    const uint32_t wantCtas = (uint32_t)ctx->smCount * 4;
    if (!multiQuery && qBlocks < wantCtas && n > 1024) {
        // ... and over a small set (a young planner tree: a few hundred queries, a few thousand points) in pieces shorter than
        // a full tile, up to 16 of them: one warp's pass over 1,024 points with its insertions is a serial chain of ~50 us for
        // k > 32 -- the whole cost of such a wave
        uint32_t ws = (wantCtas + qBlocks - 1) / qBlocks;
        if (ws > 16) ws = 16;
        uint32_t per = (((n + ws - 1) / ws + 31) / 32) * 32;
        if (per < 128) per = 128;
        if (per < tile) tile = per;
    }
    if (qBlocks < wantCtas) {
        splits = (wantCtas + qBlocks - 1) / qBlocks;
        const uint32_t maxSplits = (n + tile - 1) / tile;
        if (splits > maxSplits) splits = maxSplits ? maxSplits : 1;
        if (splits > 64) splits = 64;
    }
    uint32_t chunk = (n + splits - 1) / splits;
    chunk = ((chunk + tile - 1) / tile) * tile;
    if (chunk == 0) chunk = tile;
    splits = n ? (n + chunk - 1) / chunk : 1;

    BruteArgs<S> a{};
    a.pts = (const S*)knn->pts;
    a.stride = knn->stride;
    a.begin = begin;
    a.end = end;
    a.queries = queries;
    a.Q = Q;
    a.k = k;
    a.radius = (radius >= 0 && radius == radius) ? (S)radius : fp::consts<S>::inf();
    a.chunk = chunk;
    a.tile = tile;
    a.idxMul = knn->idxMul;
    a.idxAdd = knn->idxAdd;
    a.sp = makeDevSpace<S>(knn->space);
    a.countOut = countOut;
    uint32_t* scrIdx = nullptr;
    S* scrDist = nullptr;
    if (multiQuery && splits > 1) {
        void* pb;
        int rcb = scratch(ctx, 9, (size_t)Q * sizeof(unsigned long long), &pb);
        if (rcb) return rcb;
        MPTG_CUDA(ctx, cudaMemsetAsync(pb, 0xFF, (size_t)Q * sizeof(unsigned long long), ctx->stream));
        a.bound = (unsigned long long*)pb;
    }
    if (splits > 1) {
        void* p0;
        void* p1;
        int rc = scratch(ctx, 2, (size_t)splits * Q * k * sizeof(uint32_t), &p0);
        if (rc) return rc;
        rc = scratch(ctx, 3, (size_t)splits * Q * k * sizeof(S), &p1);
        if (rc) return rc;
        scrIdx = (uint32_t*)p0;
        scrDist = (S*)p1;
        a.idxOut = scrIdx;
        a.distOut = scrDist;
    } else {
        a.idxOut = idxOut;
        a.distOut = distOut;
    }
    const size_t smem = ((multiQuery ? 2 : 1) * (size_t)D * tile + (size_t)queriesPerCta * D) * sizeof(S);
    const dim3 grid(qBlocks, splits);
    int rc;
    switch (knn->shape) {
        case SHAPE_SE3: rc = launchBruteShape<S, SHAPE_SE3>(ctx, a, grid, smem); break;
        case SHAPE_L2_2: rc = launchBruteShape<S, SHAPE_L2_2>(ctx, a, grid, smem); break;
        case SHAPE_L2_3: rc = launchBruteShape<S, SHAPE_L2_3>(ctx, a, grid, smem); break;
        case SHAPE_L1: rc = multiQuery ? launchBruteL1<S>(ctx, a, grid, smem) : launchBruteShape<S, SHAPE_L1>(ctx, a, grid, smem); break;
        default: rc = launchBruteShape<S, SHAPE_GENERIC>(ctx, a, grid, smem); break;
    }
    if (rc) return rc;
    if (splits > 1) rc = launchMerge<S>(ctx, splits, Q, k, scrIdx, scrDist, idxOut, distOut, countOut);
    knn->stats[0] += (uint64_t)n * Q;
    return rc;
}

template <typename S>
int queryDevT(mptg_knn* knn, const S* queries, uint32_t Q, uint32_t k, double radius, uint32_t* idxOut, S* distOut,
              uint32_t* countOut) {
    mptg_ctx* ctx = knn->ctx;
    knn->stats[0] = knn->stats[1] = 0;
    int strategy = knn->strategy;
    if (strategy == MPTG_KNN_AUTO) strategy = knnAutoStrategy(knn->size, Q, (int)knn->shape, knn->D);
    uint32_t indexed = 0;
    static const bool debug = getenv("MPTG_DEBUG_KNN") != nullptr;
    if (strategy == MPTG_KNN_BVH) {
        const uint64_t before = knn->index.builds;
        int rc = knnEnsureIndex<S>(ctx, knn->index, knn->space, (const S*)knn->pts, knn->stride, knn->size);
        if (rc) return rc;
        indexed = knn->index.count;
        if (debug && knn->index.builds != before) fprintf(stderr, "[knn] rebuilt the tree over %u points (build %llu)\n", indexed, (unsigned long long)knn->index.builds);
    }
    if (debug) fprintf(stderr, "[knn] query Q=%u k=%u size=%u strategy=%d indexed=%u tail leaves=%u covered=%u\n", Q, k, knn->size, strategy, indexed, knn->tail.nLeaves, knn->tail.covered);
    knn->stats[2] = indexed;
    knn->stats[3] = (uint64_t)strategy;
    if (indexed == 0) return bruteScan<S>(knn, 0, knn->size, queries, Q, k, radius, idxOut, distOut, countOut);

    // indexed prefix via the tree; the tail through its Morton-sorted leaves (bounded by the tree's k-th distances)
    // and, for what has arrived since, by brute force; then merge the lists
    KnnTail& tl = knn->tail;
    if (tl.forBuild != knn->index.builds || tl.base != indexed) {  // the tree was rebuilt: start a new tail
        tl.forBuild = knn->index.builds;
        tl.base = indexed;
        tl.covered = 0;
        tl.nLeaves = 0;
    }
    const uint32_t rawBegin = indexed + tl.covered;
    // 1-NN waves (PRRT) scan a raw tail faster than they could sort it into leaves: the exhaustive 1-NN scan costs
    // ~3 us per 1,000 tail points and 8,192 queries, a chunk ~100 us to build; k-NN waves (PRRT*, PPRM) are the other way
    // A set that is also searched with k > 1 (PRRT*: 1-NN then k-NN in every wave) sorts its new points at the first search.
    if (k > 1) knn->sawKnn = true;
    if (knn->size - rawBegin >= TAIL_MIN_CHUNK && (knn->sawKnn || knn->size - rawBegin >= 32768u)) {
        if (int rc = knnTailAppend(ctx, tl, knn->space, (const S*)knn->pts, knn->stride, rawBegin, knn->size - rawBegin)) return rc;
    }
    const bool useLeaves = tl.nLeaves > 0;
    const bool raw = indexed + tl.covered < knn->size;
    const uint32_t parts = 1u + (useLeaves ? 1u : 0u) + (raw ? 1u : 0u);
    uint32_t* i0 = idxOut;
    S* d0 = distOut;
    if (parts > 1) {
        void* p0;
        void* p1;
        int rc = scratch(ctx, 4, (size_t)3 * Q * k * sizeof(uint32_t), &p0);
        if (rc) return rc;
        rc = scratch(ctx, 5, (size_t)3 * Q * k * sizeof(S), &p1);
        if (rc) return rc;
        i0 = (uint32_t*)p0;
        d0 = (S*)p1;
    }
    int rc = knnBvhQuery<S>(ctx, knn->index, knn->space, queries, Q, k, radius, knn->idxMul, knn->idxAdd, i0, d0,
                            parts > 1 ? nullptr : countOut, knn->stats);
    if (rc) return rc;
    uint32_t slot = 1;
    if (useLeaves) {
        BvhArgs<S> a{};
        a.leafPts = (const S*)tl.leafPts, a.perm = tl.perm, a.box[0] = (const S*)tl.box, a.nNodes[0] = tl.nLeaves, a.top = 0;
        a.queries = queries, a.Q = Q, a.k = k;
        a.radius = (radius >= 0 && radius == radius) ? (S)radius : fp::consts<S>::inf();
        a.idxMul = knn->idxMul, a.idxAdd = knn->idxAdd;
        a.idxOut = i0 + (size_t)slot * Q * k, a.distOut = d0 + (size_t)slot * Q * k;
        a.sp = makeDevSpace<S>(knn->space);
        const dim3 grid((Q + BVH_WARPS - 1) / BVH_WARPS), block(BVH_WARPS * 32);
        const size_t smem = (size_t)BVH_WARPS * knn->D * sizeof(S);
        const bool se3 = knn->shape == SHAPE_SE3;
#define MPTG_TAIL(KPL)                                                                                      \
    do {                                                                                                    \
        if (se3) knnTailKernel<S, SHAPE_SE3, KPL><<<grid, block, smem, ctx->stream>>>(a, d0);               \
        else knnTailKernel<S, SHAPE_GENERIC, KPL><<<grid, block, smem, ctx->stream>>>(a, d0);               \
    } while (0)
        if (k <= 32) MPTG_TAIL(1);
        else if (k <= 64) MPTG_TAIL(2);
        else MPTG_TAIL(4);
#undef MPTG_TAIL
        MPTG_LAUNCHED(ctx);
        knn->stats[0] += (uint64_t)tl.nLeaves * Q;  // box tests (an upper bound of the work; leaf visits are not counted)
        ++slot;
    }
    if (raw) {
        rc = bruteScan<S>(knn, indexed + tl.covered, knn->size, queries, Q, k, radius, i0 + (size_t)slot * Q * k, d0 + (size_t)slot * Q * k, nullptr);
        if (rc) return rc;
    }
    if (parts > 1) rc = launchMerge<S>(ctx, parts, Q, k, i0, d0, idxOut, distOut, countOut);
    return rc;
}

template <typename S>
int insertDevT(mptg_knn* knn, const S* aosDev, uint32_t count) {
    const size_t total = (size_t)count * knn->D;
    knnScatterKernel<S><<<(unsigned)((total + 255) / 256), 256, 0, knn->ctx->stream>>>(aosDev, count, knn->D, (S*)knn->pts,
                                                                             knn->stride, knn->size);
    MPTG_LAUNCHED(knn->ctx);
    return MPTG_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int mptg_knn_create(mptg_ctx* ctx, const mptg_space_desc* space, uint32_t capacity, mptg_knn** out) {
    if (!ctx || !space || !out || capacity == 0) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_knn_create: bad argument");
    const int D = spaceScalars(space);
    if (D <= 0) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_knn_create: malformed space descriptor");
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    auto* k = new mptg_knn();
    k->ctx = ctx;
    k->space = *space;
    k->D = D;
    k->scalar = space->scalar;
    k->shape = classifySpace(*space);
    k->capacity = capacity;
    k->stride = ((capacity + 255u) / 256u) * 256u;
    k->index.capacityHint = capacity;
    cudaError_t e = cudaMalloc(&k->pts, (size_t)D * k->stride * space->scalar);
    if (e != cudaSuccess) {
        delete k;
        return fail(ctx, e == cudaErrorMemoryAllocation ? MPTG_ERR_OOM : MPTG_ERR_CUDA, "mptg_knn_create: cudaMalloc(%zu) failed: %s",
                    (size_t)D * k->stride * space->scalar, cudaGetErrorString(e));
    }
    *out = k;
    return MPTG_OK;
}

int mptg_knn_destroy(mptg_knn* knn) {
    if (!knn) return MPTG_OK;
    cudaSetDevice(knn->ctx->device);
    cudaStreamSynchronize(knn->ctx->stream);
    knnIndexFree(knn->index);
    if (knn->tail.mem) cudaFree(knn->tail.mem);
    cudaFree(knn->gid);
    cudaFree(knn->pts);
    delete knn;
    return MPTG_OK;
}

int mptg_knn_set_strategy(mptg_knn* knn, int strategy) {
    if (!knn || strategy < MPTG_KNN_AUTO || strategy > MPTG_KNN_BVH) return fail(knn ? knn->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_knn_set_strategy: bad argument");
    knn->strategy = strategy;
    return MPTG_OK;
}

int mptg_knn_set_index_map(mptg_knn* knn, uint32_t mul, uint32_t add) {
    if (!knn || mul == 0) return fail(knn ? knn->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_knn_set_index_map: bad argument");
    knn->idxMul = mul;
    knn->idxAdd = add;
    return MPTG_OK;
}

uint32_t mptg_knn_size(const mptg_knn* knn) { return knn ? knn->size : 0; }

int mptg_knn_insert_dev(mptg_knn* knn, const void* statesDev, uint32_t count, uint32_t* firstOut) {
    if (!knn || (!statesDev && count)) return fail(knn ? knn->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_knn_insert: bad argument");
    if ((uint64_t)knn->size + count > knn->capacity)
        return fail(knn->ctx, MPTG_ERR_CAPACITY, "mptg_knn_insert: %u + %u exceeds capacity %u", knn->size, count, knn->capacity);
    if (firstOut) *firstOut = knn->size;
    if (count == 0) return MPTG_OK;
    MPTG_CUDA(knn->ctx, cudaSetDevice(knn->ctx->device));
    int rc = knn->scalar == MPTG_F32 ? insertDevT<float>(knn, (const float*)statesDev, count)
                                     : insertDevT<double>(knn, (const double*)statesDev, count);
    if (rc) return rc;
    knn->size += count;
    return MPTG_OK;
}

int mptg_knn_insert(mptg_knn* knn, const void* states, uint32_t count, uint32_t* firstOut) {
    if (!knn || (!states && count)) return fail(knn ? knn->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_knn_insert: bad argument");
    if ((uint64_t)knn->size + count > knn->capacity)
        return fail(knn->ctx, MPTG_ERR_CAPACITY, "mptg_knn_insert: %u + %u exceeds capacity %u", knn->size, count, knn->capacity);
    if (count == 0) {
        if (firstOut) *firstOut = knn->size;
        return MPTG_OK;
    }
    mptg_ctx* ctx = knn->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)count * knn->D * knn->scalar;
    void* stage;
    int rc = scratch(ctx, 0, bytes, &stage);
    if (rc) return rc;
    MPTG_CUDA(ctx, cudaMemcpyAsync(stage, states, bytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = mptg_knn_insert_dev(knn, stage, count, firstOut);
    if (rc) return rc;
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // caller may reuse `states`
    return MPTG_OK;
}

int mptg_knn_get_states(mptg_knn* knn, uint32_t first, uint32_t count, void* out) {
    if (!knn || !out || (uint64_t)first + count > knn->size) return fail(knn ? knn->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_knn_get_states: bad range");
    if (count == 0) return MPTG_OK;
    mptg_ctx* ctx = knn->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)count * knn->D * knn->scalar;
    void* stage;
    int rc = scratch(ctx, 0, bytes, &stage);
    if (rc) return rc;
    const size_t total = (size_t)count * knn->D;
    if (knn->scalar == MPTG_F32)
        knnGatherKernel<float><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const float*)knn->pts, knn->stride, first, count, knn->D, (float*)stage);
    else
        knnGatherKernel<double><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>((const double*)knn->pts, knn->stride, first, count, knn->D, (double*)stage);
    MPTG_LAUNCHED(ctx);
    MPTG_CUDA(ctx, cudaMemcpyAsync(out, stage, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPTG_OK;
}

int mptg_knn_query_dev(mptg_knn* knn, const void* queriesDev, uint32_t Q, uint32_t k, double radius,
                       uint32_t* idxOutDev, void* distOutDev, uint32_t* countOutDev) {
    if (!knn || (!queriesDev && Q) || !idxOutDev || !distOutDev)
        return fail(knn ? knn->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_knn_query: bad argument");
    if (k == 0 || k > MPTG_MAX_K) return fail(knn->ctx, MPTG_ERR_BAD_ARG, "mptg_knn_query: k=%u out of range 1..%d", k, MPTG_MAX_K);
    if (Q == 0) return MPTG_OK;
    MPTG_CUDA(knn->ctx, cudaSetDevice(knn->ctx->device));
    return knn->scalar == MPTG_F32
               ? queryDevT<float>(knn, (const float*)queriesDev, Q, k, radius, idxOutDev, (float*)distOutDev, countOutDev)
               : queryDevT<double>(knn, (const double*)queriesDev, Q, k, radius, idxOutDev, (double*)distOutDev, countOutDev);
}

int mptg_knn_query(mptg_knn* knn, const void* queries, uint32_t Q, uint32_t k, double radius, uint32_t* idxOut,
                   void* distOut, uint32_t* countOut) {
    if (!knn || (!queries && Q) || !idxOut || !distOut) return fail(knn ? knn->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_knn_query: bad argument");
    if (k == 0 || k > MPTG_MAX_K) return fail(knn->ctx, MPTG_ERR_BAD_ARG, "mptg_knn_query: k=%u out of range 1..%d", k, MPTG_MAX_K);
    if (Q == 0) return MPTG_OK;
    mptg_ctx* ctx = knn->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t qBytes = (size_t)Q * knn->D * knn->scalar;
    const size_t iBytes = (size_t)Q * k * sizeof(uint32_t), dBytes = (size_t)Q * k * knn->scalar, cBytes = (size_t)Q * sizeof(uint32_t);
    void* dq;
    void* dout;
    int rc = scratch(ctx, 0, qBytes, &dq);
    if (rc) return rc;
    rc = scratch(ctx, 1, iBytes + dBytes + cBytes, &dout);
    if (rc) return rc;
    // dist first (8-byte aligned for f64), then idx, then counts
    void* dDist = dout;
    uint32_t* dIdx = (uint32_t*)((char*)dout + dBytes);
    uint32_t* dCnt = (uint32_t*)((char*)dout + dBytes + iBytes);
    MPTG_CUDA(ctx, cudaMemcpyAsync(dq, queries, qBytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = mptg_knn_query_dev(knn, dq, Q, k, radius, dIdx, dDist, dCnt);
    if (rc) return rc;
    MPTG_CUDA(ctx, cudaMemcpyAsync(distOut, dDist, dBytes, cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaMemcpyAsync(idxOut, dIdx, iBytes, cudaMemcpyDeviceToHost, ctx->stream));
    if (countOut) MPTG_CUDA(ctx, cudaMemcpyAsync(countOut, dCnt, cBytes, cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPTG_OK;
}

int mptg_knn_build_index(mptg_knn* knn) {
    if (!knn) return fail(nullptr, MPTG_ERR_BAD_ARG, "mptg_knn_build_index: null handle");
    MPTG_CUDA(knn->ctx, cudaSetDevice(knn->ctx->device));
    knn->index.count = 0;  // force a rebuild over everything stored so far
    return knn->scalar == MPTG_F32
               ? knnBuildIndex<float>(knn->ctx, knn->index, knn->space, (const float*)knn->pts, knn->stride, knn->size)
               : knnBuildIndex<double>(knn->ctx, knn->index, knn->space, (const double*)knn->pts, knn->stride, knn->size);
}

int mptg_knn_last_stats(mptg_knn* knn, uint64_t out[4]) {
    if (!knn || !out) return fail(knn ? knn->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_knn_last_stats: bad argument");
    // device-side counters of the tree search are folded in on demand
    int rc = knnIndexReadStats(knn->ctx, knn->index, knn->stats);
    if (rc) return rc;
    for (int i = 0; i < 4; ++i) out[i] = knn->stats[i];
    return MPTG_OK;
}

int mptg_knn_insert_ids(mptg_knn* knn, const void* states, const uint32_t* ids, uint32_t count) {
    if (!knn || ((!states || !ids) && count)) return fail(knn ? knn->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_knn_insert_ids: bad argument");
    if (knn->size != 0 && !knn->gid) return fail(knn->ctx, MPTG_ERR_BAD_ARG, "mptg_knn_insert_ids: the set already holds points without ids");
    mptg_ctx* ctx = knn->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!knn->gid) MPTG_CUDA(ctx, cudaMalloc(&knn->gid, (size_t)knn->capacity * sizeof(uint32_t)));
    uint32_t first = 0;
    if (int rc = mptg_knn_insert(knn, states, count, &first)) return rc;
    if (count) MPTG_CUDA(ctx, cudaMemcpyAsync(knn->gid + first, ids, (size_t)count * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPTG_OK;
}

int mptg_knn_merge_dev(mptg_ctx* ctx, int scalar, uint32_t parts, uint32_t Q, uint32_t k, const uint32_t* idxIn,
                       const void* distIn, uint32_t* idxOut, void* distOut, uint32_t* countOut) {
    if (!ctx || !idxIn || !distIn || !idxOut || !distOut || parts == 0 || k == 0 || k > MPTG_MAX_K)
        return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_knn_merge_dev: bad argument");
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (scalar == MPTG_F32) return launchMerge<float>(ctx, parts, Q, k, idxIn, (const float*)distIn, idxOut, (float*)distOut, countOut);
    if (scalar == MPTG_F64) return launchMerge<double>(ctx, parts, Q, k, idxIn, (const double*)distIn, idxOut, (double*)distOut, countOut);
    return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_knn_merge_dev: bad scalar");
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// internal entry points of the sharded search (comm.cu)
// ---------------------------------------------------------------------------------------------
namespace mptg {

mptg_ctx* knnShardCtx(mptg_knn* knn) { return knn->ctx; }
int knnShardScalar(const mptg_knn* knn) { return knn->scalar; }
int knnShardScalars(const mptg_knn* knn) { return knn->D; }
const mptg_space_desc* knnShardSpace(const mptg_knn* knn) { return &knn->space; }
uint32_t knnShardSize(const mptg_knn* knn) { return knn->size; }
uint64_t knnShardBuilds(const mptg_knn* knn) { return knn->index.builds; }

namespace {
// lb[g][q] for every shard g from the synchronised shard boxes (generic box bound of knn_bvh.cuh: conservative for every
// space), rounded down to float; home(q) = the shard with the smallest bound (lowest rank on ties); cap = +inf for this
// rank's home queries, -1 for the others (the home search skips them).  One thread per query, one box per shard.
template <typename S>
__global__ void __launch_bounds__(256) shardRootKernel(DevSpace<S> sp, const S* __restrict__ shardBox, const uint32_t* __restrict__ peerN, int world, int rank,
                                                       const S* __restrict__ queries, uint32_t Q, float* __restrict__ lbMine, uint8_t* __restrict__ home,
                                                       S* __restrict__ cap) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const int D = sp.D;
    const S* myq = queries + (size_t)q * D;
    int best = 0;
    float bl = INFINITY, mine = INFINITY;
    for (int g = 0; g < world; ++g) {
        float v = INFINITY;
        if (peerN[g] != 0u) {
            const S* b = shardBox + (size_t)g * (size_t)(2 * D);
            const S lb = dev::boxLowerBound<S>(
                sp, [&](int c) { return b[c]; }, [&](int c) { return b[D + c]; }, [&](int c) { return myq[c]; });
            v = __uint_as_float(boundKey<S>(lb));
        }
        if (v < bl) bl = v, best = g;
        if (g == rank) mine = v;
    }
    lbMine[q] = mine;
    home[q] = (uint8_t)best;
    cap[q] = best == rank ? (S)INFINITY : S(-1);
}
// one box per shard: the union of the boxes of the children of its top node ([2 D][32] rows -> [2 D])
template <typename S>
__global__ void shardUnionKernel(const S* __restrict__ peerBox, const uint32_t* __restrict__ peerN, int world, int D, S* __restrict__ shardBox) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= world * D) return;
    const int g = t / D, c = t % D;
    const S* b = peerBox + (size_t)g * (size_t)(2 * D) * 32u;
    S lo = (S)INFINITY, hi = (S)-INFINITY;
    for (uint32_t j = 0; j < peerN[g] && j < 32u; ++j) {
        lo = fmin(lo, b[(size_t)c * 32u + j]);
        hi = fmax(hi, b[(size_t)(D + c) * 32u + j]);
    }
    shardBox[(size_t)g * 2 * D + c] = lo;
    shardBox[(size_t)g * 2 * D + D + c] = hi;
}
template <typename S>
int shardQueryT(mptg_knn* knn, const S* q, uint32_t Q, uint32_t k, double radius, const S* qcap, uint32_t* idxOut, S* distOut) {
    return knnBvhQuery<S>(knn->ctx, knn->index, knn->space, q, Q, k, radius, knn->idxMul, knn->idxAdd, idxOut, distOut, nullptr, knn->stats, knn->gid, qcap,
                          nullptr);
}
}  // namespace

// The sharded search goes through the tree only (per-query radius caps live in the tree kernels): index everything stored,
// and hand out the boxes of the top node's children (device, [2 D][32] scalars, lo rows then hi rows) and their number.
int knnShardIndexAll(mptg_knn* knn, const void** topBoxDev, uint32_t* nTop) {
    *topBoxDev = nullptr;
    *nTop = 0;
    if (knn->size == 0) return MPTG_OK;
    if (knn->index.count != knn->size) {
        knn->index.count = 0;
        const int rc = knn->scalar == MPTG_F32 ? knnBuildIndex<float>(knn->ctx, knn->index, knn->space, (const float*)knn->pts, knn->stride, knn->size)
                                               : knnBuildIndex<double>(knn->ctx, knn->index, knn->space, (const double*)knn->pts, knn->stride, knn->size);
        if (rc) return rc;
    }
    *topBoxDev = knn->index.box[knn->index.top];
    *nTop = knn->index.nNodes[knn->index.top];
    return MPTG_OK;
}
// one box per shard from the synchronised top boxes (after mptg_knn_shard_sync's all-gather)
int knnShardUnion(mptg_knn* knn, const void* peerBox, const uint32_t* peerN, int world, void* shardBox) {
    mptg_ctx* ctx = knn->ctx;
    const int n = world * knn->D;
    if (knn->scalar == MPTG_F32) shardUnionKernel<float><<<(n + 127) / 128, 128, 0, ctx->stream>>>((const float*)peerBox, peerN, world, knn->D, (float*)shardBox);
    else shardUnionKernel<double><<<(n + 127) / 128, 128, 0, ctx->stream>>>((const double*)peerBox, peerN, world, knn->D, (double*)shardBox);
    MPTG_LAUNCHED(ctx);
    return MPTG_OK;
}
// bounds of every query to every shard, home shard, and the cap array of the home search (see shardRootKernel)
int knnShardRootAll(mptg_knn* knn, const void* shardBox, const uint32_t* peerN, int world, int rank, const void* queriesDev, uint32_t Q, float* lbMine,
                    uint8_t* home, void* cap) {
    mptg_ctx* ctx = knn->ctx;
    const dim3 grid((Q + 255) / 256), block(256);
    if (knn->scalar == MPTG_F32)
        shardRootKernel<float><<<grid, block, 0, ctx->stream>>>(makeDevSpace<float>(knn->space), (const float*)shardBox, peerN, world, rank, (const float*)queriesDev,
                                                                Q, lbMine, home, (float*)cap);
    else
        shardRootKernel<double><<<grid, block, 0, ctx->stream>>>(makeDevSpace<double>(knn->space), (const double*)shardBox, peerN, world, rank,
                                                                 (const double*)queriesDev, Q, lbMine, home, (double*)cap);
    MPTG_LAUNCHED(ctx);
    return MPTG_OK;
}
// search with a per-query radius cap (scalar type of the space; < 0: skip the query, its output row is left alone)
int knnShardQuery(mptg_knn* knn, const void* queriesDev, uint32_t Q, uint32_t k, double radius, const void* qcapDev, uint32_t* idxOut,
                  void* distOut) {
    if (knn->size == 0) return MPTG_OK;
    return knn->scalar == MPTG_F32 ? shardQueryT<float>(knn, (const float*)queriesDev, Q, k, radius, (const float*)qcapDev, idxOut, (float*)distOut)
                                   : shardQueryT<double>(knn, (const double*)queriesDev, Q, k, radius, (const double*)qcapDev, idxOut, (double*)distOut);
}

}  // namespace mptg
