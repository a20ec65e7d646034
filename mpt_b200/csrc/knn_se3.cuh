// knn_se3.cuh -- the SE(3)/float32 search over the cap image of the index (KnnIndex::cap, KnnIndex::leafH), the kernel
// the C5 wave of BASELINE.json spends its time in.
//
// Node bound.  A node stores, for its rotations, a geodesic cap on RP^3 (unit quaternions up to sign): a centre c and an
// angular radius rho with angle(p, c) <= rho for every member p, and for its translations an axis-aligned box.  By the
// triangle inequality of the angle metric, with theta = angle(q, c):
//      |q.p| <= |q||p| cos(max(0, theta - rho)),
// and the weighted sum of acos of that and of the distance to the box is a lower bound of the SE(3) distance to every
// member.  The r1 index bounded the dot product over an axis-aligned box of the quaternion COEFFICIENTS instead; with
// 32,768 leaves over 1 M uniform rotations such a box is half a unit wide per coefficient and its corner maximum of the
// dot product is ~1, i.e. no rotation pruning at all (10,291 points examined per query against 6,421 inside the
// translation ball alone).  Caps halve the leaves a query has to visit (tools/knn_tree_shape_experiment.py: floor
// 270 -> 125 per query; measured 321 -> 156).
//
// Leaf prefilter.  Two leaves per pass, one point of each per lane, in packed half precision (HFMA2 / HMUL2 / HADD2 on
// the lane's two points at once, no conversions, no square roots):  with A = (w0 chord / T)^2 and B = (w1 r / T)^2,
//      w0 chord + w1 r <= T   <=>   C = 1 - A - B >= 0  and  4AB <= C^2,
// chord = sqrt(2 - 2|dot|) <= acos|dot|.  Every rounding is pushed to the accepting side (see Se3Walk::refreshThr); a
// lane that passes is evaluated exactly, in the operation order of mptg_space.h, against the float copy of the leaf --
// so the result is the exhaustive scan's, bit for bit.
#pragma once

#include <cuda_fp16.h>

namespace mptg {

// Per-query block in shared memory (16 floats, 16-byte aligned), written once per query by se3QueryPrep and read with
// 128-bit loads where it is needed -- the traversal keeps only the packed half-precision copies in registers:
//   [0..3]  query quaternion as given      [4..6] query translation, [7] unused
//   [8..11] query quaternion times normMax (|p| <= normMax for every indexed p)
//   [12]    squared norm of that, rounded up   [13] its norm, rounded up   [14] w0   [15] w1
constexpr int SE3_QF = 16;

__device__ __forceinline__ void se3QueryPrep(float* sq, const BvhArgs<float>& a, int lane) {
    __syncwarp();
    if (lane == 0) {
        const float n = a.normMax;
        const float q0 = sq[0] * n, q1 = sq[1] * n, q2 = sq[2] * n, q3 = sq[3] * n;
        const float nq2 = __fmaf_rn(q3, q3, __fmaf_rn(q2, q2, __fmaf_rn(q1, q1, q0 * q0))) * (1.0f + 2e-6f);
        sq[7] = 0.0f;
        sq[8] = q0, sq[9] = q1, sq[10] = q2, sq[11] = q3;
        sq[12] = nq2;
        sq[13] = sqrtf(nq2) * (1.0f + 1e-6f);
        sq[14] = a.sp.weighted[0] ? a.sp.weight[0] : 1.0f;
        sq[15] = a.sp.weighted[1] ? a.sp.weight[1] : 1.0f;
    }
    __syncwarp();
}

// lower-bound key of child `lane` of `block` at one level of the cap image.
// Rounding: x carries <= 2.4e-7 |qs|, and the derivative of x cos rho + sqrt(n^2 - x^2) sin rho in x is <= 2 cos rho on
// theta >= rho; nq2 / nqn are rounded up; the approximate square root adds 2.4e-7 relative; the centre's norm is 1 to
// 1.2e-7; the distance's own fma-chain dot carries <= 2.4e-7 |q||p| -- all inside the slack 3e-6 |qs|.  Then
// acos(ad) >= sqrt(2 - 2 ad), with the slacks of se3CheapBound for the square roots and the final sum.
__device__ __forceinline__ uint32_t se3CapKey(const float4* __restrict__ cap, uint32_t block, int lane, const float* sq) {
    const float4* cp = cap + ((size_t)block * 3u) * 32u + lane;
    const float4 c = __ldg(cp), r = __ldg(cp + 32), t = __ldg(cp + 64);
    const float4 qs = *reinterpret_cast<const float4*>(sq + 8), nw = *reinterpret_cast<const float4*>(sq + 12);
    const float4 qt = *reinterpret_cast<const float4*>(sq + 4);
    float x = c.x * qs.x;
    x = __fmaf_rn(c.y, qs.y, x);
    x = __fmaf_rn(c.z, qs.z, x);
    x = fabsf(__fmaf_rn(c.w, qs.w, x));
    const float sn = sqrtApprox(fmaxf(__fmaf_rn(-x, x, nw.x), 0.0f));
    float y = __fmaf_rn(sn, r.y, x * r.x);
    y = x >= nw.y * r.x ? nw.y : y;  // inside the cap: no rotation bound (NaN query: comparison false, y NaN, ad = 1)
    const float ad = fminf(1.0f, __fmaf_rn(nw.y, 3e-6f, y));
    const float e0 = fmaxf(fmaxf(r.z - qt.x, qt.x - r.w), 0.0f);
    const float e1 = fmaxf(fmaxf(t.x - qt.y, qt.y - t.y), 0.0f);
    const float e2 = fmaxf(fmaxf(t.z - qt.z, qt.z - t.w), 0.0f);
    const float acc = __fmaf_rn(e2, e2, __fmaf_rn(e1, e1, e0 * e0));
    return __float_as_uint(se3CheapBound(ad, acc, nw.z, nw.w) + 0.0f);
}

template <int KPL>
struct Se3Walk {
    const BvhArgs<float>& a;
    const float* sq;  // the query block in shared memory
    int lane;
    __half2 hq01, hq23, hnt01, hnt2;  // query quaternion; minus the query translation in the leaf copies' scale
    __half2 hG, hKa, hKb;             // prefilter scales for the current threshold
    float thrF, rad;
    WarpTopK<float, KPL> top;
    uint32_t& leaves;
    uint32_t& inner;
#ifdef MPTG_KNN_PROBE
    uint32_t useful = 0, cand = 0;
#endif

    __device__ __forceinline__ Se3Walk(const BvhArgs<float>& args, const float* q, int ln, uint32_t& lv, uint32_t& in)
        : a(args), sq(q), lane(ln), leaves(lv), inner(in) {}

    // Prefilter constants.  u = 2^-11 (half rounding).  Everything errs towards "pass":
    //  * dot: the stored copy is off by <= qabs errQ, the half query by <= u |q||p|, the three half operations by
    //    <= 3u |q||p|, the two scale constants below by the equivalent of 2u: dotC = 2 - 2 (qabs errQ + 12 u |qs|)
    //    (twice the sum); the chord term is then scaled by ka4 = 3.95 (w0 / Teff)^2 (3.95 for 4: roundings of the
    //    products), rounded down, clamped to the half range (a smaller scale only passes more)
    //  * translation: |p_h - q_h| >= |p - q| - sqrt(3) (errT + u max|q_t|), folded into the threshold:
    //    Teff = T (1 + 4e-3) + w1 sqrt(3)(errT + u max|q_t|); the differences are scaled by g = w1 tScale^-1 / Teff
    //    (rounded down) before they are squared, so nothing overflows for thresholds the half range can express
    //  * C = (1 + 8e-3) - A - B absorbs the ~10 u of the remaining sums and products
    //  * thresholds that do not fit (infinite: fewer than k found so far; g below 2.5e-4) zero all three scales:
    //    A = B = 0, every lane of a visited leaf passes and is evaluated exactly until the threshold has come down.
    //    (0 x finite = 0: the stored copies are finite and the half query is clamped to the half range.)
    __device__ __forceinline__ void refreshThr() {
        thrF = top.kthD < rad ? top.kthD : rad;
        // (recomputed here rather than kept in registers: this runs only after an insertion)
        const float u = 4.8828125e-4f;
        const float qabs = fabsf(sq[0]) + fabsf(sq[1]) + fabsf(sq[2]) + fabsf(sq[3]);
        const float dotC = 2.0f - 2.0f * (qabs * a.errQ + 12.0f * u * sq[13]);
        const float tmax = fmaxf(fabsf(sq[4]), fmaxf(fabsf(sq[5]), fabsf(sq[6])));
        const float slackT = sq[15] * 1.7321f * (a.errT * a.tScaleInv + u * tmax) + 1e-30f;
        const float Teff = __fmaf_rn(thrF, 1.004f, slackT);
        float g = __fdividef(sq[15] * a.tScaleInv, Teff);  // e = (p_scaled - q_scaled) g = (p - q) w1 / Teff
        const bool fits = g >= 2.5e-4f;                    // false also for NaN
        g = fits ? fminf(g, 128.0f) : 0.0f;
        const float sa = __fdividef(sq[14], Teff);
        const float ka4 = fits ? fminf(sa * sa * 3.95f, 15000.0f) : 0.0f;
        const __half hg = __float2half_rd(g);
        hG = __halves2half2(hg, hg);
        const __half ha = __float2half_rd(-2.0f * ka4), hb = __float2half_rd(ka4 * dotC);
        hKa = __halves2half2(ha, ha);
        hKb = __halves2half2(hb, hb);
    }
    __device__ __forceinline__ void start(float radius) {
        top.init(a.k);
        rad = radius;
        hq01 = __floats2half2_rn(sq[0], sq[1]);
        hq23 = __floats2half2_rn(sq[2], sq[3]);
        const float sc = -a.tScale, lim = 60000.0f;  // clamped: an out-of-range query only under-estimates its distances
        hnt01 = __floats2half2_rn(fminf(fmaxf(sq[4] * sc, -lim), lim), fminf(fmaxf(sq[5] * sc, -lim), lim));
        hnt2 = __floats2half2_rn(fminf(fmaxf(sq[6] * sc, -lim), lim), 0.0f);
        refreshThr();
    }

    // exact distance of this lane's point (operation order of mptg_space.h) and offer
    __device__ __forceinline__ void leafExact(uint32_t node, bool maybe) {
        const uint32_t p = node * 32u + (uint32_t)lane;
        const uint32_t orig = __ldg(a.perm + p);
        const float* pt = a.leafPts + ((size_t)node * 7u) * 32u + lane;
        float pv[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) pv[c] = __ldg(pt + c * 32);
        const float4 qq = *reinterpret_cast<const float4*>(sq), qt = *reinterpret_cast<const float4*>(sq + 4);
        float dot = pv[0] * qq.x;
        dot = __fmaf_rn(pv[1], qq.y, dot);
        dot = __fmaf_rn(pv[2], qq.z, dot);
        dot = __fmaf_rn(pv[3], qq.w, dot);
        const float d0 = pv[4] - qt.x, d1 = pv[5] - qt.y, d2 = pv[6] - qt.z;
        float s2 = d0 * d0;
        s2 = __fmaf_rn(d1, d1, s2);
        s2 = __fmaf_rn(d2, d2, s2);
        const float ad = fminf(1.0f, fabsf(dot));
        float dr = fp::acos01(ad);
        if (a.sp.weighted[0]) dr = dr * a.sp.weight[0];
        float dt = fp::sqrt_(s2);
        if (a.sp.weighted[1]) dt = dt * a.sp.weight[1];
        const uint32_t rep = a.gid ? (orig != MPTG_NO_INDEX ? __ldg(a.gid + orig) : MPTG_NO_INDEX) : orig * a.idxMul + a.idxAdd;
        top.offer(maybe && orig != MPTG_NO_INDEX, dr + dt, rep, rad, lane);
        refreshThr();
    }

    // one pass over two leaves (A in the low halves, B in the high halves); `two` false: B repeats A and is ignored
    __device__ __forceinline__ void evalPair(uint32_t nodeA, uint32_t nodeB, bool two, const uint4 A, const uint4 B) {
        const __half2 zero = __float2half2_rn(0.0f);
        const __half2 a0 = *reinterpret_cast<const __half2*>(&A.x), a1 = *reinterpret_cast<const __half2*>(&A.y);
        const __half2 a2 = *reinterpret_cast<const __half2*>(&A.z), a3 = *reinterpret_cast<const __half2*>(&A.w);
        const __half2 b0 = *reinterpret_cast<const __half2*>(&B.x), b1 = *reinterpret_cast<const __half2*>(&B.y);
        const __half2 b2 = *reinterpret_cast<const __half2*>(&B.z), b3 = *reinterpret_cast<const __half2*>(&B.w);
        const __half2 m = __hfma2(a1, hq23, __hmul2(a0, hq01));  // (x qx + z qz, y qy + w qw) of A's point
        const __half2 n = __hfma2(b1, hq23, __hmul2(b0, hq01));
        const __half2 D = __hadd2(__lows2half2(m, n), __highs2half2(m, n));  // (dot A, dot B)
        const __half2 ea0 = __hmul2(__hadd2(a2, hnt01), hG), ea1 = __hmul2(__hadd2(a3, hnt2), hG);
        const __half2 eb0 = __hmul2(__hadd2(b2, hnt01), hG), eb1 = __hmul2(__hadd2(b3, hnt2), hG);
        const __half2 sa = __hfma2(ea1, ea1, __hmul2(ea0, ea0));  // (ex^2 + ez^2, ey^2)
        const __half2 sb = __hfma2(eb1, eb1, __hmul2(eb0, eb0));
        const __half2 Bq = __hadd2(__lows2half2(sa, sb), __highs2half2(sa, sb));  // (w1 r / Teff)^2 of A, B
        const __half2 a4 = __hmax2(__hfma2(__habs2(D), hKa, hKb), zero);           // ~4 (w0 chord / Teff)^2
        const __half2 C = __hsub2(__hfma2(a4, __float2half2_rn(-0.25f), __float2half2_rn(1.008f)), Bq);
        const unsigned pass = __hge2_mask(C, zero) & __hle2_mask(__hmul2(a4, Bq), __hmul2(C, C));
#ifdef MPTG_KNN_PROBE
        {
            const unsigned pa = __ballot_sync(FULL_MASK, (pass & 0xffffu) != 0u), pb = two ? __ballot_sync(FULL_MASK, (pass >> 16) != 0u) : 0u;
            useful += (pa != 0u) + (pb != 0u);
            cand += __popc(pa) + __popc(pb);
        }
#endif
        if (!__any_sync(FULL_MASK, pass != 0u)) return;
        // rare: fetch the exact points and evaluate the true distance
        if (__any_sync(FULL_MASK, (pass & 0xffffu) != 0u)) leafExact(nodeA, (pass & 0xffffu) != 0u);
        if (two && __any_sync(FULL_MASK, (pass >> 16) != 0u)) leafExact(nodeB, (pass >> 16) != 0u);
    }

    // the leaves of one level-0 block: the nearest one and the first other candidate as the first pair (the nearest
    // tightens the threshold early), the rest two at a time in lane order against the threshold as it shrinks
    __device__ __forceinline__ void leafBlock(uint32_t block) {
        const uint32_t key = block * 32u + (uint32_t)lane < a.nNodes[0] ? se3CapKey(a.cap[0], block, lane, sq) : BVH_DEAD;
        const uint32_t best = __reduce_min_sync(FULL_MASK, key);
        if (best == BVH_DEAD || __uint_as_float(best) > thrF) return;
        unsigned m = __ballot_sync(FULL_MASK, __uint_as_float(key) <= thrF);  // BVH_DEAD is a NaN pattern
        const char* base = reinterpret_cast<const char*>(a.leafH) + ((size_t)block << 14) + ((uint32_t)lane << 4);
        const uint32_t node0 = block * 32u;
        uint32_t s0 = (uint32_t)__ffs(__ballot_sync(FULL_MASK, key == best)) - 1u;
        m &= ~(1u << s0);
        for (;;) {
            const bool two = m != 0u;
            const uint32_t s1 = two ? (uint32_t)__ffs(m) - 1u : s0;
            m &= m - 1u;  // m == 0 stays 0
            const uint4 A = __ldg(reinterpret_cast<const uint4*>(base + (s0 << 9)));
            const uint4 B = __ldg(reinterpret_cast<const uint4*>(base + (s1 << 9)));
            leaves += two ? 2u : 1u;
            const float before = thrF;
            evalPair(node0 + s0, node0 + s1, two, A, B);
            if (m == 0u) return;
            if (thrF != before) {
                m &= __ballot_sync(FULL_MASK, __uint_as_float(key) <= thrF);
                if (m == 0u) return;
            }
            s0 = (uint32_t)__ffs(m) - 1u;
            m &= m - 1u;
        }
    }

    // inner levels: children nearest bound first
    template <int L>
    __device__ __forceinline__ void descend(uint32_t block) {
        if constexpr (L == 0) {
            leafBlock(block);
        } else {
            uint32_t key = block * 32u + (uint32_t)lane < a.nNodes[L] ? se3CapKey(a.cap[L], block, lane, sq) : BVH_DEAD;
            for (;;) {
                const uint32_t best = __reduce_min_sync(FULL_MASK, key);
                if (best == BVH_DEAD || __uint_as_float(best) > thrF) return;
                const int src = __ffs(__ballot_sync(FULL_MASK, key == best)) - 1;
                if (lane == src) key = BVH_DEAD;
                ++inner;
                descend<L - 1>(block * 32u + (uint32_t)src);
            }
        }
    }
};

#ifndef MPTG_SE3_MIN_CTAS
#define MPTG_SE3_MIN_CTAS 4
#endif
// Persistent warps: the grid is sized to the machine and every warp draws the next position of the (spatially sorted)
// wave from a counter until the wave is exhausted.  A search takes 0.5x to 3x the mean, and with one query per warp of
// an 8-warp CTA a slot stayed occupied until its slowest warp was done: ncu showed 23 of the 32 resident warps active
// and the kernel waiting on loads (long scoreboard 3.3 per issue) rather than issuing.
template <int KPL>
__global__ void __launch_bounds__(BVH_WARPS * 32, KPL == 1 ? MPTG_SE3_MIN_CTAS : (KPL == 2 ? 3 : 2)) knnSe3Kernel(const BvhArgs<float> a) {
    __shared__ __align__(16) float qsm[BVH_WARPS][SE3_QF];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* sq = qsm[warp];
    uint32_t leaves = 0, inner = 0;
#ifdef MPTG_KNN_PROBE
    uint32_t useful = 0, cand = 0;
#endif
    const uint32_t nSlots = a.nActive ? __ldg(a.nActive) : a.Q;
    for (;;) {
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(a.cursor, 1u);
        slot = __shfl_sync(FULL_MASK, slot, 0);
        if (slot >= nSlots) break;  // warp-uniform; no block-wide barriers in this kernel
        const uint32_t q = a.order ? __ldg(a.order + slot) : slot;
        float radius = a.radius;
        if (a.qcap) {  // sharded search: per-query cap, negative = not this shard's query
            const float c = __ldg(a.qcap + q);
            if (c < 0.0f) continue;
            radius = fminf(radius, c);
        }
        __syncwarp();
        if (lane < 7) sq[lane] = a.queries[(size_t)q * 7u + lane];
        se3QueryPrep(sq, a, lane);
        Se3Walk<KPL> w(a, sq, lane, leaves, inner);
        w.start(radius);
        switch (a.top) {
            case 0: w.template descend<0>(0); break;
            case 1: w.template descend<1>(0); break;
            case 2: w.template descend<2>(0); break;
            case 3: w.template descend<3>(0); break;
            default: w.template descend<4>(0); break;
        }
        const uint32_t count = w.top.store(a.k, a.idxOut + (size_t)q * a.k, a.distOut + (size_t)q * a.k, lane);
        if (a.countOut && lane == 0) a.countOut[q] = count;
#ifdef MPTG_KNN_PROBE
        useful += w.useful;
        cand += w.cand;
#endif
    }
    if (lane == 0 && a.stats) {
        atomicAdd(a.stats + 0, (unsigned long long)leaves);
        atomicAdd(a.stats + 1, (unsigned long long)inner);
#ifdef MPTG_KNN_PROBE
        atomicAdd(a.stats + 2, (unsigned long long)useful);
        atomicAdd(a.stats + 3, (unsigned long long)cand);
#endif
    }
}

// query ordering key for the cap image: the leaf reached by always following the smallest child bound
__global__ void __launch_bounds__(BVH_WARPS * 32) knnSe3KeyKernel(const BvhArgs<float> a) {
    __shared__ __align__(16) float qsm[BVH_WARPS][SE3_QF];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * BVH_WARPS + warp;
    if (q >= a.Q) return;
    if (a.qcap && a.qcap[q] < 0.0f) return;  // not searched in this pass: not in the order
    float* sq = qsm[warp];
    if (lane < 7) sq[lane] = a.queries[(size_t)q * 7u + lane];
    se3QueryPrep(sq, a, lane);
    uint32_t node = 0;
    for (int l = a.top; l >= 0; --l) {
        const uint32_t key = node * 32u + (uint32_t)lane < a.nNodes[l] ? se3CapKey(a.cap[l], node, lane, sq) : BVH_DEAD;
        const uint32_t best = __reduce_min_sync(FULL_MASK, key);
        const int src = __ffs(__ballot_sync(FULL_MASK, key == best)) - 1;
        node = node * 32u + (uint32_t)(src < 0 ? 0 : src);
    }
    if (lane == 0) {
        a.orderKeys[q] = node;
        atomicAdd(a.orderHist + (node >> a.orderShift), 1u);
    }
}


// lower bound of every query to the whole indexed set (see knnBvhRootBoundKernel)
__global__ void __launch_bounds__(BVH_WARPS * 32) knnSe3RootBoundKernel(const BvhArgs<float> a) {
    __shared__ __align__(16) float qsm[BVH_WARPS][SE3_QF];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * BVH_WARPS + warp;
    if (q >= a.Q) return;
    float* sq = qsm[warp];
    if (lane < 7) sq[lane] = a.queries[(size_t)q * 7u + lane];
    se3QueryPrep(sq, a, lane);
    const uint32_t key = (uint32_t)lane < a.nNodes[a.top] ? se3CapKey(a.cap[a.top], 0, lane, sq) : BVH_DEAD;
    const uint32_t best = __reduce_min_sync(FULL_MASK, key);
    if (lane == 0) a.rootLb[q] = best == BVH_DEAD ? INFINITY : __uint_as_float(best);
}

}  // namespace mptg
