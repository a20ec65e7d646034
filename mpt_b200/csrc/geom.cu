// geom.cu -- batched state / edge validity for the occupancy grid, balls & rectangles and the planar
// N-link arm, plus the geometry part of the C ABI.
//
// Edge checks of these scenarios are recursive midpoint bisections in the reference
// (demo/png_2d_scenario.hpp:152-165, demo/shape_hierarchy.hpp:191-203,
// demo/link_manipulator_scenario.hpp:125-138).  The recursion tree depends only on the floating
// point endpoints (mid = (a+b)/2, stop test on the CURRENT pair), and the answer is the AND over
// every midpoint in that tree, so visiting order is free.  One warp takes one edge: the 31 midpoints
// of the five top levels are probed one per lane in a single parallel step, then lane L runs subtree
// L (rooted at depth five) depth first with an explicit stack; a warp vote per iteration stops all
// lanes at the first invalid probe.  All midpoints are produced by the same sequence of
// floating point operations as the reference's recursion, so decisions are bit-identical.
#include <cub/device/device_scan.cuh>
#include <cstdio>
#include <cstdlib>

#include "geom.cuh"
#include "nao.cuh"

namespace mptg {

// ------------------------------------------------------------------ validators
constexpr int FLAT_MAX_LEVELS = 24;

// number of levels k >= 0 with d / 2^k >= thresh - slack: no node of the recursion is deeper (see Validator::levels)
template <typename S>
__device__ __forceinline__ int levelsFor(S d, S t) {
    if (!(t > S(0))) return FLAT_MAX_LEVELS + 1;
    int L = 0;
    if (!(d == d)) return 1;  // NaN never meets the stop test: the root is probed (and fails)
    while (d >= t && L <= FLAT_MAX_LEVELS) {
        d = d * S(0.5);
        ++L;
    }
    return L;
}

template <typename S>
struct GridValidator {
    static constexpr int MAXD = 2;
    static constexpr bool CHECK_ENDS = true;
    static constexpr int MAXDEPTH = 48;
    const uint32_t* bits;
    int width, height;
    __device__ __forceinline__ int dims() const { return 2; }
    // demo/png_2d_scenario.hpp:104-110.  Out-of-range linear index (the reference reads out of
    // bounds there) -> obstacle.
    __device__ __forceinline__ bool valid(const S* q) const {
        const int x = (int)(q[0] + S(0.5));
        const int y = (int)(q[1] + S(0.5));
        const long long idx = (long long)width * y + x;
        if (idx < 0 || idx >= (long long)width * height) return false;
        return ((__ldg(bits + (idx >> 5)) >> (idx & 31)) & 1u) == 0;
    }
    __device__ __forceinline__ bool endpointsValid(const S* a, const S* b) const { return valid(a) && valid(b); }
    // :155-158  (b - a).squaredNorm() < 1
    __device__ __forceinline__ bool stop(const S* a, const S* b) const {
        const S dx = b[0] - a[0], dy = b[1] - a[1];
        return dx * dx + dy * dy < S(1);
    }
};

template <typename S>
struct ShapesValidator {
    static constexpr int MAXD = 2;
    static constexpr bool CHECK_ENDS = true;
    static constexpr int MAXDEPTH = 48;
    const S* rects;
    int nRects;
    __device__ __forceinline__ int dims() const { return 2; }
    // shape_hierarchy.hpp:177-182, AND over rectangles
    __device__ __forceinline__ bool valid(const S* p) const {
        for (int j = 0; j < nRects; ++j) {
            const S* r = rects + 4 * j;
            if (p[0] >= r[0] && p[0] <= r[2] && p[1] >= r[1] && p[1] <= r[3]) return false;
        }
        return true;
    }
    __device__ __forceinline__ bool endpointsValid(const S* a, const S* b) const { return valid(a) && valid(b); }
    __device__ __forceinline__ bool stop(const S* a, const S* b) const {  // :194-197
        const S dx = b[0] - a[0], dy = b[1] - a[1];
        return dx * dx + dy * dy < S(1);
    }
};

// shape_hierarchy.hpp:259-270 for 2-D points
template <typename S>
__device__ __forceinline__ S distPointSegmentSquared2(const S* pt, const S* s0, const S* s1) {
    const S vx = s1[0] - s0[0], vy = s1[1] - s0[1];
    const S wx = pt[0] - s0[0], wy = pt[1] - s0[1];
    const S c1 = vx * wx + vy * wy;
    if (c1 <= S(0)) return wx * wx + wy * wy;
    const S c2 = vx * vx + vy * vy;
    if (c2 <= c1) {
        const S ex = pt[0] - s1[0], ey = pt[1] - s1[1];
        return ex * ex + ey * ey;
    }
    const S f = fp::div_(c1, c2);
    const S ex = s0[0] - pt[0] + vx * f, ey = s0[1] - pt[1] + vy * f;
    return ex * ex + ey * ey;
}

// EXACT: the arm has exactly MAXD_ links, so every loop over the links has a compile-time trip count and the state arrays
// of the edge kernels sit in registers (with a run-time count they are indexed dynamically and live in local memory)
template <typename S, int MAXD_, bool EXACT = false>
struct ArmValidator {
    static constexpr int MAXD = MAXD_;
    static constexpr bool CHECK_ENDS = true;
    static constexpr int MAXDEPTH = 8;  // beyond the 5 split levels: |a-b|_inf up to 0.02 * 2^13 rad
    const S* lengths;
    const S* circles;
    int nLinks, nCircles;
    S linkRadius;
    __device__ __forceinline__ int dims() const { return EXACT ? MAXD_ : nLinks; }
    // demo/link_manipulator_scenario.hpp:99-116
    __device__ bool valid(const S* q) const {
        S from[2] = {S(0), S(0)}, to[2];
        S angle = S(0);
#pragma unroll
        for (int i = 0; i < (EXACT ? MAXD_ : nLinks); ++i) {
            angle = angle + q[i];
            S sn, cs;
            fp::sincos_(angle, &sn, &cs);
            const S len = lengths[i];
            to[0] = from[0] + len * cs;
            to[1] = from[1] + len * sn;
            // Far circles are passed without the exact test.  With w = centre - from, v = to - from as computed:
            // the distance from w to the segment [0, v] is >= |w| - |v|, and the value the exact test computes
            // is within a few eps * |w| of it (case c1 <= 0 returns this very |w|^2; the far-end case is |w - v|
            // up to one rounding per coordinate; the projection adds v * f with |v f| <= |v| < |w| to -w, which is
            // exact).  So |w|^2 > ((|v| + rr) * (1 + 1024 eps))^2, all in relative terms, implies the exact
            // test's "distance^2 > rr^2": same decision, about a third of the arithmetic.  NaNs fail the
            // comparison and take the exact test.
            const S vx = to[0] - from[0], vy = to[1] - from[1];
            const S reach = fp::sqrt_(vx * vx + vy * vy) * (S(1) + S(1024) * fp::consts<S>::eps());
            for (int c = 0; c < nCircles; ++c) {
                const S rr = circles[3 * c + 2] + linkRadius;
                const S wx = circles[3 * c] - from[0], wy = circles[3 * c + 1] - from[1];
                const S far = reach + rr;
                if (wx * wx + wy * wy > far * far) continue;
                if (!(distPointSegmentSquared2<S>(circles + 3 * c, from, to) > rr * rr)) return false;
            }
            from[0] = to[0];
            from[1] = to[1];
        }
        return true;
    }
    __device__ __forceinline__ bool endpointsValid(const S* a, const S* b) const { return valid(a) && valid(b); }
    // :127-131  (a - b).lpNorm<Infinity>() < 0.02
    __device__ __forceinline__ bool stop(const S* a, const S* b) const {
        S m = S(0);
#pragma unroll
        for (int i = 0; i < (EXACT ? MAXD_ : nLinks); ++i) {
            const S d = fp::abs_(a[i] - b[i]);
            m = d > m ? d : m;
        }
        return m < S(0.02);
    }
    // bound of the recursion depth for the flat edge check: the ends of a node at depth k differ by maxdiff / 2^k up to
    // the rounding of k midpoints (each within an ulp of the coordinates' magnitude); 1/64 relative + that absolute slack
    __device__ __forceinline__ int levels(const S* a, const S* b) const {
        S m = S(0), mag = S(1);
#pragma unroll
        for (int i = 0; i < (EXACT ? MAXD_ : nLinks); ++i) {
            const S d = fp::abs_(a[i] - b[i]);
            m = d > m ? d : m;
            mag = fmax(mag, fmax(fp::abs_(a[i]), fp::abs_(b[i])));
        }
        // an infinite coordinate: the end it belongs to is invalid (its sines are NaN and no circle test passes on NaN),
        // which the reference finds before it bisects (:118-123); said here without a 2^24-node tree
        if (!(m < fp::consts<S>::inf())) return -1;
        return levelsFor<S>(m, S(0.02) * S(63.0 / 64.0) - S(256) * fp::consts<S>::eps() * mag);
    }
};

// ------------------------------------------------------------------ bisection edge kernel
constexpr int SPLIT_LEVELS = 5;  // 32 lanes = 32 subtrees at depth 5

template <typename S, typename V>
__global__ void __launch_bounds__(256) bisectLinkKernel(const V v, const S* __restrict__ from, const S* __restrict__ to,
                                                        uint32_t n, uint8_t* __restrict__ ok,
                                                        unsigned long long* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
    const int D = v.dims();
    unsigned long long probes = 0;
    bool overflow = false;
    for (uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < n; e += warpsPerGrid) {
        S a[V::MAXD], b[V::MAXD], mid[V::MAXD];
        for (int i = 0; i < D; ++i) {
            a[i] = from[(size_t)e * D + i];
            b[i] = to[(size_t)e * D + i];
        }
        // endpoints: lanes 0-15 probe a, lanes 16-31 probe b (one probe latency instead of two)
        bool good = true;
        if (V::CHECK_ENDS) {  // (the Nao scenario's link assumes valid ends and does not look at them)
            for (int i = 0; i < D; ++i) mid[i] = lane < 16 ? a[i] : b[i];
            good = __all_sync(FULL_MASK_, v.valid(mid));
        }
        // Phase A -- the 31 midpoints of the five top levels, ONE per lane: lane j owns heap node j+1
        // (depth floor(log2(j+1))), walks down to it computing midpoints only, and probes just that
        // node.  A node exists iff none of its ancestors met the stop test.
        if (good) {
            const int h = lane + 1;
            const int d = 31 - __clz(h);  // lane 31 (h = 32) has no node in this phase
            bool exists = lane < 31;
            for (int level = 0; level < d && exists; ++level) {
                if (v.stop(a, b)) {
                    exists = false;
                    break;
                }
                for (int i = 0; i < D; ++i) mid[i] = (a[i] + b[i]) / S(2);
                if ((h >> (d - 1 - level)) & 1) {
                    for (int i = 0; i < D; ++i) a[i] = mid[i];
                } else {
                    for (int i = 0; i < D; ++i) b[i] = mid[i];
                }
            }
            bool mineOk = true;
            if (exists && !v.stop(a, b)) {
                for (int i = 0; i < D; ++i) mid[i] = (a[i] + b[i]) / S(2);
                ++probes;
                mineOk = v.valid(mid);
            }
            good = __all_sync(FULL_MASK_, mineOk);
        }
        // Phase B -- lane L descends (midpoints only, they were probed in phase A) to the root of
        // subtree L at depth five and then runs that subtree depth first.
        bool done = !good;
        if (!done) {
            for (int i = 0; i < D; ++i) {
                a[i] = from[(size_t)e * D + i];
                b[i] = to[(size_t)e * D + i];
            }
            for (int level = 0; level < SPLIT_LEVELS; ++level) {
                if (v.stop(a, b)) {
                    done = true;  // the tree ends above depth five on this path: nothing below to check
                    break;
                }
                for (int i = 0; i < D; ++i) mid[i] = (a[i] + b[i]) / S(2);
                if ((lane >> (SPLIT_LEVELS - 1 - level)) & 1) {
                    for (int i = 0; i < D; ++i) a[i] = mid[i];
                } else {
                    for (int i = 0; i < D; ++i) b[i] = mid[i];
                }
            }
        }
        // Lanes whose path stopped early duplicate a sibling's (empty) subtree: nothing left to do.
        // Depth-first over the remaining subtree.  Only the right end of each pending right half is
        // stacked: its left end is always the right end of the leaf just finished (the rightmost
        // leaf of a left subtree ends exactly at the parent's midpoint).
        S stackB[V::MAXDEPTH][V::MAXD];
        int sp = 0;
        bool havePair = !done;
        while (true) {
            const unsigned bad = __ballot_sync(FULL_MASK_, !good);
            if (bad) {
                good = false;
                break;
            }
            if (!__any_sync(FULL_MASK_, havePair)) break;
            if (havePair) {
                if (v.stop(a, b)) {
                    // leaf: continue with the next pending right half (current b, stacked b)
                    if (sp > 0) {
                        --sp;
                        for (int i = 0; i < D; ++i) {
                            a[i] = b[i];
                            b[i] = stackB[sp][i];
                        }
                    } else {
                        havePair = false;
                    }
                } else {
                    for (int i = 0; i < D; ++i) mid[i] = (a[i] + b[i]) / S(2);
                    ++probes;
                    if (!v.valid(mid)) {
                        good = false;
                    } else if (sp >= V::MAXDEPTH) {
                        overflow = true;
                        good = false;
                    } else {
                        // remember the right half (mid, b), continue with the left half (a, mid)
                        for (int i = 0; i < D; ++i) {
                            stackB[sp][i] = b[i];
                            b[i] = mid[i];
                        }
                        ++sp;
                    }
                }
            }
        }
        if (lane == 0) ok[e] = good ? 1 : 0;
    }
    // counters
    for (int o = 16; o > 0; o >>= 1) probes += __shfl_down_sync(FULL_MASK_, probes, o);
    if (lane == 0 && stats) {
        atomicAdd(stats + 2, probes);
    }
    if (overflow && stats) atomicOr(stats + 4, (unsigned long long)GEOM_ERR_STACK);
}

template <typename S, typename V>
__global__ void validKernel(const V v, const S* __restrict__ states, uint32_t n, uint8_t* __restrict__ ok) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    S q[V::MAXD];
    const int D = v.dims();
    for (int c = 0; c < D; ++c) q[c] = states[(size_t)i * D + c];
    ok[i] = v.valid(q) ? 1 : 0;
}

// ------------------------------------------------------------------ flat edge check (expensive validators)
// bisectLinkKernel gives an edge to a warp.  That suits probes of a few instructions (a grid cell); for the Nao scenario
// (one probe = two arms of forward kinematics and 209 pair tests, ~6.5 K operations) and planner-size edges (5 - 60
// midpoints) it left most lanes idle: 149 M midpoints/s against 2.4 G states/s of validKernel on the same validator.
// Here ALL midpoints of ALL edges of the batch form one flat list -- edge e contributes its ends (validators that check
// them) and the heap nodes 1 .. 2^L - 1 of its recursion tree, L a conservative bound of the tree's depth -- and a thread
// takes one item: it finds the edge by binary search in the prefix sums, reaches its node by midpoint arithmetic alone
// (the reference's (a+b)/2 sequence and stop test at every level, so the node exists exactly when the reference's
// recursion creates it) and probes it.  A failed probe clears ok[e]; items of an edge already cleared are skipped.
// `coarse`: the most tree nodes an edge contributes to the FIRST list (heap nodes 1 .. coarse); the rest, 2^L - 1 - coarse
// per edge, form a second list that is built after the first has been checked -- from the edges that are still valid
// (flatPlanRestKernel).  Large batches use coarse = 7 (the three top levels), see flatLink.
template <typename S, typename V>
__global__ void flatPlanKernel(const V v, const S* __restrict__ from, const S* __restrict__ to, uint32_t n, uint32_t coarse,
                               unsigned long long* __restrict__ counts, int8_t* __restrict__ levelsOut, uint8_t* __restrict__ ok,
                               unsigned long long* __restrict__ stats) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (e == 0) counts[n] = 0;
    S a[V::MAXD], b[V::MAXD];
    const int D = v.dims();
    for (int i = 0; i < D; ++i) a[i] = from[(size_t)e * D + i], b[i] = to[(size_t)e * D + i];
    const int L = v.levels(a, b);
    levelsOut[e] = (int8_t)(L < 0 || L > FLAT_MAX_LEVELS ? 0 : L);
    if (L < 0) {  // the validator declares the edge invalid outright (a length that is not finite)
        ok[e] = 0;
        counts[e] = 0;
        return;
    }
    if (L > FLAT_MAX_LEVELS) {
        atomicOr(stats + 4, (unsigned long long)GEOM_ERR_STEPS);
        ok[e] = 0;
        counts[e] = 0;
        return;
    }
    ok[e] = 1;
    const unsigned long long nodes = (1ull << L) - 1ull;
    counts[e] = (V::CHECK_ENDS ? 2ull : 0ull) + (nodes < coarse ? nodes : (unsigned long long)coarse);
}
// second list: tree nodes coarse + 1 .. 2^L - 1 of the edges the first list left valid
__global__ void flatPlanRestKernel(uint32_t n, uint32_t coarse, const int8_t* __restrict__ levels, const uint8_t* __restrict__ ok,
                                   unsigned long long* __restrict__ counts) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (e == 0) counts[n] = 0;
    const unsigned long long nodes = (1ull << levels[e]) - 1ull;
    counts[e] = ok[e] && nodes > coarse ? nodes - coarse : 0ull;
}

template <typename S, typename V>
__global__ void __launch_bounds__(256) flatLinkKernel(const V v, const S* __restrict__ from, const S* __restrict__ to, uint32_t n,
                                                      const unsigned long long* __restrict__ offs, uint8_t* ok,
                                                      unsigned long long* __restrict__ stats, bool withEnds, uint32_t firstNode) {
    const unsigned long long total = offs[n];
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const int D = v.dims();
    unsigned long long probes = 0;
    // The trip count is the same for the 32 lanes of a warp and every lane reaches the __syncwarp() in front of the
    // probe: a first version that left the loop body with `continue` had its lanes drift apart for good (independent
    // thread scheduling does not rejoin them at the back edge) and ran the probe with 5.9 of 32 lanes on average.
    for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x - lane); base < total; base += stride) {
        const unsigned long long it = base + lane;
        bool live = it < total;
        uint32_t e = 0;
        S a[V::MAXD], b[V::MAXD], mid[V::MAXD];
        if (live) {
            uint32_t lo = 0, hi = n;  // largest e with offs[e] <= it
            while (hi - lo > 1) {
                const uint32_t m = (lo + hi) >> 1;
                if (__ldg(offs + m) <= it) lo = m;
                else hi = m;
            }
            e = lo;
            live = *(volatile uint8_t*)(ok + e) != 0;
        }
        if (live) {
            unsigned long long j = it - __ldg(offs + e);
            for (int i = 0; i < D; ++i) a[i] = __ldg(from + (size_t)e * D + i), b[i] = __ldg(to + (size_t)e * D + i);
            if (V::CHECK_ENDS && withEnds && j < 2) {
                for (int i = 0; i < D; ++i) mid[i] = j == 0 ? a[i] : b[i];
            } else {
                if (V::CHECK_ENDS && withEnds) j -= 2;
                const unsigned long long h = j + firstNode;  // heap index: 1 = the edge's own midpoint
                const int d = 63 - __clzll(h);
                for (int level = 0; level < d; ++level) {
                    if (v.stop(a, b)) {
                        live = false;
                        break;
                    }
                    for (int i = 0; i < D; ++i) mid[i] = (a[i] + b[i]) / S(2);
                    if ((h >> (d - 1 - level)) & 1) {
                        for (int i = 0; i < D; ++i) a[i] = mid[i];
                    } else {
                        for (int i = 0; i < D; ++i) b[i] = mid[i];
                    }
                }
                if (live && v.stop(a, b)) live = false;
                for (int i = 0; i < D; ++i) mid[i] = (a[i] + b[i]) / S(2);
            }
        }
        __syncwarp();
        if (live) {
            ++probes;
            if (!v.valid(mid)) ok[e] = 0;
        }
        __syncwarp();
    }
    for (int o = 16; o > 0; o >>= 1) probes += __shfl_down_sync(FULL_MASK_, probes, o);
    if (lane == 0 && stats && probes) atomicAdd(stats + 2, probes);
}

// ------------------------------------------------------------------ balls (any dimension)
template <typename S>
struct BallsData {
    const S* balls;  // nBalls * (dim + 1)
    int nBalls, dim;
    const S* rects;
    int nRects;
};

// shape_hierarchy.hpp:222-226 + rect point test; holonomic_2d_point_scenario.hpp:95-103
template <typename S>
__global__ void shapesValidKernel(BallsData<S> g, const S* __restrict__ states, uint32_t n, uint8_t* __restrict__ ok) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const S* p = states + (size_t)i * g.dim;
    bool good = true;
    for (int j = 0; j < g.nBalls && good; ++j) {
        const S* c = g.balls + (size_t)j * (g.dim + 1);
        S acc = S(0);
        for (int d = 0; d < g.dim; ++d) {
            const S e = p[d] - c[d];
            acc = (d == 0) ? e * e : acc + e * e;
        }
        good = acc > c[g.dim] * c[g.dim];
    }
    for (int j = 0; j < g.nRects && good; ++j) {
        const S* r = g.rects + 4 * j;
        good = !(p[0] >= r[0] && p[0] <= r[2] && p[1] >= r[1] && p[1] <= r[3]);
    }
    ok[i] = good ? 1 : 0;
}

// Circle::segmentIsValid (shape_hierarchy.hpp:228-231,259-270), generalised to `dim` coordinates
// (test/planner_integration_test.hpp:143-149 uses the 3-D form).  One thread per edge.
// ok[] is AND-ed with the result (the rect bisection kernel may have written it first).
template <typename S>
__global__ void ballsLinkKernel(BallsData<S> g, const S* __restrict__ from, const S* __restrict__ to, uint32_t n,
                                uint8_t* __restrict__ ok, int andWithExisting) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const S* s0 = from + (size_t)i * g.dim;
    const S* s1 = to + (size_t)i * g.dim;
    bool good = true;
    for (int j = 0; j < g.nBalls && good; ++j) {
        const S* pt = g.balls + (size_t)j * (g.dim + 1);
        S c1 = S(0), c2 = S(0), ww = S(0);
        for (int d = 0; d < g.dim; ++d) {
            const S vv = s1[d] - s0[d], w = pt[d] - s0[d];
            c1 = (d == 0) ? vv * w : c1 + vv * w;
            c2 = (d == 0) ? vv * vv : c2 + vv * vv;
            ww = (d == 0) ? w * w : ww + w * w;
        }
        S dist2;
        if (c1 <= S(0)) {
            dist2 = ww;
        } else if (c2 <= c1) {
            S acc = S(0);
            for (int d = 0; d < g.dim; ++d) {
                const S e = pt[d] - s1[d];
                acc = (d == 0) ? e * e : acc + e * e;
            }
            dist2 = acc;
        } else {
            const S f = fp::div_(c1, c2);
            S acc = S(0);
            for (int d = 0; d < g.dim; ++d) {
                const S vv = s1[d] - s0[d];
                const S e = s0[d] - pt[d] + vv * f;
                acc = (d == 0) ? e * e : acc + e * e;
            }
            dist2 = acc;
        }
        good = dist2 > pt[g.dim] * pt[g.dim];
    }
    if (andWithExisting) good = good && ok[i];
    ok[i] = good ? 1 : 0;
}

}  // namespace mptg

using namespace mptg;

namespace {

int warpGrid(mptg_ctx* ctx, uint32_t n) {
    // persistent-style: enough 8-warp CTAs to fill the machine, never more than one warp per edge
    const uint32_t want = (n + 7) / 8;
    const uint32_t cap = (uint32_t)ctx->smCount * 8;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

template <typename S>
int uploadScalars(mptg_ctx* ctx, const double* src, size_t count, void** dst) {
    std::vector<S> tmp(count);
    for (size_t i = 0; i < count; ++i) tmp[i] = (S)src[i];
    MPTG_CUDA(ctx, cudaMalloc(dst, (count ? count : 1) * sizeof(S)));
    return uploadSync(ctx, *dst, tmp.data(), count * sizeof(S));
}

int newGeom(mptg_ctx* ctx, int kind, int scalar, mptg_geom** out) {
    auto* g = new mptg_geom();
    g->ctx = ctx;
    g->kind = kind;
    g->scalar = scalar;
    cudaError_t e = cudaMalloc(&g->devStats, 8 * sizeof(unsigned long long));
    if (e != cudaSuccess || memsetSync(ctx, g->devStats, 0, 8 * sizeof(unsigned long long)) != MPTG_OK) {
        if (e == cudaSuccess) cudaFree(g->devStats);
        delete g;
        return fail(ctx, MPTG_ERR_CUDA, "geometry create: %s", e != cudaSuccess ? cudaGetErrorString(e) : "memset failed");
    }
    *out = g;
    return MPTG_OK;
}

template <typename S>
int validDevT(mptg_geom* g, const S* states, uint32_t n, uint8_t* ok) {
    mptg_ctx* ctx = g->ctx;
    const dim3 grid((n + 127) / 128), block(128);
    switch (g->kind) {
        case MPTG_GEOM_GRID: {
            GridValidator<S> v{g->gridBits, g->width, g->height};
            validKernel<S, GridValidator<S>><<<grid, block, 0, ctx->stream>>>(v, states, n, ok);
            break;
        }
        case MPTG_GEOM_SHAPES: {
            BallsData<S> b{(const S*)g->balls, g->nBalls, g->dim, (const S*)g->rects, g->nRects};
            shapesValidKernel<S><<<grid, block, 0, ctx->stream>>>(b, states, n, ok);
            break;
        }
        case MPTG_GEOM_LINKARM: {
#define MPTG_ARM_VALID(MD)                                                                                         \
    {                                                                                                              \
        ArmValidator<S, MD> v{(const S*)g->lengths, (const S*)g->circles, g->nLinks, g->nCircles, (S)g->linkRadius}; \
        validKernel<S, ArmValidator<S, MD>><<<grid, block, 0, ctx->stream>>>(v, states, n, ok);                    \
    }
            if (g->nLinks <= 8) MPTG_ARM_VALID(8)
            else if (g->nLinks <= 16) MPTG_ARM_VALID(16)
            else if (g->nLinks <= 32) MPTG_ARM_VALID(32)
            else MPTG_ARM_VALID(64)
#undef MPTG_ARM_VALID
            break;
        }
        case MPTG_GEOM_NAOCUP: {
            nao::Validator<S> v{*(const nao::Model<S>*)g->naoModel};
            validKernel<S, nao::Validator<S>><<<grid, block, 0, ctx->stream>>>(v, states, n, ok);
            break;
        }
        default: return fail(ctx, MPTG_ERR_BAD_ARG, "valid: unknown geometry kind %d", g->kind);
    }
    MPTG_LAUNCHED(ctx);
    return MPTG_OK;
}

// flat edge check: plan (levels per edge) -> prefix sums (cub, library code) -> one item per thread
template <typename S, typename V>
int flatLink(mptg_geom* g, const V& v, const S* from, const S* to, uint32_t n, uint8_t* ok) {
    mptg_ctx* ctx = g->ctx;
    size_t scanBytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)(n + 1u));
    const size_t listBytes = ((size_t)(n + 1u) * sizeof(unsigned long long) + 255) & ~(size_t)255;
    const size_t levBytes = ((size_t)n + 255) & ~(size_t)255;
    const size_t want = 2 * listBytes + levBytes + scanBytes;
    if (g->flatBytes < want) {
        MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (g->flatBuf) MPTG_CUDA(ctx, cudaFree(g->flatBuf));
        g->flatBuf = nullptr, g->flatBytes = 0;
        MPTG_CUDA(ctx, cudaMalloc(&g->flatBuf, want + want / 2));
        g->flatBytes = want + want / 2;
    }
    auto* counts = (unsigned long long*)g->flatBuf;
    auto* offs = (unsigned long long*)((char*)g->flatBuf + listBytes);
    auto* levels = (int8_t*)((char*)g->flatBuf + 2 * listBytes);
    void* temp = (char*)g->flatBuf + 2 * listBytes + levBytes;
    // Large batches go coarse to fine: the three top levels of every edge first (7 midpoints and the ends), then the deeper
    // nodes of the edges that survived -- an invalid edge (most of a roadmap planner's long candidate edges) rarely gets
    // past the first list, as it rarely gets past its first midpoints in the reference's recursion.  Small batches (the
    // waves of a young tree) take one list: three launches less.
    const uint32_t coarse = n >= 4096 ? 7u : 0xFFFFFFFFu;
    int perSm = 0;
    MPTG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, flatLinkKernel<S, V>, 256, 0));
    const unsigned grid = (unsigned)ctx->smCount * (perSm > 0 ? perSm : 1);
    flatPlanKernel<S, V><<<(n + 255) / 256, 256, 0, ctx->stream>>>(v, from, to, n, coarse, counts, levels, ok, g->devStats);
    MPTG_LAUNCHED(ctx);
    MPTG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(temp, scanBytes, counts, offs, (int)(n + 1u), ctx->stream));
    MPTG_LAUNCHED(ctx);
    // the length of a list is only known on the device: a resident grid strides over it
    flatLinkKernel<S, V><<<grid, 256, 0, ctx->stream>>>(v, from, to, n, offs, ok, g->devStats, true, 1u);
    MPTG_LAUNCHED(ctx);
    if (coarse != 0xFFFFFFFFu) {
        flatPlanRestKernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(n, coarse, levels, ok, counts);
        MPTG_LAUNCHED(ctx);
        MPTG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(temp, scanBytes, counts, offs, (int)(n + 1u), ctx->stream));
        MPTG_LAUNCHED(ctx);
        flatLinkKernel<S, V><<<grid, 256, 0, ctx->stream>>>(v, from, to, n, offs, ok, g->devStats, false, coarse + 1u);
        MPTG_LAUNCHED(ctx);
    }
    return MPTG_OK;
}

template <typename S>
int linkDevT(mptg_geom* g, const S* from, const S* to, uint32_t n, uint8_t* ok) {
    mptg_ctx* ctx = g->ctx;
    const dim3 wgrid(warpGrid(ctx, n)), wblock(256);
    switch (g->kind) {
        case MPTG_GEOM_GRID: {
            GridValidator<S> v{g->gridBits, g->width, g->height};
            bisectLinkKernel<S, GridValidator<S>><<<wgrid, wblock, 0, ctx->stream>>>(v, from, to, n, ok, g->devStats);
            MPTG_LAUNCHED(ctx);
            break;
        }
        case MPTG_GEOM_SHAPES: {
            int haveRects = 0;
            if (g->nRects > 0) {
                ShapesValidator<S> v{(const S*)g->rects, g->nRects};
                bisectLinkKernel<S, ShapesValidator<S>><<<wgrid, wblock, 0, ctx->stream>>>(v, from, to, n, ok, g->devStats);
                MPTG_LAUNCHED(ctx);
                haveRects = 1;
            }
            BallsData<S> b{(const S*)g->balls, g->nBalls, g->dim, (const S*)g->rects, g->nRects};
            ballsLinkKernel<S><<<(n + 127) / 128, 128, 0, ctx->stream>>>(b, from, to, n, ok, haveRects);
            MPTG_LAUNCHED(ctx);
            break;
        }
        case MPTG_GEOM_LINKARM: {
#define MPTG_ARM_LINK(MD)                                                                                          \
    {                                                                                                              \
        ArmValidator<S, MD> v{(const S*)g->lengths, (const S*)g->circles, g->nLinks, g->nCircles, (S)g->linkRadius}; \
        if (armFlat && g->nLinks == MD && MD <= armExactMax && MD <= 32) {                                         \
            if constexpr (MD <= 32) { /* the register form of a 64-link arm does not exist */                      \
                ArmValidator<S, MD, true> ve{(const S*)g->lengths, (const S*)g->circles, g->nLinks, g->nCircles, (S)g->linkRadius}; \
                if (int rc = flatLink<S, ArmValidator<S, MD, true>>(g, ve, from, to, n, ok)) return rc;            \
            }                                                                                                      \
        } else if (armFlat) {                                                                                      \
            if (int rc = flatLink<S, ArmValidator<S, MD>>(g, v, from, to, n, ok)) return rc;                       \
        } else {                                                                                                   \
            bisectLinkKernel<S, ArmValidator<S, MD>><<<wgrid, wblock, 0, ctx->stream>>>(v, from, to, n, ok, g->devStats); \
            MPTG_LAUNCHED(ctx);                                                                                    \
        }                                                                                                          \
    }
            static const bool armFlat = getenv("MPTG_ARM_WARP_PER_EDGE") == nullptr;  // the earlier kernel stays selectable for comparisons
            static const int armExactMax = getenv("MPTG_ARM_EXACT_MAX") ? atoi(getenv("MPTG_ARM_EXACT_MAX")) : 16;
            if (g->nLinks <= 8) MPTG_ARM_LINK(8)
            else if (g->nLinks <= 16) MPTG_ARM_LINK(16)
            else if (g->nLinks <= 32) MPTG_ARM_LINK(32)
            else MPTG_ARM_LINK(64)
#undef MPTG_ARM_LINK
            break;
        }
        case MPTG_GEOM_NAOCUP: {
            nao::Validator<S> v{*(const nao::Model<S>*)g->naoModel};
            if (int rc = flatLink<S, nao::Validator<S>>(g, v, from, to, n, ok)) return rc;
            break;
        }
        default: return fail(ctx, MPTG_ERR_BAD_ARG, "link: unknown geometry kind %d", g->kind);
    }
    return MPTG_OK;
}

int resetStats(mptg_geom* g) {
    MPTG_CUDA(g->ctx, cudaMemsetAsync(g->devStats, 0, 8 * sizeof(unsigned long long), g->ctx->stream));
    return MPTG_OK;
}

}  // namespace

extern "C" {

int mptg_grid_create(mptg_ctx* ctx, int scalar, int32_t width, int32_t height, const uint8_t* occupancy, mptg_geom** out) {
    if (!ctx || !out || !occupancy || width <= 0 || height <= 0 || (scalar != MPTG_F32 && scalar != MPTG_F64))
        return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_grid_create: bad argument");
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    mptg_geom* g;
    int rc = newGeom(ctx, MPTG_GEOM_GRID, scalar, &g);
    if (rc) return rc;
    g->D = 2;
    g->width = width;
    g->height = height;
    const size_t cells = (size_t)width * height, words = (cells + 31) / 32;
    std::vector<uint32_t> bits(words, 0u);
    for (size_t i = 0; i < cells; ++i)
        if (occupancy[i]) bits[i >> 5] |= 1u << (i & 31);
    cudaError_t e = cudaMalloc(&g->gridBits, words * 4);
    if (e != cudaSuccess) {
        mptg_geom_destroy(g);
        return fail(ctx, MPTG_ERR_CUDA, "mptg_grid_create: %s", cudaGetErrorString(e));
    }
    rc = uploadSync(ctx, g->gridBits, bits.data(), words * 4);
    if (rc) {
        mptg_geom_destroy(g);
        return rc;
    }
    *out = g;
    return MPTG_OK;
}

int mptg_shapes_create(mptg_ctx* ctx, int scalar, int32_t dim, int32_t nBalls, const double* centres, const double* radii,
                       int32_t nRects, const double* rects, mptg_geom** out) {
    if (!ctx || !out || dim < 1 || dim > MPTG_MAX_SCALARS || nBalls < 0 || nRects < 0 || (nBalls && (!centres || !radii)) ||
        (nRects && (!rects || dim != 2)) || (scalar != MPTG_F32 && scalar != MPTG_F64))
        return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_shapes_create: bad argument (rectangles need dim == 2)");
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    mptg_geom* g;
    int rc = newGeom(ctx, MPTG_GEOM_SHAPES, scalar, &g);
    if (rc) return rc;
    g->D = g->dim = dim;
    g->nBalls = nBalls;
    g->nRects = nRects;
    std::vector<double> packed((size_t)nBalls * (dim + 1));
    for (int j = 0; j < nBalls; ++j) {
        for (int d = 0; d < dim; ++d) packed[(size_t)j * (dim + 1) + d] = centres[(size_t)j * dim + d];
        packed[(size_t)j * (dim + 1) + dim] = radii[j];
    }
    rc = scalar == MPTG_F32 ? uploadScalars<float>(ctx, packed.data(), packed.size(), &g->balls)
                            : uploadScalars<double>(ctx, packed.data(), packed.size(), &g->balls);
    if (!rc)
        rc = scalar == MPTG_F32 ? uploadScalars<float>(ctx, rects, (size_t)nRects * 4, &g->rects)
                                : uploadScalars<double>(ctx, rects, (size_t)nRects * 4, &g->rects);
    if (rc) {
        mptg_geom_destroy(g);
        return rc;
    }
    *out = g;
    return MPTG_OK;
}

int mptg_linkarm_create(mptg_ctx* ctx, int scalar, int32_t nLinks, const double* lengths, double linkRadius,
                        int32_t nCircles, const double* cxcyr, mptg_geom** out) {
    if (!ctx || !out || nLinks < 1 || nLinks > MPTG_MAX_SCALARS || !lengths || nCircles < 0 || (nCircles && !cxcyr) ||
        (scalar != MPTG_F32 && scalar != MPTG_F64))
        return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_linkarm_create: bad argument");
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    mptg_geom* g;
    int rc = newGeom(ctx, MPTG_GEOM_LINKARM, scalar, &g);
    if (rc) return rc;
    g->D = g->nLinks = nLinks;
    g->nCircles = nCircles;
    g->linkRadius = linkRadius;
    rc = scalar == MPTG_F32 ? uploadScalars<float>(ctx, lengths, nLinks, &g->lengths)
                            : uploadScalars<double>(ctx, lengths, nLinks, &g->lengths);
    if (!rc)
        rc = scalar == MPTG_F32 ? uploadScalars<float>(ctx, cxcyr, (size_t)nCircles * 3, &g->circles)
                                : uploadScalars<double>(ctx, cxcyr, (size_t)nCircles * 3, &g->circles);
    if (rc) {
        mptg_geom_destroy(g);
        return rc;
    }
    *out = g;
    return MPTG_OK;
}

int mptg_mesh_pair_create(mptg_ctx* ctx, int scalar, uint32_t nr, const float* robotTris, uint32_t ne, const float* envTris,
                          mptg_geom** out) {
    if (!ctx || !out || (nr && !robotTris) || (ne && !envTris)) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_mesh_pair_create: bad argument");
    if (scalar != MPTG_F32 && scalar != MPTG_F64) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_mesh_pair_create: scalar must be MPTG_F32 or MPTG_F64");
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    mptg_geom* g;
    int rc = newGeom(ctx, MPTG_GEOM_MESH, scalar, &g);
    if (rc) return rc;
    g->D = 7;
    rc = meshCreate(ctx, scalar, nr, robotTris, ne, envTris, &g->mesh);
    if (rc) {
        mptg_geom_destroy(g);
        return rc;
    }
    *out = g;
    return MPTG_OK;
}

int mptg_naocup_create(mptg_ctx* ctx, int scalar, mptg_geom** out) {
    if (!ctx || !out || (scalar != MPTG_F32 && scalar != MPTG_F64)) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_naocup_create: bad argument");
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    mptg_geom* g;
    int rc = newGeom(ctx, MPTG_GEOM_NAOCUP, scalar, &g);
    if (rc) return rc;
    g->D = nao::DIM;
    // the model is a block of scalars evaluated on the host; it travels to the kernels as a launch parameter
    if (scalar == MPTG_F32) g->naoModel = new nao::Model<float>(nao::makeModel<float>());
    else g->naoModel = new nao::Model<double>(nao::makeModel<double>());
    *out = g;
    return MPTG_OK;
}

int mptg_naocup_configs(int scalar, double* start, double* goal, double* lo, double* hi) {
    if (scalar != MPTG_F32 && scalar != MPTG_F64) return fail(nullptr, MPTG_ERR_BAD_ARG, "mptg_naocup_configs: scalar must be MPTG_F32 or MPTG_F64");
    // naocup.hpp:254-301; limits in degrees :196-219, converted as RAD(x) = x * (PI / 180) in the scalar type
    static const double startC[10] = {1.125998, -0.691876, 1.888312, 0.776246, 0.245398, 1.259372, 0.279146, -1.587732, -0.510780, -1.823800};
    static const double goalC[10] = {0.258284303377494,  -0.2699099199363406, -0.01113121187052224, 1.2053012757652763,  1.2716626717484503,
                                     -0.9826967097045605, 0.07355836822937814, 0.25450053440459897,  -0.9512909033938429, -0.5297424293532234};
    static const double loDeg[10] = {-119.5, -94.5, -119.5, 0.5, -104.5, -119.5, 0.5, -119.5, -89.5, -104.5};
    static const double hiDeg[10] = {119.5, -0.5, 119.5, 89.5, 104.5, 119.5, 94.5, 119.5, -0.5, 104.5};
    for (int i = 0; i < 10; ++i) {
        if (scalar == MPTG_F32) {
            const float k = float(float(3.14159265358979323846) / float(180.0));
            if (start) start[i] = (double)(float)startC[i];
            if (goal) goal[i] = (double)(float)goalC[i];
            if (lo) lo[i] = (double)(float(loDeg[i]) * k);
            if (hi) hi[i] = (double)(float(hiDeg[i]) * k);
        } else {
            const double k = 3.14159265358979323846 / 180.0;
            if (start) start[i] = startC[i];
            if (goal) goal[i] = goalC[i];
            if (lo) lo[i] = loDeg[i] * k;
            if (hi) hi[i] = hiDeg[i] * k;
        }
    }
    return MPTG_OK;
}

int mptg_geom_destroy(mptg_geom* g) {
    if (!g) return MPTG_OK;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
    cudaFree(g->gridBits);
    cudaFree(g->balls);
    cudaFree(g->rects);
    cudaFree(g->lengths);
    cudaFree(g->circles);
    cudaFree(g->devStats);
    cudaFree(g->flatBuf);
    if (g->mesh) meshDestroy(g->mesh);
    if (g->naoModel) {
        if (g->scalar == MPTG_F32) delete (nao::Model<float>*)g->naoModel;
        else delete (nao::Model<double>*)g->naoModel;
    }
    delete g;
    return MPTG_OK;
}

int mptg_geom_kind(const mptg_geom* g) { return g ? g->kind : 0; }

int mptg_geom_contact_band(const mptg_geom* g, double* out) {
    if (!g || !out) return fail(g ? g->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_geom_contact_band: bad argument");
    *out = g->kind == MPTG_GEOM_MESH ? meshBand(g->mesh) : 0.0;
    return MPTG_OK;
}

int mptg_valid_batch_dev(mptg_geom* g, const void* states, uint32_t n, uint8_t* ok, uint8_t* nearOut) {
    if (!g || (n && (!states || !ok))) return fail(g ? g->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_valid_batch: bad argument");
    if (n == 0) return MPTG_OK;
    MPTG_CUDA(g->ctx, cudaSetDevice(g->ctx->device));
    int rc = resetStats(g);
    if (rc) return rc;
    if (g->kind == MPTG_GEOM_MESH) return meshValidDev(g, states, n, ok, nearOut);
    if (nearOut) MPTG_CUDA(g->ctx, cudaMemsetAsync(nearOut, 0, n, g->ctx->stream));  // bit-identical decisions: no band
    return g->scalar == MPTG_F32 ? validDevT<float>(g, (const float*)states, n, ok) : validDevT<double>(g, (const double*)states, n, ok);
}

int mptg_link_batch_dev(mptg_geom* g, const mptg_space_desc* space, const void* from, const void* to, uint32_t n, double step,
                        uint8_t* ok, uint8_t* nearOut) {
    if (!g || (n && (!from || !to || !ok))) return fail(g ? g->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_link_batch: bad argument");
    if (n == 0) return MPTG_OK;
    MPTG_CUDA(g->ctx, cudaSetDevice(g->ctx->device));
    int rc = resetStats(g);
    if (rc) return rc;
    if (g->kind == MPTG_GEOM_MESH) {
        if (!space || spaceScalars(space) != 7 || space->scalar != g->scalar || !(step > 0))
            return fail(g->ctx, MPTG_ERR_BAD_ARG, "mptg_link_batch: mesh edges need an SE(3) space of the mesh's scalar type and step > 0");
        return meshLinkDev(g, space, from, to, n, step, ok, nearOut);
    }
    if (nearOut) MPTG_CUDA(g->ctx, cudaMemsetAsync(nearOut, 0, n, g->ctx->stream));
    return g->scalar == MPTG_F32 ? linkDevT<float>(g, (const float*)from, (const float*)to, n, ok)
                                 : linkDevT<double>(g, (const double*)from, (const double*)to, n, ok);
}

static int checkDeviceErrors(mptg_geom* g) {
    // called after a synchronise: surface device-side error flags
    unsigned long long host[8];
    MPTG_CUDA(g->ctx, cudaMemcpyAsync(host, g->devStats, sizeof host, cudaMemcpyDeviceToHost, g->ctx->stream));
    MPTG_CUDA(g->ctx, cudaStreamSynchronize(g->ctx->stream));
    for (int i = 0; i < 4; ++i) g->stats[i] = host[i];
    if (getenv("MPTG_DEBUG_STATS"))
        fprintf(stderr, "[mptg] geom stats: states %llu bv %llu prim %llu items %llu | dbg %llu %llu %llu\n", host[0], host[1], host[2],
                host[3], host[5], host[6], host[7]);
    if (host[4] & GEOM_ERR_STACK) return fail(g->ctx, MPTG_ERR_CAPACITY, "edge check: traversal stack overflow (edge too long / geometry too deep)");
    if (host[4] & GEOM_ERR_STEPS) return fail(g->ctx, MPTG_ERR_CAPACITY, "edge check: too many interpolation steps on one edge");
    if (host[4] & GEOM_ERR_SCHED) return fail(g->ctx, MPTG_ERR_CUDA, "mesh check: a warp waited for donated work for seconds (scheduling fault); results are not valid");
    return MPTG_OK;
}

int mptg_valid_batch(mptg_geom* g, const void* states, uint32_t n, uint8_t* ok, uint8_t* nearOut) {
    if (!g || (n && (!states || !ok))) return fail(g ? g->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_valid_batch: bad argument");
    if (n == 0) return MPTG_OK;
    mptg_ctx* ctx = g->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t sb = (size_t)n * g->D * g->scalar;
    void* dIn;
    void* dOut;
    int rc = scratch(ctx, 0, sb, &dIn);
    if (rc) return rc;
    rc = scratch(ctx, 1, 2 * (size_t)n, &dOut);
    if (rc) return rc;
    uint8_t* dNear = nearOut ? (uint8_t*)dOut + n : nullptr;
    MPTG_CUDA(ctx, cudaMemcpyAsync(dIn, states, sb, cudaMemcpyHostToDevice, ctx->stream));
    rc = mptg_valid_batch_dev(g, dIn, n, (uint8_t*)dOut, dNear);
    if (rc) return rc;
    MPTG_CUDA(ctx, cudaMemcpyAsync(ok, dOut, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (nearOut) MPTG_CUDA(ctx, cudaMemcpyAsync(nearOut, dNear, n, cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return checkDeviceErrors(g);
}

int mptg_link_batch(mptg_geom* g, const mptg_space_desc* space, const void* from, const void* to, uint32_t n, double step,
                    uint8_t* ok, uint8_t* nearOut) {
    if (!g || (n && (!from || !to || !ok))) return fail(g ? g->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_link_batch: bad argument");
    if (n == 0) return MPTG_OK;
    mptg_ctx* ctx = g->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t sb = (size_t)n * g->D * g->scalar, sbp = (sb + 255) & ~(size_t)255;
    void* dIn;
    void* dOut;
    int rc = scratch(ctx, 0, 2 * sbp, &dIn);
    if (rc) return rc;
    rc = scratch(ctx, 1, 2 * (size_t)n, &dOut);
    if (rc) return rc;
    void* dTo = (char*)dIn + sbp;
    uint8_t* dNear = nearOut ? (uint8_t*)dOut + n : nullptr;
    MPTG_CUDA(ctx, cudaMemcpyAsync(dIn, from, sb, cudaMemcpyHostToDevice, ctx->stream));
    MPTG_CUDA(ctx, cudaMemcpyAsync(dTo, to, sb, cudaMemcpyHostToDevice, ctx->stream));
    rc = mptg_link_batch_dev(g, space, dIn, dTo, n, step, (uint8_t*)dOut, dNear);
    if (rc) return rc;
    MPTG_CUDA(ctx, cudaMemcpyAsync(ok, dOut, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (nearOut) MPTG_CUDA(ctx, cudaMemcpyAsync(nearOut, dNear, n, cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return checkDeviceErrors(g);
}

int mptg_geom_last_stats(mptg_geom* g, uint64_t out[4]) {
    if (!g || !out) return fail(g ? g->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_geom_last_stats: bad argument");
    MPTG_CUDA(g->ctx, cudaSetDevice(g->ctx->device));
    MPTG_CUDA(g->ctx, cudaStreamSynchronize(g->ctx->stream));
    int rc = checkDeviceErrors(g);
    for (int i = 0; i < 4; ++i) out[i] = g->stats[i];
    return rc;
}
}
