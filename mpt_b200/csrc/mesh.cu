// mesh.cu -- rigid-body mesh-vs-mesh validity for SE(3) states and edges (SURVEY.md section 8 rows
// a6, a7).  Replaces fcl::collide(robotBVH, T(q), envBVH, I) under the default request
// (demo/se3_rigid_body_scenario.hpp:282-296) and DiscreteMotionValidator::operator()
// (src/mpt/discrete_motion_validator.hpp:71-130) for that scenario.
//
// Decision being computed (same definition as the CPU oracle; FCL itself is not available):
//   collide(q) = exists (i,j): aabb(T(q) robotTri_i) overlaps aabb(envTri_j)   [closed intervals]
//                              && the 17-axis separating-axis test finds no strictly separating axis.
// Node bounds only prune (they are padded to stay conservative), so the answer does not depend on
// the hierarchy.  Vertex transform and the SAT use plain, unfused arithmetic in a fixed order; the
// box tests use fused arithmetic freely.
//
// Execution: one warp per work item (a state, or an edge whose states are visited in the
// reference's bisection order with its early exit).  Warps pull items from a global counter, so
// long and short items balance.  The simultaneous descent of the two binary AABB trees is
// breadth-limited: the warp keeps a stack of node pairs in shared memory, each round the 32 lanes
// test up to 32 pairs from the top and push the surviving children; leaf-leaf survivors go to a
// triangle-pair queue that is drained 32 at a time so the SAT runs without divergence.
#include "geom.cuh"

namespace mptg {

struct __align__(16) BvhNode {
    float lo[3];
    int left;  // >= 0: internal (children left,right); < 0: leaf holding triangle (-1 - left)
    float hi[3];
    int right;
};
struct __align__(16) TriPad {
    float v[3][4];
};

struct MeshDev {
    const BvhNode* rNodes;
    const TriPad* rTris;
    const BvhNode* eNodes;
    const TriPad* eTris;
    uint32_t nR, nE;  // triangle counts
};

}  // namespace mptg

struct MeshData {
    mptg::MeshDev dev{};
    void* mem[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned int* workCounter = nullptr;
    int depthR = 0, depthE = 0;
};

namespace mptg {

constexpr int MESH_WARPS = 4;
constexpr int MESH_MIN_CTAS = 8;  // 32 warps per SM: caps registers at 64
constexpr int NODE_STACK = 512;
constexpr int TRI_QUEUE = 64;
constexpr int DMV_QUEUE = 256;  // fixedBisectQueueSize_, discrete_motion_validator.hpp:54

struct WarpCounters {
    unsigned int bv = 0, tri = 0, states = 0;  // per lane (bv, tri) / per warp (states); summed into 64-bit totals at exit
};

// Eigen quaternion -> rotation matrix operation order (same as the oracle)
__device__ __forceinline__ void quatToRot(const float* q, float R[9]) {
    const float x = q[0], y = q[1], z = q[2], w = q[3];
    const float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
    const float twx = tx * w, twy = ty * w, twz = tz * w;
    const float txx = tx * x, txy = ty * x, txz = tz * x;
    const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1.0f - (tyy + tzz);
    R[1] = txy - twz;
    R[2] = txz + twy;
    R[3] = txy + twz;
    R[4] = 1.0f - (txx + tzz);
    R[5] = tyz - twx;
    R[6] = txz - twy;
    R[7] = tyz + twx;
    R[8] = 1.0f - (txx + tyy);
}

__device__ __forceinline__ void xformPoint(const float R[9], const float t[3], const float* v, float out[3]) {
    out[0] = ((R[0] * v[0] + R[1] * v[1]) + R[2] * v[2]) + t[0];
    out[1] = ((R[3] * v[0] + R[4] * v[1]) + R[5] * v[2]) + t[1];
    out[2] = ((R[6] * v[0] + R[7] * v[1]) + R[8] * v[2]) + t[2];
}

__device__ __forceinline__ void cross3(const float a[3], const float b[3], float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ float dot3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

// true when the projections on `ax` overlap or touch
__device__ __forceinline__ bool project6(const float ax[3], const float p[3][3], const float q[3][3]) {
    const float P0 = dot3(ax, p[0]), P1 = dot3(ax, p[1]), P2 = dot3(ax, p[2]);
    const float Q0 = dot3(ax, q[0]), Q1 = dot3(ax, q[1]), Q2 = dot3(ax, q[2]);
    const float mx1 = fmaxf(P0, fmaxf(P1, P2)), mn1 = fminf(P0, fminf(P1, P2));
    const float mx2 = fmaxf(Q0, fmaxf(Q1, Q2)), mn2 = fminf(Q0, fminf(Q1, Q2));
    if (mn1 > mx2) return false;
    if (mn2 > mx1) return false;
    return true;
}

// 17-axis SAT (PQP TriContact / FCL Intersect::intersect_Triangle scheme), coordinates relative to P[0]
__device__ __noinline__ bool triTriIntersect(const float P[3][3], const float Qt[3][3]) {
    float p[3][3], q[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            p[i][c] = P[i][c] - P[0][c];
            q[i][c] = Qt[i][c] - P[0][c];
        }
    float e[3][3], f[3][3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        e[0][c] = p[1][c] - p[0][c];
        e[1][c] = p[2][c] - p[1][c];
        e[2][c] = p[0][c] - p[2][c];
        f[0][c] = q[1][c] - q[0][c];
        f[1][c] = q[2][c] - q[1][c];
        f[2][c] = q[0][c] - q[2][c];
    }
    float n1[3], m1[3], ax[3];
    cross3(e[0], e[1], n1);
    if (!project6(n1, p, q)) return false;
    cross3(f[0], f[1], m1);
    if (!project6(m1, p, q)) return false;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            cross3(e[i], f[j], ax);
            if (!project6(ax, p, q)) return false;
        }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        cross3(e[i], n1, ax);
        if (!project6(ax, p, q)) return false;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        cross3(f[i], m1, ax);
        if (!project6(ax, p, q)) return false;
    }
    return true;
}

__device__ __forceinline__ BvhNode loadNode(const BvhNode* nodes, int i) {
    const float4* p = reinterpret_cast<const float4*>(nodes + i);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    BvhNode n;
    n.lo[0] = a.x, n.lo[1] = a.y, n.lo[2] = a.z, n.left = __float_as_int(a.w);
    n.hi[0] = b.x, n.hi[1] = b.y, n.hi[2] = b.z, n.right = __float_as_int(b.w);
    return n;
}

// Collision test of up to MESH_SLOTS rigid-body states at once by a full warp.  The states' transforms
// (R row-major 9 floats, t 3 floats) sit in shared memory, xf[slot][12]; every stack / queue entry
// carries its slot in the top 4 bits of the robot index, so the 32 lanes always draw from one mixed
// frontier and stay busy while a single state's frontier is still narrow.  Returns the mask of slots
// found in collision; with stopAtFirst it returns as soon as any slot collides (edge checks: one
// invalid state invalidates the edge).
// err: set to GEOM_ERR_STACK if the pair stack would overflow.
constexpr int MESH_SLOTS = 8;
constexpr unsigned SLOT_SHIFT = 28;
constexpr unsigned NODE_MASK = (1u << SLOT_SHIFT) - 1u;

__device__ unsigned warpCollideMulti(const MeshDev& m, int nSlots, const float (*xf)[12], uint2* stack, uint2* triQ, int lane,
                                     bool stopAtFirst, WarpCounters& cnt, unsigned long long& err) {
    if (m.nR == 0 || m.nE == 0 || nSlots == 0) return 0u;
    const unsigned allMask = (1u << nSlots) - 1u;
    unsigned hitMask = 0u;
    int n = nSlots, nt = 0;
    if (lane < nSlots) stack[lane] = make_uint2((unsigned)lane << SLOT_SHIFT, 0u);
    __syncwarp();
    const unsigned ltMask = (1u << lane) - 1u;
    while (n > 0 || nt > 0) {
        if (n > 0) {
            const int p = (n > NODE_STACK - 160) ? 1 : (n < 32 ? n : 32);
            bool mine = lane < p;
            uint2 pr = make_uint2(0u, 0u);
            if (mine) pr = stack[n - 1 - lane];
            __syncwarp();
            n -= p;
            const unsigned slot = pr.x >> SLOT_SHIFT;
            mine = mine && !((hitMask >> slot) & 1u);  // pairs of a state already known to collide are dropped
            int kind = 0;  // 1: triangle pair, 2: expand robot node, 3: expand env node
            int c0 = 0, c1 = 0;
            if (mine) {
                const BvhNode a = loadNode(m.rNodes, (int)(pr.x & NODE_MASK));
                const BvhNode b = loadNode(m.eNodes, (int)pr.y);
                const float* X = xf[slot];
                ++cnt.bv;
                // world AABB of the rotated local box: centre +- |R| h, padded
                const float cx = 0.5f * (a.lo[0] + a.hi[0]), cy = 0.5f * (a.lo[1] + a.hi[1]), cz = 0.5f * (a.lo[2] + a.hi[2]);
                const float hx = 0.5f * (a.hi[0] - a.lo[0]), hy = 0.5f * (a.hi[1] - a.lo[1]), hz = 0.5f * (a.hi[2] - a.lo[2]);
                bool ov = true;
                float mag = 0.0f;
                float cw[3], hw[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float r0 = X[3 * r], r1 = X[3 * r + 1], r2 = X[3 * r + 2];
                    cw[r] = __fmaf_rn(r2, cz, __fmaf_rn(r1, cy, __fmaf_rn(r0, cx, X[9 + r])));
                    hw[r] = __fmaf_rn(fabsf(r2), hz, __fmaf_rn(fabsf(r1), hy, fabsf(r0) * hx));
                    mag += fabsf(cw[r]) + hw[r];
                }
                const float pad = 64.0f * 1.1920928955078125e-07f * mag;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const float lo = cw[r] - hw[r] - pad, hi = cw[r] + hw[r] + pad;
                    ov = ov && !(lo > b.hi[r] || b.lo[r] > hi);
                }
                if (ov) {
                    const bool leafA = a.left < 0, leafB = b.left < 0;
                    if (leafA && leafB) {
                        kind = 1;
                        c0 = -1 - a.left;
                        c1 = -1 - b.left;
                    } else {
                        bool descendRobot;
                        if (leafA) descendRobot = false;
                        else if (leafB) descendRobot = true;
                        else {
                            const float ea = 2.0f * fmaxf(hx, fmaxf(hy, hz));
                            const float eb = fmaxf(b.hi[0] - b.lo[0], fmaxf(b.hi[1] - b.lo[1], b.hi[2] - b.lo[2]));
                            descendRobot = ea > eb;
                        }
                        if (descendRobot) {
                            kind = 2;
                            c0 = a.left;
                            c1 = a.right;
                        } else {
                            kind = 3;
                            c0 = b.left;
                            c1 = b.right;
                        }
                    }
                }
            }
            const unsigned tag = slot << SLOT_SHIFT;
            const unsigned mt = __ballot_sync(FULL_MASK_, kind == 1);
            const unsigned mx = __ballot_sync(FULL_MASK_, kind >= 2);
            if (kind == 1) triQ[nt + __popc(mt & ltMask)] = make_uint2((unsigned)c0 | tag, (unsigned)c1);
            if (kind >= 2) {
                const int o = n + 2 * __popc(mx & ltMask);
                if (kind == 2) {
                    stack[o] = make_uint2((unsigned)c0 | tag, pr.y);
                    stack[o + 1] = make_uint2((unsigned)c1 | tag, pr.y);
                } else {
                    stack[o] = make_uint2(pr.x, (unsigned)c0);
                    stack[o + 1] = make_uint2(pr.x, (unsigned)c1);
                }
            }
            nt += __popc(mt);
            n += 2 * __popc(mx);
            if (n > NODE_STACK - 64) {  // cannot happen with the throttle above unless trees are > ~60 deep
                err |= GEOM_ERR_STACK;
                return allMask;
            }
            __syncwarp();
        }
        if (nt >= 32 || (n == 0 && nt > 0)) {
            const int p = nt < 32 ? nt : 32;
            bool mine = lane < p;
            unsigned hitBit = 0u;
            uint2 tp = make_uint2(0u, 0u);
            if (mine) tp = triQ[nt - 1 - lane];
            const unsigned slot = tp.x >> SLOT_SHIFT;
            mine = mine && !((hitMask >> slot) & 1u);
            if (mine) {
                ++cnt.tri;
                const float* X = xf[slot];
                float R[9], t[3];
#pragma unroll
                for (int k = 0; k < 9; ++k) R[k] = X[k];
#pragma unroll
                for (int k = 0; k < 3; ++k) t[k] = X[9 + k];
                const float4* rp = reinterpret_cast<const float4*>(m.rTris + (tp.x & NODE_MASK));
                const float4* ep = reinterpret_cast<const float4*>(m.eTris + tp.y);
                float P[3][3], Q[3][3];
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    const float4 rv = __ldg(rp + v), ev = __ldg(ep + v);
                    const float loc[3] = {rv.x, rv.y, rv.z};
                    xformPoint(R, t, loc, P[v]);
                    Q[v][0] = ev.x, Q[v][1] = ev.y, Q[v][2] = ev.z;
                }
                bool ov = true;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float plo = fminf(P[0][c], fminf(P[1][c], P[2][c])), phi = fmaxf(P[0][c], fmaxf(P[1][c], P[2][c]));
                    const float qlo = fminf(Q[0][c], fminf(Q[1][c], Q[2][c])), qhi = fmaxf(Q[0][c], fmaxf(Q[1][c], Q[2][c]));
                    ov = ov && !(plo > qhi || qlo > phi);
                }
                if (ov && triTriIntersect(P, Q)) hitBit = 1u << slot;
            }
            __syncwarp();
            nt -= p;
            hitMask |= __reduce_or_sync(FULL_MASK_, hitBit);
            if (hitMask && (stopAtFirst || hitMask == allMask)) return hitMask;
        }
    }
    return hitMask;
}

__device__ __forceinline__ uint32_t fetchItems(unsigned int* counter, uint32_t count, int lane) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(counter, count);
    return __shfl_sync(FULL_MASK_, i, 0);
}

__device__ __forceinline__ void flushCounters(WarpCounters& c, unsigned long long err, unsigned long long* stats, int lane) {
    unsigned long long bv = c.bv, tri = c.tri;
    for (int o = 16; o > 0; o >>= 1) {
        bv += __shfl_down_sync(FULL_MASK_, bv, o);
        tri += __shfl_down_sync(FULL_MASK_, tri, o);
    }
    if (lane == 0) {
        atomicAdd(stats + 0, (unsigned long long)c.states);
        atomicAdd(stats + 1, bv);
        atomicAdd(stats + 2, tri);
        if (err) atomicOr(stats + 4, err);
    }
}

// state -> transform in shared memory (lane-private work: one lane per slot)
__device__ __forceinline__ void storeTransform(const float* q, float* X) {
    float R[9];
    quatToRot(q, R);
#pragma unroll
    for (int k = 0; k < 9; ++k) X[k] = R[k];
    X[9] = q[4], X[10] = q[5], X[11] = q[6];
}

// ------------------------------------------------------------------ valid(q) for a batch of states
__global__ void __launch_bounds__(MESH_WARPS * 32, MESH_MIN_CTAS) meshValidKernel(MeshDev m, const float* __restrict__ states, uint32_t n,
                                                                   uint8_t* __restrict__ ok, unsigned int* counter,
                                                                   unsigned long long* stats) {
    __shared__ uint2 sStack[MESH_WARPS][NODE_STACK];
    __shared__ uint2 sTri[MESH_WARPS][TRI_QUEUE];
    __shared__ __align__(16) float sXf[MESH_WARPS][MESH_SLOTS][12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpCounters cnt;
    unsigned long long err = 0;
    for (;;) {
        const uint32_t i0 = fetchItems(counter, MESH_SLOTS, lane);
        if (i0 >= n) break;
        const int nb = (int)min((uint32_t)MESH_SLOTS, n - i0);
        __syncwarp();
        if (lane < nb) {
            float q[7];
#pragma unroll
            for (int c = 0; c < 7; ++c) q[c] = __ldg(states + (size_t)(i0 + lane) * 7 + c);
            storeTransform(q, sXf[warp][lane]);
        }
        __syncwarp();
        cnt.states += nb;
        const unsigned hit = warpCollideMulti(m, nb, sXf[warp], sStack[warp], sTri[warp], lane, false, cnt, err);
        if (lane < nb) ok[i0 + lane] = ((hit >> lane) & 1u) ? 0 : 1;
    }
    flushCounters(cnt, err, stats, lane);
}

// ------------------------------------------------------------------ link(a,b): DiscreteMotionValidator
// src/mpt/discrete_motion_validator.hpp:71-130, one warp per edge.  State indices are produced in the
// reference's order (valid(to) first, then the breadth-first bisection of 1..steps-1 through the
// 256-entry ring with its sequential fallback; children are queued before the check, which is
// harmless because a failed check ends the edge) and checked MESH_SLOTS at a time; the edge is
// invalid as soon as any of them collides.  The decision is the reference's AND over the same set of
// states; up to MESH_SLOTS-1 more states than the reference's sequential early exit may be touched.
__global__ void __launch_bounds__(MESH_WARPS * 32, MESH_MIN_CTAS) meshLinkKernel(MeshDev m, DevSpace<float> sp, const float* __restrict__ from,
                                                                  const float* __restrict__ to, uint32_t n, float invStep,
                                                                  uint8_t* __restrict__ ok, unsigned int* counter,
                                                                  unsigned long long* stats) {
    __shared__ uint2 sStack[MESH_WARPS][NODE_STACK];
    __shared__ uint2 sTri[MESH_WARPS][TRI_QUEUE];
    __shared__ uint2 sQueue[MESH_WARPS][DMV_QUEUE];
    __shared__ float sEnds[MESH_WARPS][16];  // from[7], to[7]
    __shared__ uint32_t sIdx[MESH_WARPS][MESH_SLOTS];
    __shared__ __align__(16) float sXf[MESH_WARPS][MESH_SLOTS][12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr uint32_t IDX_TO = 0xFFFFFFFFu;
    WarpCounters cnt;
    unsigned long long err = 0;
    uint2* queue = sQueue[warp];
    float* ends = sEnds[warp];
    for (;;) {
        const uint32_t e = fetchItems(counter, 1u, lane);
        if (e >= n) break;
        __syncwarp();
        if (lane < 7) ends[lane] = __ldg(from + (size_t)e * 7 + lane);
        else if (lane < 14) ends[lane] = __ldg(to + (size_t)e * 7 + (lane - 7));
        __syncwarp();
        // :78  steps = ceil(distance(from,to) * invStepSize)
        const float dist = dev::distance<float>(sp, [&](int c) { return ends[c]; }, [&](int c) { return ends[7 + c]; });
        const float fs = ceilf(dist * invStep);
        bool good = true;
        if (!(fs < 2147483648.0f)) {
            err |= GEOM_ERR_STEPS;
            good = false;
        }
        const uint32_t steps = good ? (uint32_t)fs : 0u;
        const float delta = steps >= 2 ? fp::div_(1.0f, (float)steps) : 0.0f;  // :82
        uint32_t qStart = 0, qEnd = 0, seqCur = 1, seqEnd = 0;
        if (steps >= 2) {  // :79,:99-101
            if (lane == 0) queue[0] = make_uint2(1u, steps - 1u);
            qEnd = 1;
        }
        bool first = true;
        __syncwarp();
        while (good) {
            // next batch of state indices, in the reference's order
            int nb = 0;
            while (nb < MESH_SLOTS) {
                uint32_t i;
                if (first) {  // :75  valid(to)
                    first = false;
                    i = IDX_TO;
                } else if (seqCur <= seqEnd) {  // :119-126 sequential fallback in progress
                    i = seqCur++;
                } else {
                    if (qStart == qEnd) break;
                    const uint2 r = queue[qStart % DMV_QUEUE];
                    ++qStart;
                    __syncwarp();
                    if (r.x == r.y) {  // :104-106
                        i = r.x;
                    } else if (qEnd + 2 < qStart + DMV_QUEUE) {  // :107-114
                        i = (r.x + r.y) / 2;
                        if (r.x < i) {
                            if (lane == 0) queue[qEnd % DMV_QUEUE] = make_uint2(r.x, i - 1);
                            ++qEnd;
                        }
                        if (i < r.y) {
                            if (lane == 0) queue[qEnd % DMV_QUEUE] = make_uint2(i + 1, r.y);
                            ++qEnd;
                        }
                        __syncwarp();
                    } else {
                        seqCur = r.x;
                        seqEnd = r.y;
                        continue;
                    }
                }
                if (lane == 0) sIdx[warp][nb] = i;
                ++nb;
            }
            if (nb == 0) break;
            __syncwarp();
            if (lane < nb) {  // one lane per state: interpolate and build its transform
                const uint32_t i = sIdx[warp][lane];
                float q[7];
                if (i == IDX_TO) {
#pragma unroll
                    for (int c = 0; c < 7; ++c) q[c] = ends[7 + c];
                } else {
                    dev::interpolate<float>(sp, ends, ends + 7, (float)i * delta, q);
                }
                storeTransform(q, sXf[warp][lane]);
            }
            __syncwarp();
            cnt.states += nb;
            if (warpCollideMulti(m, nb, sXf[warp], sStack[warp], sTri[warp], lane, true, cnt, err)) good = false;
        }
        if (lane == 0) ok[e] = good ? 1 : 0;
    }
    flushCounters(cnt, err, stats, lane);
}

// ------------------------------------------------------------------ host: BVH build
namespace {

struct HostTri {
    float v[3][3];
};

struct Builder {
    const std::vector<HostTri>& tris;
    std::vector<BvhNode> nodes;
    int maxDepth = 0;
    explicit Builder(const std::vector<HostTri>& t) : tris(t) {}

    int build(std::vector<int>& ids, int lo, int hi, int depth) {
        if (depth > maxDepth) maxDepth = depth;
        BvhNode n;
        for (int c = 0; c < 3; ++c) n.lo[c] = INFINITY, n.hi[c] = -INFINITY;
        for (int i = lo; i < hi; ++i)
            for (int v = 0; v < 3; ++v)
                for (int c = 0; c < 3; ++c) {
                    n.lo[c] = std::fmin(n.lo[c], tris[ids[i]].v[v][c]);
                    n.hi[c] = std::fmax(n.hi[c], tris[ids[i]].v[v][c]);
                }
        n.left = n.right = -1;
        const int me = (int)nodes.size();
        nodes.push_back(n);
        if (hi - lo == 1) {
            nodes[me].left = -1 - ids[lo];
            return me;
        }
        int axis = 0;
        float ext = -1;
        for (int c = 0; c < 3; ++c)
            if (n.hi[c] - n.lo[c] > ext) ext = n.hi[c] - n.lo[c], axis = c;
        const int mid = (lo + hi) / 2;
        std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int x, int y) {
            const float cx = tris[x].v[0][axis] + tris[x].v[1][axis] + tris[x].v[2][axis];
            const float cy = tris[y].v[0][axis] + tris[y].v[1][axis] + tris[y].v[2][axis];
            return cx < cy || (cx == cy && x < y);
        });
        const int l = build(ids, lo, mid, depth + 1);
        const int r = build(ids, mid, hi, depth + 1);
        nodes[me].left = l;
        nodes[me].right = r;
        return me;
    }
};

int uploadMesh(mptg_ctx* ctx, const float* tris9, uint32_t n, void** nodesDev, void** trisDev, int* depth) {
    std::vector<HostTri> tris(n);
    std::vector<TriPad> pad(n ? n : 1);
    for (uint32_t i = 0; i < n; ++i)
        for (int v = 0; v < 3; ++v) {
            for (int c = 0; c < 3; ++c) tris[i].v[v][c] = pad[i].v[v][c] = tris9[(size_t)i * 9 + v * 3 + c];
            pad[i].v[v][3] = 0.0f;
        }
    Builder b(tris);
    if (n) {
        std::vector<int> ids(n);
        for (uint32_t i = 0; i < n; ++i) ids[i] = (int)i;
        b.nodes.reserve(2 * (size_t)n);
        b.build(ids, 0, (int)n, 0);
    } else {
        b.nodes.push_back(BvhNode{});
    }
    *depth = b.maxDepth;
    MPTG_CUDA(ctx, cudaMalloc(nodesDev, b.nodes.size() * sizeof(BvhNode)));
    if (int rc = uploadSync(ctx, *nodesDev, b.nodes.data(), b.nodes.size() * sizeof(BvhNode))) return rc;
    MPTG_CUDA(ctx, cudaMalloc(trisDev, pad.size() * sizeof(TriPad)));
    if (int rc = uploadSync(ctx, *trisDev, pad.data(), pad.size() * sizeof(TriPad))) return rc;
    return MPTG_OK;
}

int meshGrid(mptg_ctx* ctx, uint32_t n) {
    const uint32_t want = (n + MESH_WARPS - 1) / MESH_WARPS;
    const uint32_t cap = (uint32_t)ctx->smCount * MESH_MIN_CTAS;  // persistent grid: every resident slot gets a CTA
    return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace

int meshCreate(mptg_ctx* ctx, int /*scalar*/, uint32_t nr, const float* robotTris, uint32_t ne, const float* envTris,
               MeshData** out) {
    auto* m = new MeshData();
    int rc = uploadMesh(ctx, robotTris, nr, &m->mem[0], &m->mem[1], &m->depthR);
    if (!rc) rc = uploadMesh(ctx, envTris, ne, &m->mem[2], &m->mem[3], &m->depthE);
    if (!rc) {
        cudaError_t e = cudaMalloc(&m->workCounter, sizeof(unsigned int));
        if (e != cudaSuccess) rc = fail(ctx, MPTG_ERR_CUDA, "meshCreate: %s", cudaGetErrorString(e));
    }
    if (!rc && m->depthR + m->depthE > 56)
        rc = fail(ctx, MPTG_ERR_CAPACITY, "meshCreate: BVH depth %d + %d exceeds the traversal stack budget", m->depthR, m->depthE);
    if (rc) {
        meshDestroy(m);
        return rc;
    }
    m->dev.rNodes = (const BvhNode*)m->mem[0];
    m->dev.rTris = (const TriPad*)m->mem[1];
    m->dev.eNodes = (const BvhNode*)m->mem[2];
    m->dev.eTris = (const TriPad*)m->mem[3];
    m->dev.nR = nr;
    m->dev.nE = ne;
    *out = m;
    return MPTG_OK;
}

void meshDestroy(MeshData* m) {
    if (!m) return;
    for (void* p : m->mem) cudaFree(p);
    cudaFree(m->workCounter);
    delete m;
}

int meshValidDev(mptg_geom* g, const void* states, uint32_t n, uint8_t* ok) {
    mptg_ctx* ctx = g->ctx;
    MPTG_CUDA(ctx, cudaMemsetAsync(g->mesh->workCounter, 0, sizeof(unsigned int), ctx->stream));
    meshValidKernel<<<meshGrid(ctx, n), MESH_WARPS * 32, 0, ctx->stream>>>(g->mesh->dev, (const float*)states, n, ok,
                                                                          g->mesh->workCounter, g->devStats);
    MPTG_LAUNCHED(ctx);
    return MPTG_OK;
}

int meshLinkDev(mptg_geom* g, const mptg_space_desc* space, const void* from, const void* to, uint32_t n, double step,
                uint8_t* ok) {
    mptg_ctx* ctx = g->ctx;
    MPTG_CUDA(ctx, cudaMemsetAsync(g->mesh->workCounter, 0, sizeof(unsigned int), ctx->stream));
    const float invStep = 1.0f / (float)step;  // discrete_motion_validator.hpp:64
    meshLinkKernel<<<meshGrid(ctx, n), MESH_WARPS * 32, 0, ctx->stream>>>(g->mesh->dev, makeDevSpace<float>(*space),
                                                                         (const float*)from, (const float*)to, n, invStep, ok,
                                                                         g->mesh->workCounter, g->devStats);
    MPTG_LAUNCHED(ctx);
    return MPTG_OK;
}

}  // namespace mptg
