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
// Execution: persistent warps over a flat list of states (see "flat work list" below): every warp keeps
// up to 32 states of any items in flight and its 32 lanes test node pairs from one mixed frontier held
// in shared memory; leaf-leaf survivors go to a triangle-pair queue that is drained 32 at a time so the
// SAT runs without divergence.  States arrive as float or double (the edge discretisation runs in the
// state's scalar type); the collision test itself is always float, on the pose rounded to float.
// #define MESH_DEBUG_STATS 1
#include <cub/device/device_scan.cuh>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#include "geom.cuh"

namespace mptg {

struct __align__(16) BvhNode {
    float lo[3];  // host build: box minimum; device image: box CENTRE
    int left;     // >= 0: internal (children left,right); < 0: leaf holding triangle (-1 - left)
    float hi[3];  // host build: box maximum; device image: HALF EXTENT (rounded outwards)
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
    uint32_t stageR, stageE;  // nodes [0, stage) of each tree are copied into shared memory by every CTA (breadth-first prefix)
    uint32_t rootX, rootY;    // the first entry of every state (meshRootEntry): the children of one root against the other root ...
    uint32_t rootIsTri;       // ... or, when both hierarchies are a single triangle, that triangle pair
};

}  // namespace mptg

namespace mptg {
struct DonBlock;
}
struct MeshData {
    mptg::MeshDev dev{};
    void* mem[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned long long* workCounter = nullptr;  // [0] pass 1 / states, [1] pass 2; followed by the donation control words
    mptg::DonBlock* poolBlocks = nullptr;
    int depthR = 0, depthE = 0;
    float band = 0.0f;  // contact band: 1e-6 x the diagonal of the environment's bounding box (absolute length)
    // per-edge work buffers of link batches (grow-only)
    uint32_t* steps = nullptr;
    unsigned long long* counts = nullptr;
    unsigned long long* offs = nullptr;
    void* scanTemp = nullptr;
    size_t scanBytes = 0;
    uint32_t workEdges = 0;
};

namespace mptg {

// One large CTA per SM: its warps share ONE copy of the hierarchies' upper levels in shared memory (north_star:
// "obstacle BVHs staged through shared memory/TMA"), brought in by two bulk copies (cp.async.bulk + mbarrier) when the
// persistent CTA starts.  r1 ran 6 CTAs of 4 warps per SM and fetched every node pair through L1/L2 (ncu: 2.06 warps
// per issue waiting on the long scoreboard, 19 of 64 warps active).
#ifndef MESH_WARPS_PER_CTA
#define MESH_WARPS_PER_CTA 20
#endif
constexpr int MESH_WARPS = MESH_WARPS_PER_CTA;
#ifndef MESH_CTAS
#define MESH_CTAS 1
#endif
constexpr int MESH_MIN_CTAS = MESH_CTAS;  // resident CTAs per SM (caps registers per thread)
constexpr size_t MESH_SMEM_LIMIT = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100
constexpr int NODE_STACK = 512;
constexpr size_t MESH_WARP_SMEM = 512 * 8 + 96 * 8 + 32 * 12 * 4 + 32 * 4;  // stack + triangle queue + transforms + items
constexpr int STACK_SOFT = NODE_STACK - 64 - 60;  // see the pop rule in meshFlatKernel
#ifndef MESH_TRI_DRAIN_AT
#define MESH_TRI_DRAIN_AT 32
#endif
constexpr int TRI_DRAIN_AT = MESH_TRI_DRAIN_AT;  // triangle pairs pending before the SAT runs (<= 32: at most 31 + 64 are ever queued)
constexpr int TRI_QUEUE = 96;  // fewer than 32 pending when a round starts, at most two more per lane in it

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

// The same test with the contact band made visible (near_contact_out of the C ABI).  Per axis g = signed gap of the two
// projections (> 0: separated), compared with +-tol |axis|:
//   some axis with g > tol |axis|        -> 0: clearly apart
//   otherwise bit 0: no axis separates (contact), bit 1: some axis has |g| within tol |axis| -- the largest normalised
//   gap, which is what decides, lies in (-tol, tol]: a perturbation of that size can flip the decision.
__device__ __noinline__ int triTriNear(const float P[3][3], const float Qt[3][3], float tol) {
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
    bool hit = true, within = false, apart = false;
    auto axis = [&](const float ax[3]) {
        const float P0 = dot3(ax, p[0]), P1 = dot3(ax, p[1]), P2 = dot3(ax, p[2]);
        const float Q0 = dot3(ax, q[0]), Q1 = dot3(ax, q[1]), Q2 = dot3(ax, q[2]);
        const float mx1 = fmaxf(P0, fmaxf(P1, P2)), mn1 = fminf(P0, fminf(P1, P2));
        const float mx2 = fmaxf(Q0, fmaxf(Q1, Q2)), mn2 = fminf(Q0, fminf(Q1, Q2));
        if (mn1 > mx2 || mn2 > mx1) hit = false;  // the decision itself: exactly project6
        const float g = fmaxf(mn1 - mx2, mn2 - mx1);
        const float tl = tol * sqrtf(dot3(ax, ax));
        if (g > tl) apart = true;
        else if (g > -tl) within = true;
    };
    float n1[3], m1[3], ax[3];
    cross3(e[0], e[1], n1);
    axis(n1);
    cross3(f[0], f[1], m1);
    axis(m1);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            cross3(e[i], f[j], ax);
            axis(ax);
        }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        cross3(e[i], n1, ax);
        axis(ax);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        cross3(f[i], m1, ax);
        axis(ax);
    }
    if (apart) return 0;
    return (hit ? 1 : 0) | (within ? 2 : 0);
}

__device__ __forceinline__ BvhNode loadNode(const BvhNode* nodes, int i) {
    const float4* p = reinterpret_cast<const float4*>(nodes + i);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    BvhNode n;
    n.lo[0] = a.x, n.lo[1] = a.y, n.lo[2] = a.z, n.left = __float_as_int(a.w);
    n.hi[0] = b.x, n.hi[1] = b.y, n.hi[2] = b.z, n.right = __float_as_int(b.w);
    return n;
}

// node i of a tree: from the staged prefix in shared memory, or from global memory below it
__device__ __forceinline__ BvhNode loadNodeStaged(const BvhNode* __restrict__ sm, uint32_t staged, const BvhNode* __restrict__ gl, int i) {
    if ((uint32_t)i < staged) {
        const float4* p = reinterpret_cast<const float4*>(sm + i);
        const float4 a = p[0], b = p[1];
        BvhNode n;
        n.lo[0] = a.x, n.lo[1] = a.y, n.lo[2] = a.z, n.left = __float_as_int(a.w);
        n.hi[0] = b.x, n.hi[1] = b.y, n.hi[2] = b.z, n.right = __float_as_int(b.w);
        return n;
    }
    return loadNode(gl, i);
}

// ---- bulk copy global -> shared, completion on an mbarrier (TMA's 1-D form; SASS: UBLKCP)
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkCopyG2S(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dst)), "l"(src),
                 "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smemAddr(bar)),
        "r"(parity)
        : "memory");
}

// ------------------------------------------------------------------ flat work list + mixed frontier
// Work is a flat list of rigid-body states, each belonging to an ITEM (a state of a valid() batch, or an
// edge of a link() batch).  A warp keeps up to MESH_SLOTS states in flight: their transforms sit in
// shared memory (R row-major, t), and every stack / queue entry carries its slot in the top 5 bits of the
// robot index, so the 32 lanes draw BV pair tests from ONE mixed frontier.  Whenever the frontier gets
// narrow (<= REFILL_AT pairs) the warp pulls more work ids from a global counter into the slots that
// have nothing pending, so lanes stay busy across states and across edges.  The first collision found
// for an item clears ok[item]; its remaining pairs are dropped, and states of an edge already known to
// be invalid are skipped when they are pulled.
//
// Edges (DiscreteMotionValidator::operator(), src/mpt/discrete_motion_validator.hpp:71-130): the
// reference's decision is  valid(to) && AND_{i=1..steps-1} valid(interpolate(from, to, i * (1/steps)))
// with steps = ceil(distance * (1/stepSize)) (:64,:78,:82); its bisection queue (:99-126) only fixes the
// ORDER in which that set is visited, i.e. how early an invalid edge is abandoned.  The decision is
// order-independent, so the same set is enumerated here coarse-to-fine over ALL edges in one list:
//   first 8 ids per edge -- `to`, then the multiples of P/8 inside 1..steps-1 in bisection order
//         (P = max(8, steps rounded up to a power of two)),
//   then, for the edges with steps > 8, their other interior indices (located through a prefix sum of
//         the per-edge counts); by the time these are pulled the coarse states of the edge have been
//         checked, and an edge already found invalid is skipped.
constexpr int MESH_SLOTS = 32;
constexpr unsigned SLOT_SHIFT = 27;
constexpr unsigned SIDE_ENV = 1u << 26;  // stack entry: the pair of siblings to test is on the environment's side
constexpr unsigned NODE_MASK = SIDE_ENV - 1u;
constexpr int XF_STRIDE = 12;   // 48-byte rows: three 128-bit loads per box test, conflict-free for 8 consecutive slots
#ifndef MESH_REFILL_AT
#define MESH_REFILL_AT 24
#endif
#ifndef MESH_REFILL_MAX
#define MESH_REFILL_MAX 24
#endif
constexpr int REFILL_AT = MESH_REFILL_AT;    // refill when at most this many node pairs are pending
constexpr int REFILL_MAX = MESH_REFILL_MAX;  // states started per refill
constexpr int COARSE_IDS = 8;   // work ids per edge in pass 1

enum { WORK_STATES = 0, WORK_EDGES = 1 };

template <typename S>
struct MeshWork {
    const S* a;     // states (WORK_STATES) / edge starts
    const S* b;     // edge ends
    uint32_t n;     // states / edges
    const uint32_t* steps;           // per edge
    const unsigned long long* offs;  // exclusive prefix sums of the per-edge counts of non-coarse states, n+1 entries
    uint8_t* ok;                     // per item, preset to 1
    uint8_t* nearOut;                // per item, preset to 0 (NEAR kernels only): some checked state came within `tol` of contact
    float tol;                       // contact band, absolute
    DevSpace<S> sp;
};

// ---- work donation between warps.  A few states (deep contact, long grazing passes) need thousands of
// pair tests; left to the warp that pulled them they become the tail of the launch.  Once the work list
// is exhausted, warps with nothing left register as idle, and a warp whose stack is deep hands the
// OLDEST half of it (the largest subtrees) to an idle one through a ring of blocks in global memory.  A
// block carries the donor's whole slot table (transforms + items), so the receiver -- which is empty --
// adopts it as is.  Both sides may then work on the same state; a hit by either clears ok[item], which
// the other notices at its next periodic look at ok[] (also how warps learn that ANOTHER warp has
// already invalidated an edge).
// tuned on the C5 edge wave (B200): profiles/r1_mesh_donation_tuning.txt, and again for the two-test round
// (profiles/r2_mesh_two_test_sweep.txt: 20 warps per CTA, refill at 24, a look at the ring every second round)
#ifndef MESH_DON_MAX
#define MESH_DON_MAX 128
#define MESH_DON_MIN 64
#define MESH_DON_KEEP 32
#define MESH_DON_EVERY 1u
#define MESH_MAX_POLLERS 256
#endif
constexpr int DON_MAX = MESH_DON_MAX;        // pairs per block
constexpr int DON_MIN_STACK = MESH_DON_MIN;  // only stacks at least this deep are split
constexpr int DON_KEEP = MESH_DON_KEEP;       // pairs the donor keeps (the newest: its next round)
constexpr unsigned POOL_BLOCKS = 8192;  // ring size: more than the waiters plus one block per warp that can donate at once
enum { CTL_HEAD = 0, CTL_TAIL = 1, CTL_IDLE = 2, CTL_STARTED = 3, CTL_WORDS = 32 };  // one 128-byte line per pass
constexpr unsigned MAX_POLLERS = MESH_MAX_POLLERS;  // idle warps that queue for donated work; the others leave

struct DonBlock {
    float xf[MESH_SLOTS * 12];
    uint32_t items[MESH_SLOTS];
    uint2 pairs[DON_MAX];
    uint32_t count, dead, pad0, pad1;
};

struct MeshPool {
    DonBlock* blocks;
    unsigned int* ctl;    // CTL_WORDS
    unsigned int* ready;  // POOL_BLOCKS sequence numbers (0 = empty)
};

// control block: the two work counters (u64) on lines of their own, then per pass CTL_WORDS control words and
// POOL_BLOCKS ready flags
constexpr size_t CTL_COUNTER_STRIDE = 16;  // u64 units: 128 bytes between the counters
constexpr size_t CTL_PASS_WORDS = CTL_WORDS + POOL_BLOCKS;
constexpr size_t CTL_BYTES = 2 * CTL_COUNTER_STRIDE * sizeof(unsigned long long) + 2 * CTL_PASS_WORDS * sizeof(unsigned int);

__device__ __forceinline__ unsigned int volLoad(const unsigned int* p) { return *(const volatile unsigned int*)p; }

__device__ __forceinline__ uint32_t pow2AtLeast8(uint32_t steps) {  // steps <= 2^31
    return steps <= 8u ? 8u : (1u << (32 - __clz(steps - 1u)));
}

// state -> transform in shared memory (lane-private work: one lane per slot)
__device__ __forceinline__ void storeTransform(const float* q, float* X) {
    float R[9];
    quatToRot(q, R);
#pragma unroll
    for (int k = 0; k < 9; ++k) X[k] = R[k];
    X[9] = q[4], X[10] = q[5], X[11] = q[6];
}

// Decode one work id into (item, transform).  false: nothing to check (a hole of the enumeration, or an
// edge already known to be invalid).
template <typename S, int MODE>
__device__ __forceinline__ bool decodeWork(const MeshWork<S>& w, unsigned long long id, uint32_t& item, float* X) {
    S q[7];
    if (MODE == WORK_STATES) {
        item = (uint32_t)id;
#pragma unroll
        for (int c = 0; c < 7; ++c) q[c] = __ldg(w.a + (size_t)item * 7 + c);
    } else {
        uint32_t e, i = 0;
        const unsigned long long nCoarse = (unsigned long long)w.n * COARSE_IDS;
        if (id < nCoarse) {
            e = (uint32_t)(id / COARSE_IDS);
            const uint32_t j = (uint32_t)(id % COARSE_IDS);
            const uint32_t steps = __ldg(w.steps + e);
            if (j != 0) {
                if (steps < 2u) return false;
                const int l = 31 - __clz(j);  // bisection level 0..2, k-th node of the level
                const uint32_t k = j - (1u << l);
                i = (2u * k + 1u) * (pow2AtLeast8(steps) >> (l + 1));
                if (i > steps - 1u) return false;
            }
        } else {
            const unsigned long long r = id - nCoarse;
            uint32_t lo = 0, hi = w.n;  // largest e with offs[e] <= r (offs[n] = total > r)
            while (hi - lo > 1u) {
                const uint32_t mid = lo + (hi - lo) / 2u;
                if (__ldg(w.offs + mid) <= r) lo = mid;
                else hi = mid;
            }
            e = lo;
            i = (uint32_t)(r - __ldg(w.offs + e)) + 1u;
            if (i % (pow2AtLeast8(__ldg(w.steps + e)) >> 3) == 0u) return false;  // one of the coarse states
        }
        if (__ldcg(w.ok + e) == 0) return false;
        item = e;
        if (i == 0) {  // :75 valid(to)
#pragma unroll
            for (int c = 0; c < 7; ++c) q[c] = __ldg(w.b + (size_t)e * 7 + c);
        } else {  // :82,:110 interpolate(from, to, i * delta), delta = 1 / steps
            S a[7], b[7];
#pragma unroll
            for (int c = 0; c < 7; ++c) a[c] = __ldg(w.a + (size_t)e * 7 + c), b[c] = __ldg(w.b + (size_t)e * 7 + c);
            const S delta = fp::div_(S(1), (S)__ldg(w.steps + e));
            dev::interpolate<S>(w.sp, a, b, (S)i * delta, q);
        }
    }
    float qf[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) qf[c] = (float)q[c];
    storeTransform(qf, X);
    return true;
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

#ifdef MESH_DEBUG_STATS
__device__ unsigned long long gDbgT[3][8192][8];  // per mode, per warp: start, end, rounds, bv, tri, states, tri rounds
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

template <typename S, int MODE, bool NEAR>
__global__ void __launch_bounds__(MESH_WARPS * 32, MESH_MIN_CTAS) meshFlatKernel(MeshDev m, MeshWork<S> w, unsigned long long* counter,
                                                                                MeshPool pool, unsigned long long* stats) {
    // dynamic shared memory: [staged robot nodes][staged env nodes][per warp: stack, triangle queue, transforms, items]
    extern __shared__ __align__(128) unsigned char meshSmem[];
    __shared__ uint64_t stageBar;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned ltMask = (1u << lane) - 1u;
    BvhNode* sR = reinterpret_cast<BvhNode*>(meshSmem);
    BvhNode* sE = sR + m.stageR;
    unsigned char* wb = reinterpret_cast<unsigned char*>(sE + m.stageE) + (size_t)warp * MESH_WARP_SMEM;
    uint2* stack = reinterpret_cast<uint2*>(wb);
    uint2* triQ = stack + NODE_STACK;
    float(*xf)[XF_STRIDE] = reinterpret_cast<float(*)[XF_STRIDE]>(triQ + TRI_QUEUE);
    uint32_t* items = reinterpret_cast<uint32_t*>(xf + MESH_SLOTS);
    {  // stage the upper levels of both hierarchies: one thread issues two bulk copies, everybody waits on the barrier
        const unsigned bytesR = m.stageR * (unsigned)sizeof(BvhNode), bytesE = m.stageE * (unsigned)sizeof(BvhNode);
        if (threadIdx.x == 0) mbarInit(&stageBar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbarExpectTx(&stageBar, bytesR + bytesE);
            if (bytesR) bulkCopyG2S(sR, m.rNodes, bytesR, &stageBar);
            if (bytesE) bulkCopyG2S(sE, m.eNodes, bytesE, &stageBar);
        }
        mbarWait(&stageBar, 0);
    }
    const unsigned long long total =
        MODE == WORK_STATES ? (unsigned long long)w.n : (unsigned long long)w.n * COARSE_IDS + __ldg(w.offs + w.n);
    const unsigned long long nWarps = (unsigned long long)gridDim.x * MESH_WARPS;
    WarpCounters cnt;
    unsigned long long err = 0;
    unsigned dead = 0u;  // slots whose item is already decided: their pending pairs are dropped
    int n = 0, nt = 0;
    bool exhausted = total == 0;
    unsigned roundNo = 0;
    unsigned fresh = 0xffffffffu;  // slots (re)filled since ok[] was last requested
    unsigned ctlSeen = 0u;
    uint8_t okSeen = 1;
    items[lane] = 0;
    if (lane == 0) atomicAdd(pool.ctl + CTL_STARTED, 1u);
    __syncwarp();
#ifdef MESH_DEBUG_STATS
    unsigned long long dbgRounds = 0, dbgThrottled = 0, dbgTriRounds = 0, dbgAdopted = 0, dbgDonated = 0;
    int dbgMaxN = 0;
    const unsigned long long dbgT0 = gtimer();
    unsigned long long dbgIdleT = 0, dbgSpin = 0, dbgPoll = 0;
#endif
    for (;;) {
        if (!exhausted && n <= REFILL_AT) {
            // slots with nothing pending in the stack or the triangle queue are free
            unsigned bits = 0u;
            if (lane < n) bits |= 1u << (stack[lane].x >> SLOT_SHIFT);
            if (lane < nt) bits |= 1u << (triQ[lane].x >> SLOT_SHIFT);
            if (lane + 32 < nt) bits |= 1u << (triQ[lane + 32].x >> SLOT_SHIFT);
            const unsigned freeSlots = ~__reduce_or_sync(FULL_MASK_, bits);
            // near the end of the list take smaller bites so that the warps finish together
            unsigned long long base = 0;
            int cap = 0;
            if (lane == 0) {
                const unsigned long long cur = *(volatile unsigned long long*)counter;
                const unsigned long long share = cur < total ? (total - cur) / nWarps + 1ull : 1ull;
                cap = (int)(share < 4ull ? 4ull : (share > (unsigned long long)REFILL_MAX ? (unsigned long long)REFILL_MAX : share));
                const int avail = __popc(freeSlots);
                cap = cap < avail ? cap : avail;
                if (cap > 0) base = atomicAdd(counter, (unsigned long long)cap);
            }
            cap = __shfl_sync(FULL_MASK_, cap, 0);
            base = __shfl_sync(FULL_MASK_, base, 0);
            if (cap > 0) {
                const int rank = __popc(freeSlots & ltMask);
                const bool take = ((freeSlots >> lane) & 1u) && rank < cap && base + (unsigned long long)rank < total;
                bool push = false;
                if (take) {
                    uint32_t item;
                    push = decodeWork<S, MODE>(w, base + (unsigned long long)rank, item, xf[lane]);
                    if (push) items[lane] = item;
                }
                const unsigned mp = __ballot_sync(FULL_MASK_, push);
                dead &= ~mp;
                fresh |= mp;
                // (the roots are not tested against each other: bounds only prune)
                if (m.rootIsTri) {
                    if (push) triQ[nt + __popc(mp & ltMask)] = make_uint2(m.rootX | ((unsigned)lane << SLOT_SHIFT), m.rootY);
                    nt += __popc(mp);
                } else {
                    if (push) stack[n + __popc(mp & ltMask)] = make_uint2(m.rootX | ((unsigned)lane << SLOT_SHIFT), m.rootY);
                    n += __popc(mp);
                }
                cnt.states += __popc(mp);
                if (base + (unsigned long long)cap >= total) exhausted = true;
                __syncwarp();
            }
        }
        if (n == 0 && nt == 0) {
            if (!exhausted) continue;
            // Nothing left here: queue for donated work (ticket = position in the ring), unless enough warps
            // are queueing already.  CTL_IDLE counts the idle warps MINUS the blocks handed out and not yet
            // picked up (the donor subtracts one per block), so IDLE >= STARTED means: every warp is idle and
            // no block is pending -- nothing can arrive any more.  Each waiter spins on its own ready flag.
            unsigned got = 0xffffffffu;
            bool watchdog = false;
#ifdef MESH_DEBUG_STATS
            if (dbgIdleT == 0) dbgIdleT = gtimer();
#endif
            if (lane == 0) {
                __threadfence();
                atomicAdd(pool.ctl + CTL_IDLE, 1u);
                const int waiting = (int)(volLoad(pool.ctl + CTL_TAIL) - volLoad(pool.ctl + CTL_HEAD));  // < 0: blocks pending
                if (waiting < (int)MAX_POLLERS) {
                    const unsigned my = atomicAdd(pool.ctl + CTL_TAIL, 1u);
                    const unsigned* flag = pool.ready + my % POOL_BLOCKS;
                    for (unsigned it = 0;; ++it) {
                        if (volLoad(flag) == my + 1u) {
                            got = my;
                            break;
                        }
                        if ((it & 3u) == 3u && (int)volLoad(pool.ctl + CTL_IDLE) >= (int)volLoad(pool.ctl + CTL_STARTED)) break;  // signed: IDLE dips below zero while blocks outnumber idle warps
                        if (it > (1u << 22)) {  // seconds of waiting: never hang the device, report instead
                            watchdog = true;
                            break;
                        }
                        __nanosleep(500);
#ifdef MESH_DEBUG_STATS
                        ++dbgPoll;
#endif
                    }
                }
            }
            if (watchdog) err |= GEOM_ERR_SCHED;
            got = __shfl_sync(FULL_MASK_, got, 0);
            if (got == 0xffffffffu) break;
            const unsigned bi = got % POOL_BLOCKS;
            __syncwarp();
            __threadfence();
            const DonBlock* blk = pool.blocks + bi;
            for (int i = lane; i < MESH_SLOTS * 12; i += 32) xf[i / 12][i % 12] = __ldcg(blk->xf + i);
            items[lane] = __ldcg(blk->items + lane);
            n = (int)__ldcg(&blk->count);
            dead = __ldcg(&blk->dead);
            for (int i = lane; i < n; i += 32) stack[i] = __ldcg(blk->pairs + i);
            fresh = 0xffffffffu;
            __syncwarp();
#ifdef MESH_DEBUG_STATS
            ++dbgAdopted;
#endif
            continue;
        }
        if ((++roundNo & MESH_DON_EVERY) == 0u) {
            // Every (MESH_DON_EVERY + 1)-th round: a look at ok[] (items invalidated by another warp are finished here too) and at
            // the donation ring.  The values used are the ones requested at the PREVIOUS look, so the loads
            // never stall the traversal; slots refilled since then are skipped.
            dead |= __ballot_sync(FULL_MASK_, okSeen == 0) & ~fresh;
            fresh = 0u;
            okSeen = __ldcg(w.ok + items[lane]);
            const int waiting = (int)(__shfl_sync(FULL_MASK_, ctlSeen, 0) - __shfl_sync(FULL_MASK_, ctlSeen, 1));  // TAIL - HEAD
            if (lane < 2 && n >= DON_MIN_STACK / 2 + 8) ctlSeen = volLoad(pool.ctl + (lane == 0 ? CTL_TAIL : CTL_HEAD));
            bool donate = n >= DON_MIN_STACK && waiting > 0;
            if (donate) {  // confirm on fresh values (rare path); a few warps may overshoot, those blocks wait for the next idle warp
                unsigned v = 0u;
                if (lane < 2) v = volLoad(pool.ctl + (lane == 0 ? CTL_TAIL : CTL_HEAD));
                donate = (int)(__shfl_sync(FULL_MASK_, v, 0) - __shfl_sync(FULL_MASK_, v, 1)) > 0;
            }
            {
                if (donate) {
                    const int d = n - DON_KEEP < DON_MAX ? n - DON_KEEP : DON_MAX;
                    unsigned idx = 0;
                    if (lane == 0) {
                        idx = atomicAdd(pool.ctl + CTL_HEAD, 1u);
                        atomicSub(pool.ctl + CTL_IDLE, 1u);  // on behalf of the warp that will pick the block up
                    }
                    idx = __shfl_sync(FULL_MASK_, idx, 0);
                    DonBlock* blk = pool.blocks + idx % POOL_BLOCKS;
                    for (int i = lane; i < d; i += 32) blk->pairs[i] = stack[i];
                    for (int i = lane; i < MESH_SLOTS * 12; i += 32) blk->xf[i] = xf[i / 12][i % 12];
                    blk->items[lane] = items[lane];
                    if (lane == 0) blk->count = (uint32_t)d, blk->dead = dead;
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) *(volatile unsigned int*)(pool.ready + idx % POOL_BLOCKS) = idx + 1u;
                    for (int base = 0; base < n - d; base += 32) {  // close the gap
                        uint2 e = make_uint2(0u, 0u);
                        if (base + lane < n - d) e = stack[d + base + lane];
                        __syncwarp();
                        if (base + lane < n - d) stack[base + lane] = e;
                        __syncwarp();
                    }
                    n -= d;
#ifdef MESH_DEBUG_STATS
                    ++dbgDonated;
#endif
                }
            }
        }
        if (n > 0) {
            // Pop up to 32 entries, fewer as the stack nears STACK_SOFT (each pop pushes at most two); from
            // STACK_SOFT on it is one entry per round, a depth-first descent that can add at most
            // depthR + depthE <= 56 more entries (checked at creation), which stays below the hard limit.
            // An entry is a pair of nodes KNOWN to overlap, one of them already replaced by the first of its two children
            // (siblings are neighbours in the node array): the lane tests both children against the other node -- two
            // independent box tests in flight -- and pushes only what survives.  (r1/r2a pushed both children untested and
            // popped them again, one test per lane and round: twice the stack traffic, rounds and votes per test.)
            int p = STACK_SOFT - n;
            p = p < 1 ? 1 : (p > 32 ? 32 : p);
            p = p < n ? p : n;
#ifdef MESH_DEBUG_STATS
            ++dbgRounds;
            if (p < 32 && p < n) ++dbgThrottled;
            if (n > dbgMaxN) dbgMaxN = n;
#endif
            bool mine = lane < p;
            uint2 pr = make_uint2(0u, 0u);
            if (mine) pr = stack[n - 1 - lane];
            __syncwarp();
            n -= p;
            const unsigned slot = pr.x >> SLOT_SHIFT;
            const unsigned tag = slot << SLOT_SHIFT;
            mine = mine && !((dead >> slot) & 1u);
            int kind0 = 0, kind1 = 0;  // per test: 1 triangle pair, 2 stack entry
            unsigned ex0 = 0u, ey0 = 0u, ex1 = 0u, ey1 = 0u;
            if (mine) {
                const bool sideE = (pr.x & SIDE_ENV) != 0u;
                const int rid = (int)(pr.x & NODE_MASK), eid = (int)pr.y;
                const BvhNode r0 = loadNodeStaged(sR, m.stageR, m.rNodes, rid);
                const BvhNode e0 = loadNodeStaged(sE, m.stageE, m.eNodes, eid);
                const BvhNode sib = loadNodeStaged(sideE ? sE : sR, sideE ? m.stageE : m.stageR, sideE ? m.eNodes : m.rNodes, (sideE ? eid : rid) + 1);
                const float* X = xf[slot];
                cnt.bv += 2;
                const float4 x0 = *reinterpret_cast<const float4*>(X), x1 = *reinterpret_cast<const float4*>(X + 4),
                             x2 = *reinterpret_cast<const float4*>(X + 8);
                const float R[9] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x};
                const float T[3] = {x2.y, x2.z, x2.w};
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    // test 0: (r0, e0); test 1: the sibling on the entry's side against the other node
                    const bool second = t == 1;
                    const bool robotSib = second && !sideE, envSib = second && sideE;
                    const float cx = robotSib ? sib.lo[0] : r0.lo[0], cy = robotSib ? sib.lo[1] : r0.lo[1], cz = robotSib ? sib.lo[2] : r0.lo[2];
                    const float hx = robotSib ? sib.hi[0] : r0.hi[0], hy = robotSib ? sib.hi[1] : r0.hi[1], hz = robotSib ? sib.hi[2] : r0.hi[2];
                    const int aLeft = robotSib ? sib.left : r0.left, bLeft = envSib ? sib.left : e0.left;
                    const unsigned aId = (unsigned)(robotSib ? rid + 1 : rid), bId = (unsigned)(envSib ? eid + 1 : eid);
                    // robot box (centre c, half extents h in the robot frame) against the env box, separating
                    // axes = the three world axes and the three robot-frame axes; all bounds padded
                    float d[3], hb[3], hw[3];
                    float mag = 0.0f;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const float cw = __fmaf_rn(R[3 * r + 2], cz, __fmaf_rn(R[3 * r + 1], cy, __fmaf_rn(R[3 * r], cx, T[r])));
                        hw[r] = __fmaf_rn(fabsf(R[3 * r + 2]), hz, __fmaf_rn(fabsf(R[3 * r + 1]), hy, fabsf(R[3 * r]) * hx));
                        const float cb = envSib ? sib.lo[r] : e0.lo[r];
                        hb[r] = envSib ? sib.hi[r] : e0.hi[r];
                        d[r] = cb - cw;
                        mag += (fabsf(cw) + hw[r]) + (fabsf(cb) + hb[r]);
                    }
                    const float pad = 64.0f * 1.1920928955078125e-07f * mag + (NEAR ? w.tol : 0.0f);
                    bool ov = true;
#pragma unroll
                    for (int r = 0; r < 3; ++r) ov = ov && !(fabsf(d[r]) > hw[r] + hb[r] + pad);
                    if (ov) {
                        const float ha[3] = {hx, hy, hz};
#pragma unroll
                        for (int j = 0; j < 3; ++j) {  // robot-frame axis j = column j of R
                            const float dl = __fmaf_rn(R[6 + j], d[2], __fmaf_rn(R[3 + j], d[1], R[j] * d[0]));
                            const float el = __fmaf_rn(fabsf(R[6 + j]), hb[2], __fmaf_rn(fabsf(R[3 + j]), hb[1], fabsf(R[j]) * hb[0]));
                            ov = ov && !(fabsf(dl) > ha[j] + el + pad);
                        }
                    }
                    int kind = 0;
                    unsigned ex = 0u, ey = 0u;
                    if (ov) {
                        const bool leafA = aLeft < 0, leafB = bLeft < 0;
                        if (leafA && leafB) {
                            kind = 1;
                            ex = (unsigned)(-1 - aLeft) | tag;
                            ey = (unsigned)(-1 - bLeft);
                        } else {
                            bool descendRobot;
                            if (leafA) descendRobot = false;
                            else if (leafB) descendRobot = true;
                            else descendRobot = fmaxf(hx, fmaxf(hy, hz)) > fmaxf(hb[0], fmaxf(hb[1], hb[2]));
                            kind = 2;
                            ex = descendRobot ? ((unsigned)aLeft | tag) : (aId | tag | SIDE_ENV);
                            ey = descendRobot ? bId : (unsigned)bLeft;
                        }
                    }
                    if (second) kind1 = kind, ex1 = ex, ey1 = ey;
                    else kind0 = kind, ex0 = ex, ey0 = ey;
                }
            }
            const unsigned t0 = __ballot_sync(FULL_MASK_, kind0 == 1), t1 = __ballot_sync(FULL_MASK_, kind1 == 1);
            const unsigned s0 = __ballot_sync(FULL_MASK_, kind0 == 2), s1 = __ballot_sync(FULL_MASK_, kind1 == 2);
            if (kind0 == 1) triQ[nt + __popc(t0 & ltMask)] = make_uint2(ex0, ey0);
            if (kind1 == 1) triQ[nt + __popc(t0) + __popc(t1 & ltMask)] = make_uint2(ex1, ey1);
            if (kind0 == 2) stack[n + __popc(s0 & ltMask)] = make_uint2(ex0, ey0);
            if (kind1 == 2) stack[n + __popc(s0) + __popc(s1 & ltMask)] = make_uint2(ex1, ey1);
            nt += __popc(t0) + __popc(t1);
            n += __popc(s0) + __popc(s1);
            if (n > NODE_STACK - 64) {  // cannot happen with the throttle above unless trees are > ~60 deep
                err |= GEOM_ERR_STACK;
                if (lane == 0) atomicAdd(pool.ctl + CTL_IDLE, 1u);  // leaving: keep the termination count right
                break;
            }
            __syncwarp();
        }
        while (nt >= TRI_DRAIN_AT || (n == 0 && nt > 0 && exhausted)) {  // (a round adds up to 64: drain below 32 before the next one)
            const int p = nt < 32 ? nt : 32;
#ifdef MESH_DEBUG_STATS
            ++dbgTriRounds;
#endif
            bool mine = lane < p;
            unsigned hitBit = 0u;
            uint2 tp = make_uint2(0u, 0u);
            if (mine) tp = triQ[nt - 1 - lane];
            const unsigned slot = tp.x >> SLOT_SHIFT;
            mine = mine && !((dead >> slot) & 1u);
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
                bool ov = true, ovBand = true;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float plo = fminf(P[0][c], fminf(P[1][c], P[2][c])), phi = fmaxf(P[0][c], fmaxf(P[1][c], P[2][c]));
                    const float qlo = fminf(Q[0][c], fminf(Q[1][c], Q[2][c])), qhi = fmaxf(Q[0][c], fmaxf(Q[1][c], Q[2][c]));
                    ov = ov && !(plo > qhi || qlo > phi);
                    if (NEAR) ovBand = ovBand && !(plo - w.tol > qhi || qlo - w.tol > phi);
                }
                if (NEAR) {
                    if (ovBand) {
                        const int r = triTriNear(P, Q, w.tol);
                        if (ov && (r & 1)) hitBit = 1u << slot;
                        if (r & 2) w.nearOut[items[slot]] = 1;
                    }
                } else if (ov && triTriIntersect(P, Q)) {
                    hitBit = 1u << slot;
                }
            }
            __syncwarp();
            nt -= p;
            const unsigned hits = __reduce_or_sync(FULL_MASK_, hitBit) & ~dead;
            if (hits) {
                if ((hits >> lane) & 1u) w.ok[items[lane]] = 0;  // lane s reports slot s
                // every slot working on the same item is finished too (stale items of free slots may match: harmless)
                const unsigned same = __match_any_sync(FULL_MASK_, items[lane]);
                dead |= __ballot_sync(FULL_MASK_, (same & hits) != 0u);
            }
        }
    }
#ifdef MESH_DEBUG_STATS
    {
        unsigned bvw = cnt.bv, triw = cnt.tri;
        for (int o = 16; o > 0; o >>= 1) bvw += __shfl_down_sync(FULL_MASK_, bvw, o), triw += __shfl_down_sync(FULL_MASK_, triw, o);
        const unsigned gw = blockIdx.x * MESH_WARPS + warp;
        if (lane == 0 && gw < 8192) gDbgT[MODE][gw][3] = bvw, gDbgT[MODE][gw][4] = triw;
    }
#endif
    flushCounters(cnt, err, stats, lane);
#ifdef MESH_DEBUG_STATS
    if (lane == 0) {
        atomicAdd(stats + 5, dbgRounds);
        atomicAdd(stats + 6, dbgThrottled);
        atomicMax(stats + 7, (unsigned long long)dbgRounds);
        const unsigned gw = blockIdx.x * MESH_WARPS + warp;
        if (gw < 8192) gDbgT[MODE][gw][0] = dbgT0, gDbgT[MODE][gw][1] = gtimer(), gDbgT[MODE][gw][2] = dbgRounds, gDbgT[MODE][gw][5] = (dbgSpin << 32) | dbgPoll, gDbgT[MODE][gw][6] = dbgIdleT, gDbgT[MODE][gw][7] = (dbgAdopted << 32) | dbgDonated;
        atomicMax(stats + 3, (unsigned long long)cnt.states);
    }
#endif
}

// per edge: steps = ceil(distance(from,to) * (1/stepSize)) (:64,:78), the number of its states outside the
// coarse set, and ok preset to 1
template <typename S>
__global__ void meshStepsKernel(DevSpace<S> sp, const S* __restrict__ from, const S* __restrict__ to, uint32_t n, S invStep,
                                uint32_t* __restrict__ steps, unsigned long long* __restrict__ counts, uint8_t* __restrict__ ok,
                                unsigned long long* stats) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (e == 0) counts[n] = 0;
    const S* a = from + (size_t)e * 7;
    const S* b = to + (size_t)e * 7;
    const S dist = dev::distance<S>(sp, [&](int c) { return __ldg(a + c); }, [&](int c) { return __ldg(b + c); });
    const S fs = ceil(dist * invStep);
    uint32_t s = 0;
    if (!(fs < S(2147483648.0))) {
        atomicOr(stats + 4, (unsigned long long)GEOM_ERR_STEPS);
        ok[e] = 0;
    } else {
        s = (uint32_t)fs;
        ok[e] = 1;
    }
    steps[e] = s;
    counts[e] = s > (uint32_t)COARSE_IDS ? s - 1u : 0u;
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
    bool sah = true;  // false: median splits along the longest axis (the r1 / early r2 builder)
    static constexpr int sahMin = 2;
    explicit Builder(const std::vector<HostTri>& t) : tris(t) {}
    float centroid(int id, int ax) const { return tris[id].v[0][ax] + tris[id].v[1][ax] + tris[id].v[2][ax]; }

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
        int mid = (lo + hi) / 2;
        bool split = false;
        if (sah && hi - lo > sahMin && depth < 24) {
            // binned surface-area heuristic over the three axes (16 bins of the centroid range): the split that minimises
            // count(left) area(left) + count(right) area(right); ties and degenerate ranges fall back to the median split
            constexpr int BINS = 16, MAXBINS = BINS;
            double best = INFINITY;
            int bestAxis = -1;
            float bestPos = 0;
            for (int ax = 0; ax < 3; ++ax) {
                float cmin = INFINITY, cmax = -INFINITY;
                for (int i = lo; i < hi; ++i) {
                    const float c = centroid(ids[i], ax);
                    cmin = std::fmin(cmin, c), cmax = std::fmax(cmax, c);
                }
                if (!(cmax > cmin)) continue;
                struct Bin { int n = 0; float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}; } bins[MAXBINS];
                const float scale = BINS / (cmax - cmin);
                for (int i = lo; i < hi; ++i) {
                    int bi = (int)((centroid(ids[i], ax) - cmin) * scale);
                    bi = bi < 0 ? 0 : (bi >= BINS ? BINS - 1 : bi);
                    Bin& bn = bins[bi];
                    ++bn.n;
                    for (int v = 0; v < 3; ++v)
                        for (int c = 0; c < 3; ++c) bn.lo[c] = std::fmin(bn.lo[c], tris[ids[i]].v[v][c]), bn.hi[c] = std::fmax(bn.hi[c], tris[ids[i]].v[v][c]);
                }
                auto area = [](const float* l, const float* h) {
                    const double dx = (double)h[0] - l[0], dy = (double)h[1] - l[1], dz = (double)h[2] - l[2];
                    return dx * dy + dy * dz + dz * dx;
                };
                double rightArea[MAXBINS];
                int rightN[MAXBINS];
                {
                    float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY};
                    int cnt = 0;
                    for (int k = BINS - 1; k > 0; --k) {
                        for (int c = 0; c < 3; ++c) l[c] = std::fmin(l[c], bins[k].lo[c]), h[c] = std::fmax(h[c], bins[k].hi[c]);
                        cnt += bins[k].n;
                        rightN[k] = cnt, rightArea[k] = cnt ? area(l, h) : 0.0;
                    }
                }
                float l[3] = {INFINITY, INFINITY, INFINITY}, h[3] = {-INFINITY, -INFINITY, -INFINITY};
                int cnt = 0;
                for (int k = 0; k + 1 < BINS; ++k) {
                    for (int c = 0; c < 3; ++c) l[c] = std::fmin(l[c], bins[k].lo[c]), h[c] = std::fmax(h[c], bins[k].hi[c]);
                    cnt += bins[k].n;
                    if (cnt == 0 || rightN[k + 1] == 0) continue;
                    const double cost = cnt * area(l, h) + rightN[k + 1] * rightArea[k + 1];
                    if (cost < best) best = cost, bestAxis = ax, bestPos = cmin + (k + 1) / scale;
                }
            }
            if (bestAxis >= 0) {
                const auto it = std::stable_partition(ids.begin() + lo, ids.begin() + hi, [&](int x) { return centroid(x, bestAxis) < bestPos; });
                const int m = (int)(it - ids.begin());
                if (m > lo && m < hi) mid = m, split = true;
            }
        }
        if (!split) {
            mid = (lo + hi) / 2;
            std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int x, int y) {
                const float cx = tris[x].v[0][axis] + tris[x].v[1][axis] + tris[x].v[2][axis];
                const float cy = tris[y].v[0][axis] + tris[y].v[1][axis] + tris[y].v[2][axis];
                return cx < cy || (cx == cy && x < y);
            });
        }
        const int l = build(ids, lo, mid, depth + 1);
        const int r = build(ids, mid, hi, depth + 1);
        nodes[me].left = l;
        nodes[me].right = r;
        return me;
    }
};

// Node order of the device image: the first `prefix` nodes breadth-first from the root (complete upper levels: what a
// CTA stages in shared memory), then the rest of the breadth-first frontier, then the subtrees below it depth-first, the
// two children of a node placed TOGETHER before either subtree.  Everywhere the right child follows the left one
// (right == left + 1): a traversal entry names a pair of siblings by the first.
std::vector<BvhNode> stagingOrder(const std::vector<BvhNode>& in, size_t prefix) {
    const size_t n = in.size();
    std::vector<int> order;  // new position -> old index
    order.reserve(n);
    std::vector<int> frontier{0};
    size_t head = 0;
    while (head < frontier.size() && order.size() < prefix) {  // breadth-first part
        const int i = frontier[head++];
        order.push_back(i);
        if (in[i].left >= 0) frontier.push_back(in[i].left), frontier.push_back(in[i].right);
    }
    const size_t firstBelow = order.size();
    for (size_t f = head; f < frontier.size(); ++f) order.push_back(frontier[f]);  // sibling pairs stay together across the cut
    std::vector<int> stack;
    for (size_t k = order.size(); k-- > firstBelow;) stack.push_back(order[k]);
    while (!stack.empty()) {
        const int i = stack.back();
        stack.pop_back();
        if (in[i].left < 0) continue;
        order.push_back(in[i].left), order.push_back(in[i].right);
        stack.push_back(in[i].right), stack.push_back(in[i].left);
    }
    std::vector<int> where(n);
    for (size_t k = 0; k < n; ++k) where[order[k]] = (int)k;
    std::vector<BvhNode> out(n);
    for (size_t k = 0; k < n; ++k) {
        out[k] = in[order[k]];
        if (out[k].left >= 0) out[k].left = where[out[k].left], out[k].right = where[out[k].right];
    }
    return out;
}

int uploadMesh(mptg_ctx* ctx, const float* tris9, uint32_t n, size_t stagePrefix, void** nodesDev, void** trisDev, int* depth, uint32_t* staged,
               int* rootLeft) {
    std::vector<HostTri> tris(n);
    std::vector<TriPad> pad(n ? n : 1);
    for (uint32_t i = 0; i < n; ++i)
        for (int v = 0; v < 3; ++v) {
            for (int c = 0; c < 3; ++c) tris[i].v[v][c] = pad[i].v[v][c] = tris9[(size_t)i * 9 + v * 3 + c];
            pad[i].v[v][3] = 0.0f;
        }
    Builder b(tris);
    b.sah = getenv("MPTG_MESH_MEDIAN_SPLIT") == nullptr;  // (the earlier builder stays selectable for comparisons)
    if (n) {
        std::vector<int> ids(n);
        for (uint32_t i = 0; i < n; ++i) ids[i] = (int)i;
        b.nodes.reserve(2 * (size_t)n);
        b.build(ids, 0, (int)n, 0);
        if (b.sah && b.maxDepth > 28) {  // a degenerate mesh: the balanced tree instead (depth <= 25 + 1), so that two trees always fit the stack budget
            b.sah = false, b.maxDepth = 0, b.nodes.clear();
            for (uint32_t i = 0; i < n; ++i) ids[i] = (int)i;
            b.build(ids, 0, (int)n, 0);
        }
    } else {
        BvhNode placeholder{};  // never traversed (an empty mesh cannot collide); a leaf, so that nothing hangs off it
        placeholder.left = placeholder.right = -1;
        b.nodes.push_back(placeholder);
    }
    *depth = b.maxDepth;
    if (n) b.nodes = stagingOrder(b.nodes, stagePrefix);
    *staged = (uint32_t)std::min(stagePrefix, b.nodes.size());
    *rootLeft = b.nodes[0].left;
    // device image: centre / half-extent form (what the box test needs), half extents rounded outwards so that
    // [c - h, c + h] contains [lo, hi]
    for (BvhNode& n : b.nodes)
        for (int c = 0; c < 3; ++c) {
            const float lo = n.lo[c], hi = n.hi[c];
            const float ctr = 0.5f * (lo + hi);
            float h = std::fmax(hi - ctr, ctr - lo);
            h = std::nextafter(h, INFINITY);
            n.lo[c] = ctr;  // device meaning: centre
            n.hi[c] = h;    // device meaning: half extent
        }
    MPTG_CUDA(ctx, cudaMalloc(nodesDev, b.nodes.size() * sizeof(BvhNode)));
    if (int rc = uploadSync(ctx, *nodesDev, b.nodes.data(), b.nodes.size() * sizeof(BvhNode))) return rc;
    MPTG_CUDA(ctx, cudaMalloc(trisDev, pad.size() * sizeof(TriPad)));
    if (int rc = uploadSync(ctx, *trisDev, pad.data(), pad.size() * sizeof(TriPad))) return rc;
    return MPTG_OK;
}

}  // namespace

int meshCreate(mptg_ctx* ctx, int /*scalar*/, uint32_t nr, const float* robotTris, uint32_t ne, const float* envTris,
               MeshData** out) {
    auto* m = new MeshData();
    // shared-memory budget of a CTA: its warps' private arrays, then the robot's hierarchy (whole, if it fits in half of
    // what is left), then the environment's upper levels
    const size_t nodeBudget = (MESH_SMEM_LIMIT - 1024 - (size_t)MESH_WARPS * MESH_WARP_SMEM) / sizeof(BvhNode);
    const size_t nodesR = nr ? 2 * (size_t)nr - 1 : 1, nodesE = ne ? 2 * (size_t)ne - 1 : 1;
    size_t stageR = std::min(nodesR, nodeBudget / 2);
    const size_t stageE = std::min(nodesE, nodeBudget - stageR);
    stageR = std::min(nodesR, nodeBudget - stageE);
    uint32_t sr = 0, se = 0;
    int rootR = -1, rootE = -1;
    int rc = (nodesR > NODE_MASK || nodesE > NODE_MASK) ? fail(ctx, MPTG_ERR_CAPACITY, "meshCreate: more than 2^25 triangles in a mesh") : MPTG_OK;
    if (!rc) rc = uploadMesh(ctx, robotTris, nr, stageR, &m->mem[0], &m->mem[1], &m->depthR, &sr, &rootR);
    if (!rc) rc = uploadMesh(ctx, envTris, ne, stageE, &m->mem[2], &m->mem[3], &m->depthE, &se, &rootE);
    // the entry every state starts from (meshFlatKernel): the children of the robot's root against the environment's
    // root; a one-triangle robot: the children of the environment's root against it; two single triangles: that pair
    if (rootR >= 0) m->dev.rootX = (uint32_t)rootR, m->dev.rootY = 0u, m->dev.rootIsTri = 0u;
    else if (rootE >= 0) m->dev.rootX = SIDE_ENV, m->dev.rootY = (uint32_t)rootE, m->dev.rootIsTri = 0u;
    else m->dev.rootX = (uint32_t)(-1 - rootR), m->dev.rootY = (uint32_t)(-1 - rootE), m->dev.rootIsTri = 1u;
    m->dev.stageR = sr;
    m->dev.stageE = se;
    if (!rc) {
        cudaError_t e = cudaMalloc(&m->workCounter, CTL_BYTES);
        if (e == cudaSuccess) e = cudaMalloc(&m->poolBlocks, (size_t)POOL_BLOCKS * sizeof(DonBlock));
        if (e != cudaSuccess) rc = fail(ctx, MPTG_ERR_CUDA, "meshCreate: %s", cudaGetErrorString(e));
    }
    if (!rc && (unsigned)(ctx->smCount * MESH_MIN_CTAS * MESH_WARPS) + 2u * MAX_POLLERS > POOL_BLOCKS)
        rc = fail(ctx, MPTG_ERR_CAPACITY, "meshCreate: device has more resident warps than the donation ring allows");
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
    if (ne) {  // contact band: 1e-6 of the diagonal of the environment's bounding box (the oracle's definition)
        double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (size_t i = 0; i < (size_t)ne * 3; ++i)
            for (int c = 0; c < 3; ++c) {
                lo[c] = std::fmin(lo[c], (double)envTris[i * 3 + c]);
                hi[c] = std::fmax(hi[c], (double)envTris[i * 3 + c]);
            }
        m->band = (float)(1e-6 * std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2])));
    }
    *out = m;
    return MPTG_OK;
}

void meshDestroy(MeshData* m) {
    if (!m) return;
    for (void* p : m->mem) cudaFree(p);
    cudaFree(m->workCounter);
    cudaFree(m->poolBlocks);
    cudaFree(m->steps), cudaFree(m->counts), cudaFree(m->offs), cudaFree(m->scanTemp);
    delete m;
}

namespace {

template <int MODE, bool NEAR, typename S>
void launchFlat(mptg_ctx* ctx, MeshData* md, const MeshWork<S>& w, unsigned long long units, int pass, unsigned long long* stats) {
    unsigned long long* counter = md->workCounter + (size_t)pass * CTL_COUNTER_STRIDE;
    unsigned int* ctl = reinterpret_cast<unsigned int*>(md->workCounter + 2 * CTL_COUNTER_STRIDE) + (size_t)pass * CTL_PASS_WORDS;
    const MeshPool pool{md->poolBlocks, ctl, ctl + CTL_WORDS};
    // persistent grid; a warp starts up to REFILL_MAX states at a time
    const unsigned long long want = units / (MESH_WARPS * REFILL_MAX) + 1ull;
    const unsigned long long cap = (unsigned long long)ctx->smCount * MESH_MIN_CTAS;
    const int grid = (int)(want < cap ? (want ? want : 1) : cap);
    const size_t smem = ((size_t)md->dev.stageR + md->dev.stageE) * sizeof(BvhNode) + (size_t)MESH_WARPS * MESH_WARP_SMEM;
    static bool attrSet = false;  // per instantiation
    if (!attrSet) {
        cudaFuncSetAttribute(meshFlatKernel<S, MODE, NEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MESH_SMEM_LIMIT - 1024));
        attrSet = true;
    }
    meshFlatKernel<S, MODE, NEAR><<<grid, MESH_WARPS * 32, smem, ctx->stream>>>(md->dev, w, counter, pool, stats);
}

int ensureEdgeWork(mptg_ctx* ctx, MeshData* md, uint32_t n) {
    size_t scanBytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)(n + 1u));
    if (md->workEdges >= n && md->scanBytes >= scanBytes) return MPTG_OK;
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(md->steps), cudaFree(md->counts), cudaFree(md->offs), cudaFree(md->scanTemp);
    md->steps = nullptr, md->counts = md->offs = nullptr, md->scanTemp = nullptr, md->workEdges = 0, md->scanBytes = 0;
    const size_t cap = (size_t)n + (size_t)n / 4 + 1024;
    cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)(cap + 1));
    MPTG_CUDA(ctx, cudaMalloc(&md->steps, cap * sizeof(uint32_t)));
    MPTG_CUDA(ctx, cudaMalloc(&md->counts, (cap + 1) * sizeof(unsigned long long)));
    MPTG_CUDA(ctx, cudaMalloc(&md->offs, (cap + 1) * sizeof(unsigned long long)));
    MPTG_CUDA(ctx, cudaMalloc(&md->scanTemp, scanBytes ? scanBytes : 16));
    md->workEdges = (uint32_t)(cap > 0xffffffffull ? 0xffffffffull : cap);
    md->scanBytes = scanBytes;
    return MPTG_OK;
}

template <typename S>
int meshValidT(mptg_geom* g, const S* states, uint32_t n, uint8_t* ok, uint8_t* nearOut) {
    mptg_ctx* ctx = g->ctx;
    MeshData* md = g->mesh;
    MPTG_CUDA(ctx, cudaMemsetAsync(ok, 1, n, ctx->stream));
    if (nearOut) MPTG_CUDA(ctx, cudaMemsetAsync(nearOut, 0, n, ctx->stream));
    if (md->dev.nR == 0 || md->dev.nE == 0) return MPTG_OK;  // nothing can collide
    MPTG_CUDA(ctx, cudaMemsetAsync(md->workCounter, 0, CTL_BYTES, ctx->stream));
    MeshWork<S> w{};
    w.a = states, w.n = n, w.ok = ok, w.nearOut = nearOut, w.tol = md->band;
    if (nearOut) launchFlat<WORK_STATES, true>(ctx, md, w, n, 0, g->devStats);
    else launchFlat<WORK_STATES, false>(ctx, md, w, n, 0, g->devStats);
    MPTG_LAUNCHED(ctx);
    return MPTG_OK;
}

template <typename S>
int meshLinkT(mptg_geom* g, const mptg_space_desc* space, const S* from, const S* to, uint32_t n, double step, uint8_t* ok, uint8_t* nearOut) {
    mptg_ctx* ctx = g->ctx;
    MeshData* md = g->mesh;
    if ((unsigned long long)n * COARSE_IDS > 0xffffffffull) return fail(ctx, MPTG_ERR_CAPACITY, "mptg_link_batch: at most 2^29 mesh edges per call");
    if (int rc = ensureEdgeWork(ctx, md, n)) return rc;
    MPTG_CUDA(ctx, cudaMemsetAsync(md->workCounter, 0, CTL_BYTES, ctx->stream));
    const DevSpace<S> sp = makeDevSpace<S>(*space);
    const S invStep = S(1) / (S)step;  // discrete_motion_validator.hpp:64
    meshStepsKernel<S><<<(n + 255) / 256, 256, 0, ctx->stream>>>(sp, from, to, n, invStep, md->steps, md->counts, ok, g->devStats);
    MPTG_LAUNCHED(ctx);
    if (nearOut) MPTG_CUDA(ctx, cudaMemsetAsync(nearOut, 0, n, ctx->stream));
    if (md->dev.nR == 0 || md->dev.nE == 0) return MPTG_OK;
    size_t scanBytes = md->scanBytes;
    MPTG_CUDA(ctx, cub::DeviceScan::ExclusiveSum(md->scanTemp, scanBytes, md->counts, md->offs, (int)(n + 1u), ctx->stream));
    MPTG_LAUNCHED(ctx);
    MeshWork<S> w{};
    w.a = from, w.b = to, w.n = n, w.steps = md->steps, w.offs = md->offs, w.ok = ok, w.sp = sp, w.nearOut = nearOut, w.tol = md->band;
    // the length of the list is only known on the device; at least the coarse ids are there
    if (nearOut) launchFlat<WORK_EDGES, true>(ctx, md, w, (unsigned long long)n * COARSE_IDS, 0, g->devStats);
    else launchFlat<WORK_EDGES, false>(ctx, md, w, (unsigned long long)n * COARSE_IDS, 0, g->devStats);
    MPTG_LAUNCHED(ctx);
#ifdef MESH_DEBUG_STATS
    if (getenv("MPTG_DEBUG_TIMELINE")) {
        cudaStreamSynchronize(ctx->stream);
        static unsigned long long h[3][8192][8];
        cudaMemcpyFromSymbol(h, gDbgT, sizeof h);
        for (int mode = 1; mode <= 1; ++mode) {
            const int nw = ctx->smCount * MESH_MIN_CTAS * MESH_WARPS;
            unsigned long long t0 = ~0ull;
            for (int i = 0; i < nw; ++i) t0 = h[mode][i][0] < t0 ? h[mode][i][0] : t0;
            std::vector<double> endv, rounds;
            for (int i = 0; i < nw; ++i) endv.push_back((h[mode][i][1] - t0) * 1e-3), rounds.push_back((double)h[mode][i][2]);
            unsigned long long totA = 0, totD = 0, nAd = 0;
            for (int i = 0; i < nw; ++i) totA += h[mode][i][7] >> 32, totD += h[mode][i][7] & 0xffffffffull, nAd += (h[mode][i][7] >> 32) ? 1 : 0;
            unsigned long long nSpin = 0, nPoll = 0, totSpin = 0, totPoll = 0;
            for (int i = 0; i < nw; ++i) {
                const unsigned long long sp = h[mode][i][5] >> 32, po = h[mode][i][5] & 0xffffffffull;
                nSpin += sp ? 1 : 0, nPoll += po ? 1 : 0, totSpin += sp, totPoll += po;
            }
            fprintf(stderr, "[mptg] mode %d donations %llu adoptions %llu by %llu warps; %llu warps spun (%llu iterations), %llu warps polled (%llu iterations)\n", mode, totD, totA, nAd, nSpin, totSpin, nPoll, totPoll);
            std::vector<double> st, si;
            for (int i = 0; i < nw; ++i) st.push_back((h[mode][i][0] - t0) * 1e-3), si.push_back(h[mode][i][6] ? (h[mode][i][6] - t0) * 1e-3 : -1.0);
            std::sort(st.begin(), st.end()), std::sort(si.begin(), si.end());
            fprintf(stderr, "[mptg] mode %d start us p50 %.0f p99 %.0f max %.0f | first idle us p10 %.0f p50 %.0f p90 %.0f max %.0f\n", mode, st[nw / 2], st[nw * 99 / 100], st[nw - 1], si[nw / 10], si[nw / 2], si[nw * 9 / 10], si[nw - 1]);
            std::vector<double> se = endv, sr = rounds;
            std::sort(se.begin(), se.end()), std::sort(sr.begin(), sr.end());
            fprintf(stderr, "[mptg] mode %d warp end us: p10 %.0f p50 %.0f p90 %.0f p99 %.0f max %.0f | rounds p50 %.0f p90 %.0f p99 %.0f max %.0f\n", mode,
                    se[nw / 10], se[nw / 2], se[nw * 9 / 10], se[nw * 99 / 100], se[nw - 1], sr[nw / 2], sr[nw * 9 / 10], sr[nw * 99 / 100], sr[nw - 1]);
            // the last 5 warps
            for (int k = 0; k < 5; ++k) {
                int best = 0;
                for (int i = 0; i < nw; ++i) if (endv[i] > endv[best]) best = i;
                fprintf(stderr, "   late warp %d (cta %d): end %.0f us rounds %.0f bv %llu tri %llu states %llu triRounds %llu adopted %llu donated %llu\n", best, best / MESH_WARPS, endv[best], rounds[best], h[mode][best][3], h[mode][best][4], h[mode][best][5], h[mode][best][6], h[mode][best][7] >> 32, h[mode][best][7] & 0xffffffffull);
                endv[best] = -1;
            }
        }
    }
#endif
    return MPTG_OK;
}

}  // namespace

int meshValidDev(mptg_geom* g, const void* states, uint32_t n, uint8_t* ok, uint8_t* nearOut) {
    return g->scalar == MPTG_F32 ? meshValidT<float>(g, (const float*)states, n, ok, nearOut) : meshValidT<double>(g, (const double*)states, n, ok, nearOut);
}

int meshLinkDev(mptg_geom* g, const mptg_space_desc* space, const void* from, const void* to, uint32_t n, double step,
                uint8_t* ok, uint8_t* nearOut) {
    return g->scalar == MPTG_F32 ? meshLinkT<float>(g, space, (const float*)from, (const float*)to, n, step, ok, nearOut)
                                 : meshLinkT<double>(g, space, (const double*)from, (const double*)to, n, step, ok, nearOut);
}

double meshBand(const MeshData* m) { return m ? (double)m->band : 0.0; }

}  // namespace mptg
