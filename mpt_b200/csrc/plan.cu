// plan.cu -- planner stages that keep a sample wave on the device:
//   * uniform sampling of box / SO(2) / SO(3) / compound states from a counter-based generator
//     (src/mpt/uniform_box_sampler.hpp:60-68, impl/uniform_sampler_so2.hpp:58-64,
//      impl/uniform_sampler_so3.hpp:55-68, impl/uniform_sampler_cartesian.hpp:75-78),
//   * a device-resident PRRT: the tree (states, parents), the nearest-neighbour structure and every
//     stage of Worker::addSample (src/mpt/impl/prrt/prrt.hpp:411-452) stay on the GPU; per wave the host
//     launches kernels and reads back two words (nodes added, goal node).
//
// Sampling.  The reference draws from std::uniform_real_distribution over a seeded std::mt19937_64, one
// generator per worker thread, seeded from std::random_device: there is no sequence to reproduce, only
// the distributions.  Here uniform j of sample g is a pure function of (seed, g, j) -- Philox4x32-10,
// counter (g_lo, g_hi, j / 4, 0), key = seed -- so a wave is generated in parallel, the same on any
// number of GPUs, and can be replayed bit for bit by the CPU oracle.
//   u in [0,1):  float  (w >> 8) * 2^-24;   double  ((w_{2j} << 32 | w_{2j+1}) >> 11) * 2^-53
//   uniform 0 is the goal-bias draw (always consumed), then per part, in part order:
//   LP coordinate   q = u * (hi - lo) + lo                 (what uniform_real_distribution(lo,hi) computes)
//   SO2 coordinate  q = u * (pi - (-pi)) + (-pi)
//   SO3             a = u0, b = u1 * 2pi, c = u2 * 2pi;  (w,x,y,z) = (sqrt(1-a) sin b, sqrt(1-a) cos b, sqrt(a) sin c, sqrt(a) cos c)
// with the shared sqrt / sincos of mptg_fpmath.h and unfused arithmetic.
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <chrono>
#include <cstdio>
#include <vector>

#include "geom.cuh"

namespace mptg {

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c[0];
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0, c[1] = n1, c[2] = n2, c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// stream of uniforms of one sample
template <typename S>
struct UniformStream {
    unsigned long long seed, g;
    uint32_t w[4];
    uint32_t word = 0;  // next 32-bit word of the sample's stream
    __device__ __forceinline__ UniformStream(unsigned long long seed_, unsigned long long g_) : seed(seed_), g(g_) {}
    __device__ __forceinline__ uint32_t next32() {
        if ((word & 3u) == 0u) {
            w[0] = (uint32_t)g, w[1] = (uint32_t)(g >> 32), w[2] = word >> 2, w[3] = 0u;
            philox4x32_10(w, (uint32_t)seed, (uint32_t)(seed >> 32));
        }
        return w[word++ & 3u];
    }
    __device__ __forceinline__ S next() {
        if (sizeof(S) == 4) {
            return (S)((float)(next32() >> 8) * 5.9604644775390625e-08f);
        } else {
            const unsigned long long hi = next32();
            const unsigned long long lo = next32();
            return (S)((double)(((hi << 32) | lo) >> 11) * 1.1102230246251565404e-16);
        }
    }
};

// uniforms -> state; U is a callable returning the next uniform
template <typename S, typename U>
__device__ __forceinline__ void transformSample(const DevSpace<S>& sp, const S* __restrict__ lo, const S* __restrict__ hi, U&& next, S* q) {
    for (int i = 0; i < sp.nParts; ++i) {
        const int off = sp.off[i];
        if (sp.kind[i] == MPTG_PART_SO3) {
            const S a = next();
            const S twoPi = S(2) * fp::consts<S>::pi();
            const S b = next() * twoPi;
            const S c = next() * twoPi;
            S sb, cb, sc, cc;
            fp::sincos_(b, &sb, &cb);
            fp::sincos_(c, &sc, &cc);
            const S r1 = fp::sqrt_(S(1) - a), r2 = fp::sqrt_(a);
            q[off + 3] = r1 * sb;  // w
            q[off + 0] = r1 * cb;  // x
            q[off + 1] = r2 * sc;  // y
            q[off + 2] = r2 * cc;  // z
        } else if (sp.kind[i] == MPTG_PART_SO2) {
            const S pi = fp::consts<S>::pi();
            for (int c = 0; c < sp.dim[i]; ++c) q[off + c] = next() * (pi - (-pi)) + (-pi);
        } else {
            for (int c = 0; c < sp.dim[i]; ++c) q[off + c] = next() * (hi[off + c] - lo[off + c]) + lo[off + c];
        }
    }
}

template <typename S>
__global__ void sampleKernel(DevSpace<S> sp, const S* __restrict__ lo, const S* __restrict__ hi, unsigned long long seed,
                             unsigned long long first, uint32_t n, const S* __restrict__ goal, S goalBias, S* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    UniformStream<S> us(seed, first + i);
    S q[MPTG_MAX_SCALARS];
    const S biasDraw = us.next();
    if (goal != nullptr && biasDraw < goalBias) {  // prrt.hpp:377-379: the goal sampler of GoalState returns the goal state
        for (int c = 0; c < sp.D; ++c) out[(size_t)i * sp.D + c] = goal[c];
        return;
    }
    transformSample<S>(sp, lo, hi, [&]() { return us.next(); }, q);
    for (int c = 0; c < sp.D; ++c) out[(size_t)i * sp.D + c] = q[c];
}

template <typename S>
__global__ void transformKernel(DevSpace<S> sp, const S* __restrict__ lo, const S* __restrict__ hi, const S* __restrict__ uniforms,
                                int perState, uint32_t n, S* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const S* u = uniforms + (size_t)i * perState;
    int j = 0;
    S q[MPTG_MAX_SCALARS];
    transformSample<S>(sp, lo, hi, [&]() { return u[j++]; }, q);
    for (int c = 0; c < sp.D; ++c) out[(size_t)i * sp.D + c] = q[c];
}

// ---- device-resident PRRT stages
// steer (prrt.hpp:416-434): near = state of the nearest node; d == 0 drops the sample; d > range pulls it in
template <typename S>
__global__ void prrtSteerKernel(DevSpace<S> sp, const S* __restrict__ nodes, const S* __restrict__ samples, const uint32_t* __restrict__ nearIdx,
                                const S* __restrict__ nearDist, const uint32_t* __restrict__ nearCnt, uint32_t n, S range, S* __restrict__ from,
                                S* __restrict__ to, uint8_t* __restrict__ alive) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int D = sp.D;
    const S* ps = samples + (size_t)i * D;
    S* pf = from + (size_t)i * D;
    S* pt = to + (size_t)i * D;
    const S d = nearDist[i];
    const bool live = nearCnt[i] != 0 && !(d == S(0));
    alive[i] = live ? 1 : 0;
    if (!live) {  // keep the buffers well defined: a zero-length edge at the sample
        for (int c = 0; c < D; ++c) pf[c] = pt[c] = ps[c];
        return;
    }
    const S* pn = nodes + (size_t)nearIdx[i] * D;
    S a[MPTG_MAX_SCALARS];
    for (int c = 0; c < D; ++c) a[c] = pf[c] = pn[c];
    if (d > range) {
        S b[MPTG_MAX_SCALARS], q[MPTG_MAX_SCALARS];
        for (int c = 0; c < D; ++c) b[c] = ps[c];
        dev::interpolate<S>(sp, a, b, fp::div_(range, d), q);
        for (int c = 0; c < D; ++c) pt[c] = q[c];
    } else {
        for (int c = 0; c < D; ++c) pt[c] = ps[c];
    }
}

// ordered compaction by one CTA (defined with the PRRT* kernels below)
__global__ void starCompactKernel(const uint8_t* __restrict__ f0, const uint8_t* __restrict__ f1, const uint8_t* __restrict__ f2, uint32_t n,
                                  uint32_t* __restrict__ out, uint32_t* __restrict__ count, uint32_t* __restrict__ init);

__global__ void prrtFlagKernel(const uint8_t* __restrict__ alive, const uint8_t* __restrict__ okValid, const uint8_t* __restrict__ okLink,
                               uint32_t n, uint8_t* __restrict__ keep) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keep[i] = (alive[i] && okValid[i] && okLink[i]) ? 1 : 0;
}

// append the survivors in sample order (prrt.hpp:441-450): node states, parents, goal test (goal_state.hpp:64-69)
template <typename S>
__global__ void prrtAppendKernel(DevSpace<S> sp, const uint32_t* __restrict__ sel, const uint32_t* __restrict__ nSel, const S* __restrict__ to,
                                 const uint32_t* __restrict__ nearIdx, uint32_t size, uint32_t capacity, const S* __restrict__ goal, S goalRadius,
                                 S* __restrict__ nodes, uint32_t* __restrict__ parent, S* __restrict__ fresh, uint32_t* __restrict__ result) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t count = *nSel;
    if (count > capacity - size) count = capacity - size;
    if (j == 0) result[0] = count;
    if (j >= count) return;
    const int D = sp.D;
    const uint32_t i = sel[j];
    const S* pt = to + (size_t)i * D;
    S* pn = nodes + (size_t)(size + j) * D;
    S* pfr = fresh + (size_t)j * D;
    for (int c = 0; c < D; ++c) pn[c] = pfr[c] = pt[c];
    parent[size + j] = nearIdx[i];
    if (goal != nullptr) {
        const S d = dev::distance<S>(sp, [&](int c) { return pt[c]; }, [&](int c) { return goal[c]; });
        if (d <= goalRadius) atomicMin(result + 1, size + j);  // the first goal node in insertion order
    }
}

}  // namespace mptg

using namespace mptg;

namespace {

int spaceUniforms(const mptg_space_desc* sp) {
    int n = 0;
    for (int i = 0; i < sp->n_parts; ++i) n += sp->part[i].kind == MPTG_PART_SO3 ? 3 : sp->part[i].dim;
    return n;
}

// bounds (doubles, one per scalar; SO2 / SO3 entries ignored) as lo then hi in the space's scalar type
std::vector<unsigned char> packBounds(const mptg_space_desc* space, const double* lo, const double* hi) {
    const int D = spaceScalars(space);
    std::vector<unsigned char> host(2 * (size_t)D * (size_t)space->scalar);
    for (int c = 0; c < D; ++c) {
        const double l = lo ? lo[c] : 0.0, h = hi ? hi[c] : 0.0;
        if (space->scalar == MPTG_F32) {
            ((float*)host.data())[c] = (float)l;
            ((float*)host.data())[D + c] = (float)h;
        } else {
            ((double*)host.data())[c] = l;
            ((double*)host.data())[D + c] = h;
        }
    }
    return host;
}
// persistent copy (planner handles)
int uploadBounds(mptg_ctx* ctx, const mptg_space_desc* space, const double* lo, const double* hi, void** out) {
    const std::vector<unsigned char> host = packBounds(space, lo, hi);
    MPTG_CUDA(ctx, cudaMalloc(out, host.size()));
    return uploadSync(ctx, *out, host.data(), host.size());
}
// per-call copy in scratch slot 8, stream-ordered (the pageable source is staged before the call returns)
int stageBounds(mptg_ctx* ctx, const mptg_space_desc* space, const double* lo, const double* hi, void** out) {
    const std::vector<unsigned char> host = packBounds(space, lo, hi);
    if (int rc = scratch(ctx, 8, host.size(), out)) return rc;
    MPTG_CUDA(ctx, cudaMemcpyAsync(*out, host.data(), host.size(), cudaMemcpyHostToDevice, ctx->stream));
    return MPTG_OK;
}

bool spaceOk(const mptg_space_desc* space) {
    const int D = space ? spaceScalars(space) : 0;
    return space && D > 0 && D <= MPTG_MAX_SCALARS && (space->scalar == MPTG_F32 || space->scalar == MPTG_F64);
}

template <typename S>
void launchSample(mptg_ctx* ctx, const mptg_space_desc* space, const void* bounds, uint64_t seed, uint64_t first, uint32_t n,
                  const void* goalDev, double goalBias, void* outDev) {
    const int D = spaceScalars(space);
    sampleKernel<S><<<(n + 127) / 128, 128, 0, ctx->stream>>>(makeDevSpace<S>(*space), (const S*)bounds, (const S*)bounds + D, seed, first, n,
                                                               (const S*)goalDev, (S)goalBias, (S*)outDev);
}

}  // namespace

extern "C" {

int mptg_space_uniforms(const mptg_space_desc* space) { return space ? spaceUniforms(space) : 0; }

int mptg_sample_batch_dev(mptg_ctx* ctx, const mptg_space_desc* space, const double* lo, const double* hi, uint64_t seed, uint64_t first,
                          uint32_t n, void* out_dev) {
    if (!ctx || !spaceOk(space) || (n && !out_dev)) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_sample_batch: bad argument");
    if (n == 0) return MPTG_OK;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    void* bounds = nullptr;
    if (int rc = stageBounds(ctx, space, lo, hi, &bounds)) return rc;
    if (space->scalar == MPTG_F32) launchSample<float>(ctx, space, bounds, seed, first, n, nullptr, 0.0, out_dev);
    else launchSample<double>(ctx, space, bounds, seed, first, n, nullptr, 0.0, out_dev);
    MPTG_LAUNCHED(ctx);
    return MPTG_OK;
}

int mptg_sample_batch(mptg_ctx* ctx, const mptg_space_desc* space, const double* lo, const double* hi, uint64_t seed, uint64_t first,
                      uint32_t n, void* out) {
    if (!ctx || !spaceOk(space) || (n && !out)) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_sample_batch: bad argument");
    if (n == 0) return MPTG_OK;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)n * spaceScalars(space) * space->scalar;
    void* dOut;
    if (int rc = scratch(ctx, 1, bytes, &dOut)) return rc;
    if (int rc = mptg_sample_batch_dev(ctx, space, lo, hi, seed, first, n, dOut)) return rc;
    MPTG_CUDA(ctx, cudaMemcpyAsync(out, dOut, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPTG_OK;
}

int mptg_sample_transform_batch(mptg_ctx* ctx, const mptg_space_desc* space, const double* lo, const double* hi, const void* uniforms,
                                uint32_t n, void* out) {
    if (!ctx || !spaceOk(space) || (n && (!out || !uniforms))) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_sample_transform_batch: bad argument");
    if (n == 0) return MPTG_OK;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int D = spaceScalars(space), U = spaceUniforms(space);
    const size_t ub = (size_t)n * U * space->scalar, ob = (size_t)n * D * space->scalar;
    void *dU, *dOut, *bounds = nullptr;
    if (int rc = scratch(ctx, 0, ub, &dU)) return rc;
    if (int rc = scratch(ctx, 1, ob, &dOut)) return rc;
    if (int rc = stageBounds(ctx, space, lo, hi, &bounds)) return rc;
    MPTG_CUDA(ctx, cudaMemcpyAsync(dU, uniforms, ub, cudaMemcpyHostToDevice, ctx->stream));
    if (space->scalar == MPTG_F32)
        transformKernel<float><<<(n + 127) / 128, 128, 0, ctx->stream>>>(makeDevSpace<float>(*space), (const float*)bounds, (const float*)bounds + D,
                                                                        (const float*)dU, U, n, (float*)dOut);
    else
        transformKernel<double><<<(n + 127) / 128, 128, 0, ctx->stream>>>(makeDevSpace<double>(*space), (const double*)bounds,
                                                                         (const double*)bounds + D, (const double*)dU, U, n, (double*)dOut);
    MPTG_LAUNCHED(ctx);
    MPTG_CUDA(ctx, cudaMemcpyAsync(out, dOut, ob, cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPTG_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// device-resident PRRT
// ---------------------------------------------------------------------------------------------
struct mptg_prrt {
    mptg_ctx* ctx = nullptr;
    mptg_geom* geom = nullptr;
    mptg_knn* knn = nullptr;  // owned
    mptg_space_desc space{};
    int D = 0, scalar = MPTG_F32;
    double range = 0, goalBias = 0, goalRadius = 0, linkStep = 0;
    bool hasGoal = false;
    uint64_t seed = 0, drawn = 0;  // samples drawn so far = counter of the next sample
    uint32_t capacity = 0, size = 0, maxWave = 0;
    uint32_t goalNode = MPTG_NO_INDEX;
    uint64_t waves = 0;
    // device
    void* bounds = nullptr;  // lo, hi
    void* goal = nullptr;
    void* nodes = nullptr;   // AoS [capacity][D]
    uint32_t* parent = nullptr;
    void *samples = nullptr, *from = nullptr, *to = nullptr, *fresh = nullptr, *nearDist = nullptr;
    uint32_t *nearIdx = nullptr, *nearCnt = nullptr, *sel = nullptr, *nSel = nullptr, *result = nullptr;
    uint8_t *alive = nullptr, *okValid = nullptr, *okLink = nullptr, *keep = nullptr;
    void* selTemp = nullptr;
    size_t selBytes = 0;
    uint32_t* hostResult = nullptr;  // pinned: [0] nodes added, [1] goal node
};

namespace {

void prrtFree(mptg_prrt* p) {
    if (!p) return;
    if (p->knn) mptg_knn_destroy(p->knn);
    for (void* q : {p->bounds, p->goal, p->nodes, (void*)p->parent, p->samples, p->from, p->to, p->fresh, p->nearDist, (void*)p->nearIdx,
                    (void*)p->nearCnt, (void*)p->sel, (void*)p->nSel, (void*)p->result, (void*)p->alive, (void*)p->okValid, (void*)p->okLink,
                    (void*)p->keep, p->selTemp})
        cudaFree(q);
    if (p->hostResult) cudaFreeHost(p->hostResult);
    delete p;
}

template <typename S>
int prrtWaveT(mptg_prrt* p, uint32_t W) {
    mptg_ctx* ctx = p->ctx;
    const DevSpace<S> sp = makeDevSpace<S>(p->space);
    const uint32_t grid = (W + 127) / 128;
    // sample (prrt.hpp:365-387: goal-biased only until a goal has been reached)
    const bool biased = p->hasGoal && p->goalBias > 0 && p->goalNode == MPTG_NO_INDEX;
    sampleKernel<S><<<grid, 128, 0, ctx->stream>>>(sp, (const S*)p->bounds, (const S*)p->bounds + p->D, p->seed, p->drawn, W,
                                                   biased ? (const S*)p->goal : nullptr, (S)p->goalBias, (S*)p->samples);
    MPTG_LAUNCHED(ctx);
    p->drawn += W;
    // nearest (prrt.hpp:416)
    if (int rc = mptg_knn_query_dev(p->knn, p->samples, W, 1, -1.0, p->nearIdx, p->nearDist, p->nearCnt)) return rc;
    prrtSteerKernel<S><<<grid, 128, 0, ctx->stream>>>(sp, (const S*)p->nodes, (const S*)p->samples, p->nearIdx, (const S*)p->nearDist, p->nearCnt, W,
                                                      (S)p->range, (S*)p->from, (S*)p->to, p->alive);
    MPTG_LAUNCHED(ctx);
    // valid, link (prrt.hpp:439-441)
    if (int rc = mptg_valid_batch_dev(p->geom, p->to, W, p->okValid, nullptr)) return rc;
    if (int rc = mptg_link_batch_dev(p->geom, &p->space, p->from, p->to, W, p->linkStep, p->okLink, nullptr)) return rc;
    if (W <= 16384) {  // young tree, small wave: the wave costs its number of launches -- flags and compaction in one, by one CTA
        starCompactKernel<<<1, 1024, 0, ctx->stream>>>(p->alive, p->okValid, p->okLink, W, p->sel, p->nSel, nullptr);
        MPTG_LAUNCHED(ctx);
    } else {
        prrtFlagKernel<<<grid, 128, 0, ctx->stream>>>(p->alive, p->okValid, p->okLink, W, p->keep);
        MPTG_LAUNCHED(ctx);
        size_t bytes = p->selBytes;
        MPTG_CUDA(ctx, cub::DeviceSelect::Flagged(p->selTemp, bytes, thrust::counting_iterator<uint32_t>(0), p->keep, p->sel, p->nSel, (int)W,
                                                  ctx->stream));
        MPTG_LAUNCHED(ctx);
    }
    prrtAppendKernel<S><<<grid, 128, 0, ctx->stream>>>(sp, p->sel, p->nSel, (const S*)p->to, p->nearIdx, p->size, p->capacity,
                                                       p->hasGoal ? (const S*)p->goal : nullptr, (S)p->goalRadius, (S*)p->nodes, p->parent,
                                                       (S*)p->fresh, p->result);
    MPTG_LAUNCHED(ctx);
    MPTG_CUDA(ctx, cudaMemcpyAsync(p->hostResult, p->result, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const uint32_t added = p->hostResult[0];
    if (p->hostResult[1] != MPTG_NO_INDEX && p->goalNode == MPTG_NO_INDEX) p->goalNode = p->hostResult[1];
    if (added) {
        uint32_t first = 0;
        if (int rc = mptg_knn_insert_dev(p->knn, p->fresh, added, &first)) return rc;  // prrt.hpp:447
        if (first != p->size) return fail(ctx, MPTG_ERR_CUDA, "mptg_prrt_wave: node numbering out of step");
        p->size += added;
    }
    ++p->waves;
    return MPTG_OK;
}

}  // namespace

extern "C" {

int mptg_prrt_create(mptg_ctx* ctx, mptg_geom* geom, const mptg_prrt_params* prm, mptg_prrt** out) {
    if (!ctx || !geom || !prm || !out || !spaceOk(prm->space) || prm->capacity == 0 || prm->max_wave == 0 || !(prm->range > 0))
        return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_prrt_create: bad argument");
    if (geom->ctx != ctx || geom->scalar != prm->space->scalar || geom->D != spaceScalars(prm->space))
        return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_prrt_create: the geometry's states are not states of this space");
    if (geom->kind == MPTG_GEOM_MESH && !(prm->link_step > 0)) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_prrt_create: mesh geometries need link_step > 0");
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    auto* p = new mptg_prrt();
    p->ctx = ctx, p->geom = geom, p->space = *prm->space;
    p->D = spaceScalars(prm->space), p->scalar = prm->space->scalar;
    p->range = prm->range, p->goalBias = prm->goal_bias, p->goalRadius = prm->goal_radius, p->linkStep = prm->link_step;
    p->seed = prm->seed, p->capacity = prm->capacity, p->maxWave = prm->max_wave, p->hasGoal = prm->goal_state != nullptr;
    const size_t sb = (size_t)p->D * p->scalar, W = p->maxWave;
    int rc = mptg_knn_create(ctx, prm->space, p->capacity, &p->knn);
    if (!rc) rc = uploadBounds(ctx, prm->space, prm->lo, prm->hi, &p->bounds);
    auto alloc = [&](auto** q, size_t bytes) {
        if (rc) return;
        cudaError_t e = cudaMalloc((void**)q, bytes ? bytes : 16);
        if (e != cudaSuccess) rc = fail(ctx, MPTG_ERR_OOM, "mptg_prrt_create: %s", cudaGetErrorString(e));
    };
    alloc(&p->nodes, (size_t)p->capacity * sb);
    alloc(&p->parent, (size_t)p->capacity * sizeof(uint32_t));
    alloc(&p->samples, W * sb), alloc(&p->from, W * sb), alloc(&p->to, W * sb), alloc(&p->fresh, W * sb);
    alloc(&p->nearDist, W * p->scalar), alloc(&p->nearIdx, W * 4), alloc(&p->nearCnt, W * 4), alloc(&p->sel, W * 4);
    alloc(&p->nSel, 4), alloc(&p->result, 8), alloc(&p->alive, W), alloc(&p->okValid, W), alloc(&p->okLink, W), alloc(&p->keep, W);
    if (!rc) {
        cub::DeviceSelect::Flagged(nullptr, p->selBytes, thrust::counting_iterator<uint32_t>(0), p->keep, p->sel, p->nSel, (int)W);
        alloc(&p->selTemp, p->selBytes);
    }
    if (!rc && p->hasGoal) {
        alloc(&p->goal, sb);
        if (!rc) rc = uploadSync(ctx, p->goal, prm->goal_state, sb);
    }
    if (!rc && cudaMallocHost((void**)&p->hostResult, 2 * sizeof(uint32_t)) != cudaSuccess) rc = fail(ctx, MPTG_ERR_OOM, "mptg_prrt_create: pinned allocation failed");
    if (rc) {
        prrtFree(p);
        return rc;
    }
    *out = p;
    return MPTG_OK;
}

int mptg_prrt_destroy(mptg_prrt* p) {
    if (!p) return MPTG_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    prrtFree(p);
    return MPTG_OK;
}

int mptg_prrt_add_start(mptg_prrt* p, const void* state) {
    if (!p || !state) return fail(p ? p->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_prrt_add_start: bad argument");
    if (p->size >= p->capacity) return fail(p->ctx, MPTG_ERR_CAPACITY, "mptg_prrt_add_start: tree is full");
    mptg_ctx* ctx = p->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t sb = (size_t)p->D * p->scalar;
    const uint32_t none = MPTG_NO_INDEX;
    if (int rc = uploadSync(ctx, (char*)p->nodes + (size_t)p->size * sb, state, sb)) return rc;
    if (int rc = uploadSync(ctx, p->parent + p->size, &none, sizeof none)) return rc;
    uint32_t first = 0;
    if (int rc = mptg_knn_insert(p->knn, state, 1, &first)) return rc;
    ++p->size;
    return MPTG_OK;
}

int mptg_prrt_wave(mptg_prrt* p, uint32_t n_samples, uint32_t* size_out, uint32_t* goal_node_out) {
    if (!p || n_samples == 0 || n_samples > p->maxWave) return fail(p ? p->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_prrt_wave: bad argument");
    if (p->size == 0) return fail(p->ctx, MPTG_ERR_BAD_ARG, "mptg_prrt_wave: there are no valid initial states");  // prrt.hpp:197-198
    MPTG_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
    // result = {0 nodes added, no goal node}: stream-ordered, no host round trip in front of the wave
    MPTG_CUDA(p->ctx, cudaMemsetAsync(p->result, 0, 4, p->ctx->stream));
    MPTG_CUDA(p->ctx, cudaMemsetAsync(p->result + 1, 0xFF, 4, p->ctx->stream));
    const int rc = p->scalar == MPTG_F32 ? prrtWaveT<float>(p, n_samples) : prrtWaveT<double>(p, n_samples);
    if (size_out) *size_out = p->size;
    if (goal_node_out) *goal_node_out = p->goalNode;
    return rc;
}

uint32_t mptg_prrt_size(const mptg_prrt* p) { return p ? p->size : 0; }
uint64_t mptg_prrt_samples_drawn(const mptg_prrt* p) { return p ? p->drawn : 0; }

int mptg_prrt_get_tree(mptg_prrt* p, uint32_t first, uint32_t count, void* states_out, uint32_t* parents_out) {
    if (!p || (uint64_t)first + count > p->size) return fail(p ? p->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_prrt_get_tree: bad range");
    if (count == 0) return MPTG_OK;
    mptg_ctx* ctx = p->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t sb = (size_t)p->D * p->scalar;
    if (states_out) MPTG_CUDA(ctx, cudaMemcpyAsync(states_out, (char*)p->nodes + first * sb, count * sb, cudaMemcpyDeviceToHost, ctx->stream));
    if (parents_out) MPTG_CUDA(ctx, cudaMemcpyAsync(parents_out, p->parent + first, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPTG_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// device-resident PPRM (src/mpt/impl/pprm/pprm.hpp:298-362)
// ---------------------------------------------------------------------------------------------
namespace mptg {

template <typename S>
__global__ void pprmGatherKernel(const uint32_t* __restrict__ sel, uint32_t n, int D, const S* __restrict__ src, S* __restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const S* ps = src + (size_t)sel[i] * D;
    S* pd = dst + (size_t)i * D;
    for (int c = 0; c < D; ++c) pd[c] = ps[c];
}

// pprm.hpp:306-308: a sample closer than epsilon to its nearest node is dropped
template <typename S>
__global__ void pprmKeepKernel(const S* __restrict__ dist, const uint32_t* __restrict__ cnt, uint32_t k, uint32_t n, S minDist,
                               uint8_t* __restrict__ keep) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keep[i] = (cnt != nullptr && cnt[i] > 0 && dist[(size_t)i * k] < minDist) ? 0 : 1;
}

// edges (kept sample sel[s]) -> (its neighbour j), pprm.hpp:325; slots beyond the neighbour count become zero-length
// edges at the sample and are ignored later
template <typename S>
__global__ void pprmEdgeKernel(const uint32_t* __restrict__ sel, uint32_t nSel, uint32_t k, int D, const S* __restrict__ samples,
                               const S* __restrict__ nodes, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ cnt,
                               S* __restrict__ from, S* __restrict__ to) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)nSel * k) return;
    const uint32_t s = (uint32_t)(e / k), j = (uint32_t)(e % k);
    const uint32_t i = sel[s];
    const S* ps = samples + (size_t)i * D;
    const S* pn = j < cnt[i] ? nodes + (size_t)idx[(size_t)i * k + j] * D : ps;
    S* pf = from + e * D;
    S* pt = to + e * D;
    for (int c = 0; c < D; ++c) pf[c] = ps[c], pt[c] = pn[c];
}

// append the kept samples in sample order: state, marks (goal test, pprm.hpp:312-316), own component, edge row
template <typename S>
__global__ void pprmAppendKernel(DevSpace<S> sp, const uint32_t* __restrict__ sel, uint32_t nSel, uint32_t k, uint32_t stride,
                                 const S* __restrict__ samples, const uint32_t* __restrict__ idx, const S* __restrict__ dist,
                                 const uint32_t* __restrict__ cnt, const uint8_t* __restrict__ okEdge, uint32_t size, const S* __restrict__ goal,
                                 S goalRadius, uint32_t forcedMarks, S* __restrict__ nodes, S* __restrict__ fresh, uint32_t* __restrict__ edgeIdx,
                                 S* __restrict__ edgeDist, uint8_t* __restrict__ marks, uint32_t* __restrict__ comp, uint32_t* __restrict__ lists) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nSel) return;
    const int D = sp.D;
    const uint32_t i = sel[s], id = size + s;
    const S* ps = samples + (size_t)i * D;
    S* pn = nodes + (size_t)id * D;
    S* pf = fresh + (size_t)s * D;
    for (int c = 0; c < D; ++c) pn[c] = pf[c] = ps[c];
    uint32_t m = forcedMarks;
    if (!(m & MPTG_PPRM_GOAL) && goal != nullptr) {
        const S d = dev::distance<S>(sp, [&](int c) { return ps[c]; }, [&](int c) { return goal[c]; });
        if (d <= goalRadius) m |= MPTG_PPRM_GOAL;
    }
    marks[id] = (uint8_t)m;
    comp[id] = id;
    // lists: [0] start count, [1] goal count, [2 .. 2+PPRM_STARTS) starts, then goals
    if (m & MPTG_PPRM_START) {
        const uint32_t slot = atomicAdd(lists + 0, 1u);
        if (slot < 64u) lists[2 + slot] = id;
    }
    if (m & MPTG_PPRM_GOAL) {
        const uint32_t slot = atomicAdd(lists + 1, 1u);
        if (slot < 4096u) lists[2 + 64 + slot] = id;
    }
    const uint32_t have = cnt ? cnt[i] : 0u;
    for (uint32_t j = 0; j < stride; ++j) {
        const bool live = j < k && j < have && okEdge[(size_t)s * k + j] != 0;
        edgeIdx[(size_t)id * stride + j] = live ? idx[(size_t)i * k + j] : MPTG_NO_INDEX;
        edgeDist[(size_t)id * stride + j] = live ? dist[(size_t)i * k + j] : S(0);
    }
}

// lock-free union-find: roots point to themselves, a root is only ever hooked under a smaller index (no cycles),
// path halving writes are benign (they replace a parent by one of its ancestors)
__device__ __forceinline__ uint32_t pprmFind(uint32_t* comp, uint32_t x) {
    for (;;) {
        const uint32_t p = ((volatile uint32_t*)comp)[x];
        if (p == x) return x;
        const uint32_t gp = ((volatile uint32_t*)comp)[p];
        if (gp != p) ((volatile uint32_t*)comp)[x] = gp;
        x = p;
    }
}
__device__ __forceinline__ void pprmUnite(uint32_t* comp, uint32_t a, uint32_t b) {
    for (;;) {
        a = pprmFind(comp, a);
        b = pprmFind(comp, b);
        if (a == b) return;
        if (a < b) {
            const uint32_t t = a;
            a = b;
            b = t;
        }
        if (atomicCAS(comp + a, a, b) == a) return;
    }
}
// merge the components along the new edges (pprm.hpp:327-334, 341-362)
__global__ void pprmUniteKernel(uint32_t nSel, uint32_t stride, uint32_t size, const uint32_t* __restrict__ edgeIdx, uint32_t* comp) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)nSel * stride) return;
    const uint32_t id = size + (uint32_t)(e / stride);
    const uint32_t nb = edgeIdx[(size_t)id * stride + (e % stride)];
    if (nb != MPTG_NO_INDEX) pprmUnite(comp, id, nb);
}
// component.hpp:97-99: solved when one component holds a start and a goal
__global__ void pprmSolvedKernel(uint32_t* comp, const uint32_t* __restrict__ lists, uint32_t* __restrict__ result) {
    const uint32_t nStart = min(lists[0], 64u), nGoal = min(lists[1], 4096u);
    for (uint32_t p = threadIdx.x; p < nStart * nGoal; p += blockDim.x) {
        const uint32_t s = lists[2 + p / nGoal], g = lists[2 + 64 + p % nGoal];
        if (pprmFind(comp, s) == pprmFind(comp, g)) result[0] = 1u;
    }
}
// representatives for mptg_pprm_get_graph
__global__ void pprmRootKernel(uint32_t* comp, uint32_t first, uint32_t count, uint32_t* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = pprmFind(comp, first + i);
}

// ---------------------------------------------------------------------------------------------
// PPRM-IRS on the device (src/mpt/impl/pprm_irs/pprm_irs.hpp:350-368, shortest_path_check.hpp:111-225): of the validated
// (sample, neighbour) edges of a wave, keep as SPARSE edges those whose ends the sparse roadmap does not already join by a
// path shorter than stretch x the edge.  One warp per new node runs the reference's bounded Dijkstra search over the
// roadmap as it stood when the wave began plus the node's own kept edges: neighbours nearest first, the search resumed from
// check to check (the bound only grows), a kept edge entering it at once, stale queue entries skipped.  Path costs are the
// left-to-right sums from the new node outwards, as Dijkstra forms them; by monotonicity of the rounded addition every
// exact search yields the same labels, so the kept set does not depend on the order ties are settled in.
// Working storage per node: an open-addressing table node -> best cost and a binary min-heap with lazy deletion, both
// in global memory (cap entries / 4 cap pushes; exceeding either raises the error flag -> MPTG_ERR_CAPACITY).
// Roadmap adjacency = the node's own row (edges to older nodes) + the reverse list through revHead / revNext (edges from
// newer nodes; entry id = newer node * stride + slot, so an entry's neighbour is id / stride and its length edgeDist[id]).
template <typename S>
struct SpannerScratch {
    uint32_t* tabNode;  // [wave][2 cap], MPTG_NO_INDEX = empty (memset before the launch)
    S* tabCost;         // [wave][2 cap]
    S* heapCost;        // [wave][4 cap]
    uint32_t* heapNode; // [wave][4 cap]
    uint32_t cap;
};

constexpr int SPANNER_WARPS = 4;

// One WARP per new node.  Lane 0 owns the heap; the 32 lanes read a settled node's edge row together (one coalesced load
// per 32 slots), probe the table for their neighbours in parallel (distinct nodes: a row lists a node once, and the reverse
// list holds edges from NEWER nodes only) and hand their improvements to lane 0 for queueing.  Same labels as a sequential
// search (each relaxation touches its own node), so the same kept set.
template <typename S>
__global__ void __launch_bounds__(SPANNER_WARPS * 32) pprmSpannerKernel(
    uint32_t nSel, uint32_t k, uint32_t stride, const uint32_t* __restrict__ sel, const uint32_t* __restrict__ nnIdx, const S* __restrict__ nnDist,
    const uint32_t* __restrict__ nnCnt, uint8_t* __restrict__ okEdge, const uint32_t* __restrict__ edgeIdx, const S* __restrict__ edgeDist,
    const uint32_t* __restrict__ revHead, const uint32_t* __restrict__ revNext, S stretch, SpannerScratch<S> w, uint32_t* __restrict__ err) {
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= nSel) return;
    const uint32_t i = sel[s];
    const uint32_t cnt = nnCnt[i] < k ? nnCnt[i] : k;
    const uint32_t slots = 2u * w.cap, heapCap = 4u * w.cap;
    uint32_t* tn = w.tabNode + (size_t)s * slots;
    S* tc = w.tabCost + (size_t)s * slots;
    S* hc = w.heapCost + (size_t)s * heapCap;
    uint32_t* hn = w.heapNode + (size_t)s * heapCap;
    uint32_t used = 0;   // labelled nodes (the same value in every lane)
    uint32_t heapN = 0;  // the same value in every lane; the heap itself is touched by lane 0 only
    bool overflow = false;
    // slot of `node` for reading: the slot holding it or an empty one
    auto slotOf = [&](uint32_t node) {
        uint32_t h = (node * 2654435761u) & (slots - 1u);
        for (;;) {
            const uint32_t t = *(volatile uint32_t*)(tn + h);
            if (t == MPTG_NO_INDEX || t == node) return h;
            h = (h + 1u) & (slots - 1u);
        }
    };
    // slot of `node`, claiming an empty one (lanes insert different nodes at the same time); fresh = it was not there before
    auto claim = [&](uint32_t node, bool& fresh) {
        uint32_t h = (node * 2654435761u) & (slots - 1u);
        for (;;) {
            const uint32_t t = atomicCAS(tn + h, MPTG_NO_INDEX, node);
            if (t == MPTG_NO_INDEX) {
                fresh = true;
                return h;
            }
            if (t == node) {
                fresh = false;
                return h;
            }
            h = (h + 1u) & (slots - 1u);
        }
    };
    auto push0 = [&](S c, uint32_t node) {  // lane 0
        uint32_t x = heapN;
        while (x > 0) {
            const uint32_t parent = (x - 1u) >> 1;
            if (!(c < hc[parent])) break;
            hc[x] = hc[parent], hn[x] = hn[parent];
            x = parent;
        }
        hc[x] = c, hn[x] = node;
    };
    auto pop0 = [&]() {  // lane 0; heapN already decremented
        const S c = hc[heapN];
        const uint32_t node = hn[heapN];
        uint32_t x = 0;
        for (;;) {
            uint32_t child = 2u * x + 1u;
            if (child >= heapN) break;
            if (child + 1u < heapN && hc[child + 1u] < hc[child]) ++child;
            if (!(hc[child] < c)) break;
            hc[x] = hc[child], hn[x] = hn[child];
            x = child;
        }
        if (heapN > 0) hc[x] = c, hn[x] = node;
    };
    // queue the improvements of the lanes in `mask` (their labels are already in the table)
    auto queue = [&](unsigned mask, S c, uint32_t node) {
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1u;
            const S cc = __shfl_sync(FULL, c, src);
            const uint32_t nn = __shfl_sync(FULL, node, src);
            if (heapN >= heapCap) {
                overflow = true;
                return;
            }
            if (lane == 0) push0(cc, nn);
            ++heapN;
        }
        __syncwarp();
    };
    for (uint32_t j = 0; j < cnt && !overflow; ++j) {
        uint8_t* flag = okEdge + (size_t)s * k + j;
        if (!*flag) continue;
        const uint32_t v = nnIdx[(size_t)i * k + j];
        const S d = nnDist[(size_t)i * k + j];
        const S target = stretch * d;  // pprm_irs.hpp:351
        {
            const uint32_t h = slotOf(v);
            if (*(volatile uint32_t*)(tn + h) == v && *(volatile S*)(tc + h) < target) {  // shortest_path_check.hpp:133-140
                __syncwarp();
                if (lane == 0) *flag = 0;
                continue;
            }
        }
        bool found = false;
        while (heapN > 0 && !overflow) {
            S priority = S(0);
            uint32_t top = 0;
            if (lane == 0) priority = hc[0], top = hn[0];
            priority = __shfl_sync(FULL, priority, 0);
            top = __shfl_sync(FULL, top, 0);
            const S pathCost = *(volatile S*)(tc + slotOf(top));
            if (pathCost >= target) break;  // :159-160
            --heapN;
            if (lane == 0) pop0();
            __syncwarp();
            if (pathCost != priority) continue;  // stale: settled through a shorter path (:166-171)
            found = top == v;
            // the node's own row: edges to older nodes, 32 slots at a time
            for (uint32_t base = 0; base < stride && !overflow; base += 32) {
                const uint32_t t = base + lane;
                uint32_t nbr = MPTG_NO_INDEX;
                S c = S(0);
                if (t < stride) {
                    nbr = edgeIdx[(size_t)top * stride + t];
                    if (nbr != MPTG_NO_INDEX) c = pathCost + edgeDist[(size_t)top * stride + t];
                }
                const bool live = nbr != MPTG_NO_INDEX;
                found = found || __any_sync(FULL, live && nbr == v && c < target);
                bool fresh = false, better = false;
                if (live) {
                    const uint32_t h = claim(nbr, fresh);
                    better = fresh || c < *(volatile S*)(tc + h);
                    if (better) tc[h] = c;
                }
                used += __popc(__ballot_sync(FULL, fresh));
                if (used > w.cap) overflow = true;
                __syncwarp();
                queue(__ballot_sync(FULL, better), c, nbr);
            }
            // edges from newer nodes: the reverse list, gathered 32 entries at a time (every lane walks the same chain)
            uint32_t e = revHead[top];
            while (e != MPTG_NO_INDEX && !overflow) {
                uint32_t nbr = MPTG_NO_INDEX;
                S c = S(0);
                for (int got = 0; got < 32 && e != MPTG_NO_INDEX; ++got) {
                    if (got == lane) nbr = e / stride, c = pathCost + edgeDist[e];
                    e = revNext[e];
                }
                const bool live = nbr != MPTG_NO_INDEX;
                found = found || __any_sync(FULL, live && nbr == v && c < target);
                bool fresh = false, better = false;
                if (live) {
                    const uint32_t h = claim(nbr, fresh);
                    better = fresh || c < *(volatile S*)(tc + h);
                    if (better) tc[h] = c;
                }
                used += __popc(__ballot_sync(FULL, fresh));
                if (used > w.cap) overflow = true;
                __syncwarp();
                queue(__ballot_sync(FULL, better), c, nbr);
            }
            if (found) break;  // :208-209
        }
        if (overflow) break;
        if (found) {
            if (lane == 0) *flag = 0;
            continue;
        }
        // a sparse edge: part of the search from here on (:219-222)
        bool fresh = false;
        uint32_t h = 0;
        if (lane == 0) {
            h = claim(v, fresh);
            tc[h] = d;
        }
        fresh = __shfl_sync(FULL, (int)fresh, 0) != 0;
        used += fresh ? 1u : 0u;
        if (used > w.cap) overflow = true;
        __syncwarp();
        queue(1u, d, v);
    }
    if (overflow && lane == 0) atomicOr(err, 1u);
}

// reverse lists for the sparse edges of the nodes just appended
__global__ void pprmSpannerLinkKernel(uint32_t nSel, uint32_t stride, uint32_t size, const uint32_t* __restrict__ edgeIdx, uint32_t* revHead,
                                      uint32_t* __restrict__ revNext) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)nSel * stride) return;
    const uint32_t id = size + (uint32_t)(e / stride);
    const uint32_t entry = id * stride + (uint32_t)(e % stride);
    const uint32_t nbr = edgeIdx[entry];
    if (nbr != MPTG_NO_INDEX) revNext[entry] = atomicExch(revHead + nbr, entry);
}

}  // namespace mptg

struct mptg_pprm {
    mptg_ctx* ctx = nullptr;
    mptg_geom* geom = nullptr;
    mptg_knn* knn = nullptr;  // owned
    mptg_space_desc space{};
    int D = 0, scalar = MPTG_F32, dims = 0;
    double goalRadius = 0, linkStep = 0;
    bool hasGoal = false, solved = false;
    uint64_t seed = 0, drawn = 0, waves = 0;
    uint32_t capacity = 0, size = 0, maxWave = 0, stride = 0;
    // device
    void *bounds = nullptr, *goal = nullptr, *nodes = nullptr, *edgeDist = nullptr;
    uint32_t *edgeIdx = nullptr, *comp = nullptr, *lists = nullptr;
    uint8_t* marks = nullptr;
    void *samples = nullptr, *cand = nullptr, *fresh = nullptr, *nnDist = nullptr, *from = nullptr, *to = nullptr;
    uint32_t *nnIdx = nullptr, *nnCnt = nullptr, *sel = nullptr, *sel2 = nullptr, *nSel = nullptr, *result = nullptr;
    uint8_t *okValid = nullptr, *keep = nullptr, *okEdge = nullptr;
    void* selTemp = nullptr;
    size_t selBytes = 0;
    uint32_t* host = nullptr;  // pinned: [0] selected count, [1] solved, [2] spanner error flag
    // PPRM-IRS (mptg_pprm_set_spanner): stretch > 0 switches it on
    double stretch = 0;
    uint32_t *revHead = nullptr, *revNext = nullptr, *spanErr = nullptr, spanCap = 0;
    uint8_t* okEdgeRaw = nullptr;  // the edge batch's answers before the spanner stage (a search that outgrows its storage is rerun)
    void *spanTabCost = nullptr, *spanHeapCost = nullptr;
    uint32_t *spanTabNode = nullptr, *spanHeapNode = nullptr;
};

namespace {

void pprmFree(mptg_pprm* p) {
    if (!p) return;
    if (p->knn) mptg_knn_destroy(p->knn);
    for (void* q : {p->bounds, p->goal, p->nodes, p->edgeDist, (void*)p->edgeIdx, (void*)p->comp, (void*)p->lists, (void*)p->marks, p->samples,
                    p->cand, p->fresh, p->nnDist, p->from, p->to, (void*)p->nnIdx, (void*)p->nnCnt, (void*)p->sel, (void*)p->sel2, (void*)p->nSel,
                    (void*)p->result, (void*)p->okValid, (void*)p->keep, (void*)p->okEdge, p->selTemp, (void*)p->revHead, (void*)p->revNext,
                    (void*)p->spanErr, p->spanTabCost, p->spanHeapCost, (void*)p->spanTabNode, (void*)p->spanHeapNode, (void*)p->okEdgeRaw})
        cudaFree(q);
    if (p->host) cudaFreeHost(p->host);
    delete p;
}

int spaceDimensions(const mptg_space_desc& sp) {  // Space::dimensions(): SO(3) counts 3
    int d = 0;
    for (int i = 0; i < sp.n_parts; ++i) d += sp.part[i].kind == MPTG_PART_SO3 ? 3 : sp.part[i].dim;
    return d;
}

// k = ceil(kRRG * ln(n + 1)) in the space's scalar type (pprm.hpp:146,302-303)
template <typename S>
uint32_t pprmK(int dims, uint32_t n) {
    const S e = (S)2.718281828459045235360287471352662498L;
    const S kRRG = e + e / (S)dims;
    const S logSizePlus1 = (S)std::log((double)n + 1.0);
    const int k = (int)std::ceil(kRRG * logSizePlus1);
    return (uint32_t)(k < 1 ? 1 : k);
}

template <typename T>
int selectFlagged(mptg_pprm* p, const uint8_t* flags, uint32_t n, T* out, uint32_t* countHost) {
    mptg_ctx* ctx = p->ctx;
    size_t bytes = p->selBytes;
    MPTG_CUDA(ctx, cub::DeviceSelect::Flagged(p->selTemp, bytes, thrust::counting_iterator<uint32_t>(0), flags, out, p->nSel, (int)n, ctx->stream));
    MPTG_LAUNCHED(ctx);
    MPTG_CUDA(ctx, cudaMemcpyAsync(p->host, p->nSel, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *countHost = p->host[0];
    return MPTG_OK;
}

// (re)allocate the spanner's search storage for `cap` labelled nodes per new node (rounded up to a power of two)
int spannerScratch(mptg_pprm* p, uint32_t cap) {
    mptg_ctx* ctx = p->ctx;
    uint32_t pow2 = 64;
    while (pow2 < cap) pow2 <<= 1;
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (void** q : {(void**)&p->spanTabNode, &p->spanTabCost, (void**)&p->spanHeapNode, &p->spanHeapCost}) {
        if (*q) cudaFree(*q);
        *q = nullptr;
    }
    const size_t W = p->maxWave;
    const size_t bytes[4] = {W * 2 * pow2 * 4, W * 2 * pow2 * (size_t)p->scalar, W * 4 * pow2 * 4, W * 4 * pow2 * (size_t)p->scalar};
    void** dst[4] = {(void**)&p->spanTabNode, &p->spanTabCost, (void**)&p->spanHeapNode, &p->spanHeapCost};
    for (int i = 0; i < 4; ++i) {
        cudaError_t e = cudaMalloc(dst[i], bytes[i]);
        if (e != cudaSuccess) return fail(ctx, MPTG_ERR_OOM, "PPRM-IRS search storage (%u nodes per new node, %zu MiB): %s", pow2, (bytes[0] + bytes[1] + bytes[2] + bytes[3]) >> 20,
                                          cudaGetErrorString(e));
    }
    p->spanCap = pow2;
    return MPTG_OK;
}
// four times the storage, at most one entry per node the roadmap can hold and 32 GiB in all
int spannerGrow(mptg_pprm* p) {
    uint32_t limit = 64;
    while (limit < p->capacity) limit <<= 1;
    const size_t per = 6 * (4 + (size_t)p->scalar);
    while ((size_t)limit * per * p->maxWave > ((size_t)32 << 30) && limit > 64) limit >>= 1;
    if (p->spanCap >= limit)
        return fail(p->ctx, MPTG_ERR_CAPACITY, "mptg_pprm: a spanner search labelled more than %u nodes (the limit for waves of %u samples); use smaller waves", p->spanCap, p->maxWave);
    const uint32_t next = p->spanCap > limit / 4 ? limit : p->spanCap * 4;
    return spannerScratch(p, next);
}

// Worker::addSample for the W states in p->samples (pprm.hpp:298-339); marks: start / goal marks forced on them
template <typename S>
int pprmProcessT(mptg_pprm* p, uint32_t W, uint32_t marks, uint32_t* firstOut, uint32_t* addedOut) {
    mptg_ctx* ctx = p->ctx;
    const DevSpace<S> sp = makeDevSpace<S>(p->space);
    const int D = p->D;
    cudaStream_t st = ctx->stream;
    *firstOut = p->size, *addedOut = 0;
    // valid (:299), keep the valid samples in sample order
    if (int rc = mptg_valid_batch_dev(p->geom, p->samples, W, p->okValid, nullptr)) return rc;
    uint32_t V = 0;
    if (int rc = selectFlagged(p, p->okValid, W, p->sel, &V)) return rc;
    if (V == 0) return MPTG_OK;
    pprmGatherKernel<S><<<(V + 127) / 128, 128, 0, st>>>(p->sel, V, D, (const S*)p->samples, (S*)p->cand);
    MPTG_LAUNCHED(ctx);
    // k nearest (:302-304) and the epsilon test (:306-308)
    const uint32_t n = p->size;
    uint32_t k = 1;
    uint32_t nSel = V;
    if (n > 0) {
        k = pprmK<S>(p->dims, n);
        if (k > p->stride) k = p->stride;
        if (int rc = mptg_knn_query_dev(p->knn, p->cand, V, k, -1.0, p->nnIdx, p->nnDist, p->nnCnt)) return rc;
        pprmKeepKernel<S><<<(V + 127) / 128, 128, 0, st>>>((const S*)p->nnDist, p->nnCnt, k, V, std::numeric_limits<S>::epsilon(), p->keep);
        MPTG_LAUNCHED(ctx);
        if (int rc = selectFlagged(p, p->keep, V, p->sel2, &nSel)) return rc;
    } else {
        // empty roadmap: nothing to be near to; the samples of this call do not see each other
        std::vector<uint32_t> iota(V);
        for (uint32_t i = 0; i < V; ++i) iota[i] = i;
        MPTG_CUDA(ctx, cudaMemcpyAsync(p->sel2, iota.data(), (size_t)V * 4, cudaMemcpyHostToDevice, st));
        MPTG_CUDA(ctx, cudaMemsetAsync(p->nnCnt, 0, (size_t)V * 4, st));
        MPTG_CUDA(ctx, cudaStreamSynchronize(st));
    }
    if (nSel > p->capacity - p->size) nSel = p->capacity - p->size;
    if (nSel == 0) return MPTG_OK;
    // every (sample, neighbour) edge in one batch (:325)
    if (n > 0) {
        const size_t E = (size_t)nSel * k;
        pprmEdgeKernel<S><<<(unsigned)((E + 127) / 128), 128, 0, st>>>(p->sel2, nSel, k, D, (const S*)p->cand, (const S*)p->nodes, p->nnIdx, p->nnCnt,
                                                                         (S*)p->from, (S*)p->to);
        MPTG_LAUNCHED(ctx);
        if (int rc = mptg_link_batch_dev(p->geom, &p->space, p->from, p->to, (uint32_t)E, p->linkStep, p->okEdge, nullptr)) return rc;
        if (p->stretch > 0) {  // PPRM-IRS: validated edges the spanner does not need are dropped before the rows are written
            MPTG_CUDA(ctx, cudaMemcpyAsync(p->okEdgeRaw, p->okEdge, E, cudaMemcpyDeviceToDevice, st));
            for (;;) {
                const size_t slots = 2 * (size_t)p->spanCap;
                MPTG_CUDA(ctx, cudaMemsetAsync(p->spanTabNode, 0xFF, (size_t)nSel * slots * sizeof(uint32_t), st));
                SpannerScratch<S> ws{p->spanTabNode, (S*)p->spanTabCost, (S*)p->spanHeapCost, p->spanHeapNode, p->spanCap};
                pprmSpannerKernel<S><<<(nSel + SPANNER_WARPS - 1) / SPANNER_WARPS, SPANNER_WARPS * 32, 0, st>>>(nSel, k, p->stride, p->sel2, p->nnIdx, (const S*)p->nnDist, p->nnCnt, p->okEdge,
                                                                     p->edgeIdx, (const S*)p->edgeDist, p->revHead, p->revNext, (S)p->stretch, ws, p->spanErr);
                MPTG_LAUNCHED(ctx);
                MPTG_CUDA(ctx, cudaMemcpyAsync(p->host + 2, p->spanErr, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
                MPTG_CUDA(ctx, cudaStreamSynchronize(st));
                if (!p->host[2]) break;
                // some search outgrew its table or heap: nothing has been written to the roadmap yet -- more storage, same wave again
                if (int rc = spannerGrow(p)) return rc;
                MPTG_CUDA(ctx, cudaMemcpyAsync(p->okEdge, p->okEdgeRaw, E, cudaMemcpyDeviceToDevice, st));
                MPTG_CUDA(ctx, cudaMemsetAsync(p->spanErr, 0, sizeof(uint32_t), st));
            }
        }
    }
    pprmAppendKernel<S><<<(nSel + 127) / 128, 128, 0, st>>>(sp, p->sel2, nSel, k, p->stride, (const S*)p->cand, p->nnIdx, (const S*)p->nnDist, p->nnCnt,
                                                          p->okEdge, p->size, p->hasGoal ? (const S*)p->goal : nullptr, (S)p->goalRadius, marks,
                                                          (S*)p->nodes, (S*)p->fresh, p->edgeIdx, (S*)p->edgeDist, p->marks, p->comp, p->lists);
    MPTG_LAUNCHED(ctx);
    if (n > 0) {
        const size_t E = (size_t)nSel * p->stride;
        pprmUniteKernel<<<(unsigned)((E + 127) / 128), 128, 0, st>>>(nSel, p->stride, p->size, p->edgeIdx, p->comp);
        MPTG_LAUNCHED(ctx);
        if (p->stretch > 0) {
            pprmSpannerLinkKernel<<<(unsigned)((E + 127) / 128), 128, 0, st>>>(nSel, p->stride, p->size, p->edgeIdx, p->revHead, p->revNext);
            MPTG_LAUNCHED(ctx);
        }
    }
    pprmSolvedKernel<<<1, 256, 0, st>>>(p->comp, p->lists, p->result);
    MPTG_LAUNCHED(ctx);
    MPTG_CUDA(ctx, cudaMemcpyAsync(p->host + 1, p->result, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    uint32_t first = 0;
    if (int rc = mptg_knn_insert_dev(p->knn, p->fresh, nSel, &first)) return rc;  // :337
    if (first != p->size) return fail(ctx, MPTG_ERR_CUDA, "mptg_pprm: node numbering out of step");
    MPTG_CUDA(ctx, cudaStreamSynchronize(st));
    if (p->host[1]) p->solved = true;
    p->size += nSel;
    *addedOut = nSel;
    return MPTG_OK;
}

int pprmProcess(mptg_pprm* p, uint32_t W, uint32_t marks, uint32_t* first, uint32_t* added) {
    return p->scalar == MPTG_F32 ? pprmProcessT<float>(p, W, marks, first, added) : pprmProcessT<double>(p, W, marks, first, added);
}

}  // namespace

extern "C" {

int mptg_pprm_create(mptg_ctx* ctx, mptg_geom* geom, const mptg_pprm_params* prm, mptg_pprm** out) {
    if (!ctx || !geom || !prm || !out || !spaceOk(prm->space) || prm->capacity == 0 || prm->max_wave == 0 || prm->max_k > MPTG_MAX_K)
        return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_pprm_create: bad argument");
    if (geom->ctx != ctx || geom->scalar != prm->space->scalar || geom->D != spaceScalars(prm->space))
        return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_pprm_create: the geometry's states are not states of this space");
    if (geom->kind == MPTG_GEOM_MESH && !(prm->link_step > 0)) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_pprm_create: mesh geometries need link_step > 0");
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    auto* p = new mptg_pprm();
    p->ctx = ctx, p->geom = geom, p->space = *prm->space;
    p->D = spaceScalars(prm->space), p->scalar = prm->space->scalar, p->dims = spaceDimensions(*prm->space);
    p->goalRadius = prm->goal_radius, p->linkStep = prm->link_step, p->hasGoal = prm->goal_state != nullptr;
    p->seed = prm->seed, p->capacity = prm->capacity, p->maxWave = prm->max_wave;
    const uint32_t kFull = p->scalar == MPTG_F32 ? pprmK<float>(p->dims, p->capacity) : pprmK<double>(p->dims, p->capacity);
    p->stride = prm->max_k ? prm->max_k : (kFull < MPTG_MAX_K ? kFull : MPTG_MAX_K);
    const size_t sb = (size_t)p->D * p->scalar, W = p->maxWave, K = p->stride;
    int rc = mptg_knn_create(ctx, prm->space, p->capacity, &p->knn);
    if (!rc) rc = uploadBounds(ctx, prm->space, prm->lo, prm->hi, &p->bounds);
    auto alloc = [&](auto** q, size_t bytes) {
        if (rc) return;
        cudaError_t e = cudaMalloc((void**)q, bytes ? bytes : 16);
        if (e != cudaSuccess) rc = fail(ctx, MPTG_ERR_OOM, "mptg_pprm_create: %s", cudaGetErrorString(e));
    };
    alloc(&p->nodes, (size_t)p->capacity * sb);
    alloc(&p->edgeIdx, (size_t)p->capacity * K * 4), alloc(&p->edgeDist, (size_t)p->capacity * K * p->scalar);
    alloc(&p->comp, (size_t)p->capacity * 4), alloc(&p->marks, p->capacity), alloc(&p->lists, (2 + 64 + 4096) * 4);
    alloc(&p->samples, W * sb), alloc(&p->cand, W * sb), alloc(&p->fresh, W * sb);
    alloc(&p->nnIdx, W * K * 4), alloc(&p->nnDist, W * K * p->scalar), alloc(&p->nnCnt, W * 4);
    alloc(&p->from, W * K * sb), alloc(&p->to, W * K * sb), alloc(&p->okEdge, W * K);
    alloc(&p->sel, W * 4), alloc(&p->sel2, W * 4), alloc(&p->nSel, 4), alloc(&p->result, 4), alloc(&p->okValid, W), alloc(&p->keep, W);
    if (!rc) {
        cub::DeviceSelect::Flagged(nullptr, p->selBytes, thrust::counting_iterator<uint32_t>(0), p->keep, p->sel, p->nSel, (int)W);
        alloc(&p->selTemp, p->selBytes);
    }
    if (!rc && p->hasGoal) {
        alloc(&p->goal, sb);
        if (!rc) rc = uploadSync(ctx, p->goal, prm->goal_state, sb);
    }
    if (!rc) rc = memsetSync(ctx, p->lists, 0, (2 + 64 + 4096) * 4);
    if (!rc) rc = memsetSync(ctx, p->result, 0, 4);
    if (!rc && cudaMallocHost((void**)&p->host, 4 * sizeof(uint32_t)) != cudaSuccess) rc = fail(ctx, MPTG_ERR_OOM, "mptg_pprm_create: pinned allocation failed");
    if (rc) {
        pprmFree(p);
        return rc;
    }
    *out = p;
    return MPTG_OK;
}

int mptg_pprm_set_spanner(mptg_pprm* p, double stretch_weight, uint32_t search_capacity) {
    if (!p || !(stretch_weight > 0)) return fail(p ? p->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_pprm_set_spanner: bad argument");
    if (p->size != 0 && p->stretch == 0) return fail(p->ctx, MPTG_ERR_BAD_ARG, "mptg_pprm_set_spanner: the roadmap already holds nodes added without the spanner");
    mptg_ctx* ctx = p->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!p->revHead) {
        // working storage: 72 (double) / 48 (float) bytes per search entry; default = what fits 1 GiB for a full wave, within
        // [256, 4096]; a wave whose searches outgrow it is rerun with four times as much (spannerGrow)
        uint32_t cap = search_capacity;
        if (cap == 0) {
            const size_t per = 2 * (4 + (size_t)p->scalar) + 4 * (4 + (size_t)p->scalar);
            size_t fit = ((size_t)1 << 30) / (per * p->maxWave);
            cap = (uint32_t)(fit < 256 ? 256 : (fit > 4096 ? 4096 : fit));
        }
        int rc = MPTG_OK;
        auto alloc = [&](auto** q, size_t bytes) {
            if (rc) return;
            cudaError_t e = cudaMalloc((void**)q, bytes ? bytes : 16);
            if (e != cudaSuccess) rc = fail(ctx, MPTG_ERR_OOM, "mptg_pprm_set_spanner: %s", cudaGetErrorString(e));
        };
        alloc(&p->revHead, (size_t)p->capacity * 4), alloc(&p->revNext, (size_t)p->capacity * p->stride * 4), alloc(&p->spanErr, 4);
        alloc(&p->okEdgeRaw, (size_t)p->maxWave * p->stride);
        if (!rc) rc = memsetSync(ctx, p->revHead, 0xFF, (size_t)p->capacity * 4);
        if (!rc) rc = memsetSync(ctx, p->spanErr, 0, 4);
        if (!rc) rc = spannerScratch(p, cap);
        if (rc) return rc;
        p->host[2] = 0;
    }
    p->stretch = stretch_weight;
    return MPTG_OK;
}

int mptg_pprm_destroy(mptg_pprm* p) {
    if (!p) return MPTG_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    pprmFree(p);
    return MPTG_OK;
}

int mptg_pprm_add_state(mptg_pprm* p, const void* state, uint32_t marks, uint32_t* node_out) {
    if (!p || !state || (marks & ~(MPTG_PPRM_START | MPTG_PPRM_GOAL))) return fail(p ? p->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_pprm_add_state: bad argument");
    if (p->size >= p->capacity) return fail(p->ctx, MPTG_ERR_CAPACITY, "mptg_pprm_add_state: roadmap is full");
    mptg_ctx* ctx = p->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int rc = uploadSync(ctx, p->samples, state, (size_t)p->D * p->scalar)) return rc;
    uint32_t first = 0, added = 0;
    if (int rc = pprmProcess(p, 1, marks, &first, &added)) return rc;
    if (node_out) *node_out = added ? first : MPTG_NO_INDEX;
    return MPTG_OK;
}

int mptg_pprm_wave(mptg_pprm* p, uint32_t n_samples, uint32_t* size_out, uint32_t* solved_out) {
    if (!p || n_samples == 0 || n_samples > p->maxWave) return fail(p ? p->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_pprm_wave: bad argument");
    mptg_ctx* ctx = p->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    // sample (pprm.hpp:372-374: no goal bias in PPRM); uniform 0 of every sample's stream stays the unused bias draw
    if (p->scalar == MPTG_F32) launchSample<float>(ctx, &p->space, p->bounds, p->seed, p->drawn, n_samples, nullptr, 0.0, p->samples);
    else launchSample<double>(ctx, &p->space, p->bounds, p->seed, p->drawn, n_samples, nullptr, 0.0, p->samples);
    MPTG_LAUNCHED(ctx);
    p->drawn += n_samples;
    uint32_t first = 0, added = 0;
    const int rc = pprmProcess(p, n_samples, 0u, &first, &added);
    ++p->waves;
    if (size_out) *size_out = p->size;
    if (solved_out) *solved_out = p->solved ? 1u : 0u;
    return rc;
}

uint32_t mptg_pprm_size(const mptg_pprm* p) { return p ? p->size : 0; }
uint64_t mptg_pprm_samples_drawn(const mptg_pprm* p) { return p ? p->drawn : 0; }
uint32_t mptg_pprm_row_stride(const mptg_pprm* p) { return p ? p->stride : 0; }

int mptg_pprm_get_graph(mptg_pprm* p, uint32_t first, uint32_t count, void* states_out, uint32_t* edge_idx_out, void* edge_dist_out,
                        uint8_t* marks_out, uint32_t* component_out) {
    if (!p || (uint64_t)first + count > p->size) return fail(p ? p->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_pprm_get_graph: bad range");
    if (count == 0) return MPTG_OK;
    mptg_ctx* ctx = p->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t sb = (size_t)p->D * p->scalar, K = p->stride;
    if (states_out) MPTG_CUDA(ctx, cudaMemcpyAsync(states_out, (char*)p->nodes + first * sb, count * sb, cudaMemcpyDeviceToHost, st));
    if (edge_idx_out) MPTG_CUDA(ctx, cudaMemcpyAsync(edge_idx_out, p->edgeIdx + (size_t)first * K, (size_t)count * K * 4, cudaMemcpyDeviceToHost, st));
    if (edge_dist_out)
        MPTG_CUDA(ctx, cudaMemcpyAsync(edge_dist_out, (char*)p->edgeDist + (size_t)first * K * p->scalar, (size_t)count * K * p->scalar,
                                       cudaMemcpyDeviceToHost, st));
    if (marks_out) MPTG_CUDA(ctx, cudaMemcpyAsync(marks_out, p->marks + first, count, cudaMemcpyDeviceToHost, st));
    if (component_out) {
        void* tmp;
        if (int rc = scratch(ctx, 1, (size_t)count * 4, &tmp)) return rc;
        pprmRootKernel<<<(count + 255) / 256, 256, 0, st>>>(p->comp, first, count, (uint32_t*)tmp);
        MPTG_LAUNCHED(ctx);
        MPTG_CUDA(ctx, cudaMemcpyAsync(component_out, tmp, (size_t)count * 4, cudaMemcpyDeviceToHost, st));
    }
    MPTG_CUDA(ctx, cudaStreamSynchronize(st));
    return MPTG_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// device-resident PRRT* (src/mpt/impl/prrt_star/prrt_star.hpp:510-657), wave-parallel
// ---------------------------------------------------------------------------------------------
// One wave = n_samples iterations of Worker::addSample run as a batch against the tree as it stood when the wave
// began (samples of one wave do not see each other, like the reference's concurrent workers):
//   sample -> nearest -> d == 0 drops -> steer to `range`, distance recomputed (:526-537) -> valid -> link(near, new)
//   -> k nearest, k = ceil(kRRG ln(n+1)) (rrg_rewire_neighbors.hpp:53-67) -> neighbours ranked by cost + distance,
//   tested in that order up to the near node or the cost cut-off, first valid one becomes the parent (:565-605)
//   -> append (:607-622) -> rewire: every unchecked neighbour whose cost would drop is tested (:626-656).
// Rewiring inside a wave is evaluated against the costs at the start of the wave's rewiring step: each old node takes
// the best valid offer (smallest new cost, then smallest (sample, neighbour slot)); all offers are applied at once.
// This cannot close a cycle: afterwards every node's parent has a strictly smaller pre-rewire cost than the node
// itself.  The cost decrease of a re-parented node is then pushed to its whole subtree (nonConcurrentPushUpdate,
// :664-688) by one pass in which every node walks its new ancestor chain and subtracts the decreases it meets,
// nearest ancestor first.
namespace mptg {

// Item counts of the steps of a wave.  A wave with host synchronisation launches exactly `rows` items (nDev null); a QUEUED
// wave launches an upper bound and the kernels read the count the previous step left on the device (starWaveQueuedT).
__device__ __forceinline__ uint32_t starCount(uint32_t rows, const uint32_t* nDev) {
    const uint32_t n = nDev ? *nDev : rows;
    return n < rows ? n : rows;
}

// Ordered compaction of up to a few ten thousand flags by ONE CTA (queued waves: cub::DeviceSelect is two launches and a
// dispatch on the host per call, three calls per wave): out[] = the indices i < n with f0[i] && f1[i] && f2[i] (f1, f2 may be
// null) in increasing order, *count = how many.  16 flags per thread and pass.  `init`: words to preset for the wave
// (init[0] = no goal node yet, init[1] = no rewires yet), null to leave alone.
constexpr int COMPACT_THREADS = 1024, COMPACT_PER = 16;
__global__ void __launch_bounds__(COMPACT_THREADS) starCompactKernel(const uint8_t* __restrict__ f0, const uint8_t* __restrict__ f1,
                                                                     const uint8_t* __restrict__ f2, uint32_t n, uint32_t* __restrict__ out,
                                                                     uint32_t* __restrict__ count, uint32_t* __restrict__ init) {
    __shared__ uint32_t warpSum[COMPACT_THREADS / 32];
    __shared__ uint32_t carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        carry = 0;
        if (init) init[0] = MPTG_NO_INDEX, init[1] = 0u;
    }
    __syncthreads();
    for (uint32_t base = 0; base < n; base += COMPACT_THREADS * COMPACT_PER) {
        const uint32_t first = base + threadIdx.x * COMPACT_PER;
        uint32_t bits = 0;
#pragma unroll
        for (int j = 0; j < COMPACT_PER; ++j) {
            const uint32_t i = first + j;
            if (i < n && f0[i] && (!f1 || f1[i]) && (!f2 || f2[i])) bits |= 1u << j;
        }
        const uint32_t mine = __popc(bits);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warpSum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warpSum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += v;
            }
            warpSum[lane] = w;  // inclusive over the warps
        }
        __syncthreads();
        uint32_t pos = carry + (warp ? warpSum[warp - 1] : 0u) + incl - mine;
        while (bits) {
            const int j = __ffs(bits) - 1;
            bits &= bits - 1u;
            out[pos++] = first + (uint32_t)j;
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += warpSum[COMPACT_THREADS / 32 - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry;
}

template <typename S>
__global__ void starSteerKernel(DevSpace<S> sp, const S* __restrict__ nodes, const S* __restrict__ samples, const uint32_t* __restrict__ nearIdx,
                                const S* __restrict__ nearDist, const uint32_t* __restrict__ nearCnt, uint32_t n, S range, S* __restrict__ from,
                                S* __restrict__ to, S* __restrict__ dNew, uint8_t* __restrict__ alive) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int D = sp.D;
    const S* ps = samples + (size_t)i * D;
    S* pf = from + (size_t)i * D;
    S* pt = to + (size_t)i * D;
    const S d = nearDist[i];
    const bool live = nearCnt[i] != 0 && !(d == S(0));
    alive[i] = live ? 1 : 0;
    dNew[i] = d;
    if (!live) {
        for (int c = 0; c < D; ++c) pf[c] = pt[c] = ps[c];
        return;
    }
    const S* pn = nodes + (size_t)nearIdx[i] * D;
    S a[MPTG_MAX_SCALARS];
    for (int c = 0; c < D; ++c) a[c] = pf[c] = pn[c];
    if (d > range) {
        S b[MPTG_MAX_SCALARS], q[MPTG_MAX_SCALARS];
        for (int c = 0; c < D; ++c) b[c] = ps[c];
        dev::interpolate<S>(sp, a, b, fp::div_(range, d), q);
        for (int c = 0; c < D; ++c) pt[c] = q[c];
        dNew[i] = dev::distance<S>(sp, [&](int c) { return a[c]; }, [&](int c) { return q[c]; });  // :535
    } else {
        for (int c = 0; c < D; ++c) pt[c] = ps[c];
    }
}

template <typename S>
__global__ void starGatherKernel(const uint32_t* __restrict__ sel, uint32_t rows, const uint32_t* __restrict__ nDev, int D, const S* __restrict__ to,
                                 const uint32_t* __restrict__ nearIdx, const S* __restrict__ dNew, S* __restrict__ fresh, uint32_t* __restrict__ nearOf,
                                 S* __restrict__ dOf, const S* __restrict__ filler) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= rows) return;
    if (s >= starCount(rows, nDev)) {  // queued wave: a defined state in the rows past the survivors (their results are never read)
        for (int c = 0; c < D; ++c) fresh[(size_t)s * D + c] = filler[c];
        nearOf[s] = 0;
        dOf[s] = S(0);
        return;
    }
    const uint32_t i = sel[s];
    for (int c = 0; c < D; ++c) fresh[(size_t)s * D + c] = to[(size_t)i * D + c];
    nearOf[s] = nearIdx[i];
    dOf[s] = dNew[i];
}

// one warp per survivor: rank the neighbours by cost + distance (stable), mark the ones the reference's loop could
// test before it stops (cost cut-off or the near node), :565-605
template <typename S>
__global__ void __launch_bounds__(128) starRankKernel(uint32_t rows, const uint32_t* __restrict__ nDev, uint32_t k, const uint32_t* __restrict__ nnIdx, const S* __restrict__ nnDist,
                                                      const uint32_t* __restrict__ nnCnt, const uint32_t* __restrict__ nearOf,
                                                      const S* __restrict__ dOf, const S* __restrict__ cost, uint8_t* __restrict__ order,
                                                      uint8_t* __restrict__ candFlag, uint8_t* __restrict__ checked, uint32_t* __restrict__ limit,
                                                      uint32_t* __restrict__ nearRank, S* __restrict__ defCost) {
    __shared__ S sc[4][MPTG_MAX_K];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t s = blockIdx.x * 4 + warp;
    if (s >= rows) return;
    if (s >= starCount(rows, nDev)) {  // queued wave: no candidates in the rows past the survivors
        for (uint32_t j = lane; j < k; j += 32) candFlag[(size_t)s * k + j] = 0;
        return;
    }
    const uint32_t cnt = nnCnt[s], near = nearOf[s];
    const S parentCost = cost[near] + dOf[s];
    for (uint32_t j = lane; j < cnt; j += 32) sc[warp][j] = cost[nnIdx[(size_t)s * k + j]] + nnDist[(size_t)s * k + j];
    __syncwarp();
    uint32_t cutLocal = 0, nearRankLocal = 0xFFFFFFFFu;
    uint32_t myRank[MPTG_MAX_K / 32];
#pragma unroll
    for (int r = 0; r < MPTG_MAX_K / 32; ++r) {
        const uint32_t j = (uint32_t)r * 32 + lane;
        myRank[r] = 0xFFFFFFFFu;
        if (j < cnt) {
            const S c = sc[warp][j];
            uint32_t rank = 0;
            for (uint32_t i = 0; i < cnt; ++i) {
                const S o = sc[warp][i];
                rank += (o < c || (o == c && i < j)) ? 1u : 0u;
            }
            myRank[r] = rank;
            order[(size_t)s * k + rank] = (uint8_t)j;
            if (!(c > parentCost)) ++cutLocal;
            if (nnIdx[(size_t)s * k + j] == near) nearRankLocal = rank;
        }
    }
    uint32_t rCut = cutLocal, rNear = nearRankLocal;
    for (int o = 16; o > 0; o >>= 1) {
        rCut += __shfl_xor_sync(0xffffffffu, rCut, o);
        rNear = min(rNear, __shfl_xor_sync(0xffffffffu, rNear, o));
    }
    // costs are sorted by rank, so "newCost > parentCost" holds exactly for ranks >= rCut
    const uint32_t lim = min(rCut, rNear);
    const bool nearReached = rNear < rCut;
#pragma unroll
    for (int r = 0; r < MPTG_MAX_K / 32; ++r) {
        const uint32_t j = (uint32_t)r * 32 + lane;
        if (j < k) {
            const bool test = j < cnt && myRank[r] < lim;
            candFlag[(size_t)s * k + j] = test ? 1 : 0;
            checked[(size_t)s * k + j] = 0;  // set by starAppendKernel: the loop marks what it visited before it stopped
        }
    }
    if (lane == 0) {
        limit[s] = lim;
        nearRank[s] = nearReached ? rNear : 0xFFFFFFFFu;
        S dc = parentCost;
        if (nearReached) dc = sc[warp][order[(size_t)s * k + rNear]];  // parent stays the near node, cost from the neighbour list (:588-592)
        defCost[s] = dc;
    }
}

// edges for the flagged (survivor, slot) pairs; fromNode: neighbour -> new state (parent candidates) or new state -> neighbour (rewiring)
template <typename S>
__global__ void starEdgeKernel(const uint32_t* __restrict__ ids, uint32_t rows, const uint32_t* __restrict__ nDev, uint32_t k, int D, bool towardsFresh,
                               const S* __restrict__ fresh, const S* __restrict__ nodes, const uint32_t* __restrict__ nnIdx, S* __restrict__ from,
                               S* __restrict__ to, uint32_t* __restrict__ inv) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows) return;
    if (e >= starCount(rows, nDev)) {  // queued wave: the edges past the count go from the root to itself (one state to check)
        for (int c = 0; c < D; ++c) from[(size_t)e * D + c] = to[(size_t)e * D + c] = nodes[c];
        return;
    }
    const uint32_t id = ids[e], s = id / k;
    if (inv) inv[id] = e;
    const S* pf = fresh + (size_t)s * D;
    const S* pn = nodes + (size_t)nnIdx[id] * D;
    const S* a = towardsFresh ? pn : pf;
    const S* b = towardsFresh ? pf : pn;
    for (int c = 0; c < D; ++c) from[(size_t)e * D + c] = a[c], to[(size_t)e * D + c] = b[c];
}

// first valid candidate in rank order becomes the parent; append node, parent, cost; goal test (:607-622)
template <typename S>
__global__ void starAppendKernel(DevSpace<S> sp, uint32_t rows, const uint32_t* __restrict__ nDev, uint32_t k, const S* __restrict__ fresh, const uint32_t* __restrict__ nnIdx,
                                 const S* __restrict__ nnDist, const uint8_t* __restrict__ order, const uint32_t* __restrict__ limit,
                                 const uint32_t* __restrict__ nearRank, uint8_t* __restrict__ checked, const uint32_t* __restrict__ inv,
                                 const uint8_t* __restrict__ okCand, const uint32_t* __restrict__ nearOf, const S* __restrict__ defCost, uint32_t size, const S* __restrict__ goal, S goalRadius, S* __restrict__ nodes,
                                 uint32_t* __restrict__ parent, S* __restrict__ cost, uint32_t* __restrict__ goalList,
                                 const uint32_t* __restrict__ nnCnt, uint8_t* __restrict__ rewireFlag) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= starCount(rows, nDev)) {
        if (rewireFlag && s < rows)
            for (uint32_t j = 0; j < k; ++j) rewireFlag[(size_t)s * k + j] = 0;
        return;
    }
    const int D = sp.D;
    uint32_t par = nearOf[s];
    S c = defCost[s];
    const uint32_t lim = limit[s];
    bool found = false;
    for (uint32_t r = 0; r < lim && !found; ++r) {
        const uint32_t j = order[(size_t)s * k + r];
        checked[(size_t)s * k + j] = 1;  // :586 "mark as checked"
        if (okCand[inv[(size_t)s * k + j]]) {
            par = nnIdx[(size_t)s * k + j];
            c = cost[par] + nnDist[(size_t)s * k + j];
            found = true;
        }
    }
    if (!found && nearRank[s] != 0xFFFFFFFFu) checked[(size_t)s * k + order[(size_t)s * k + nearRank[s]]] = 1;  // the loop reached the near node
    const uint32_t id = size + s;
    const S* pf = fresh + (size_t)s * D;
    for (int d = 0; d < D; ++d) nodes[(size_t)id * D + d] = pf[d];
    parent[id] = par;
    cost[id] = c;
    if (goal != nullptr) {
        const S dg = dev::distance<S>(sp, [&](int d) { return pf[d]; }, [&](int d) { return goal[d]; });
        if (dg <= goalRadius) {
            const uint32_t slot = atomicAdd(goalList, 1u);
            if (slot < 65535u) goalList[1 + slot] = id;
        }
    }
    if (rewireFlag) {  // queued wave: the offers of this row (starRewireFlagKernel) -- they need this row's cost and marks only
        const uint32_t cnt = nnCnt[s];
        for (uint32_t j = 0; j < k; ++j) {
            const size_t e = (size_t)s * k + j;
            bool f = false;
            if (j < cnt && !checked[e]) f = c + nnDist[e] < cost[nnIdx[e]];
            rewireFlag[e] = f ? 1 : 0;
        }
    }
}

// rewiring offers: unchecked neighbours whose cost would drop (:626-637)
template <typename S>
__global__ void starRewireFlagKernel(uint32_t rows, const uint32_t* __restrict__ nDev, uint32_t k, uint32_t size, const uint32_t* __restrict__ nnIdx, const S* __restrict__ nnDist,
                                     const uint32_t* __restrict__ nnCnt, const uint8_t* __restrict__ checked, const S* __restrict__ cost,
                                     uint8_t* __restrict__ flag) {
    const size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (size_t)rows * k) return;
    const uint32_t s = (uint32_t)(id / k), j = (uint32_t)(id % k);
    bool f = false;
    if (s < starCount(rows, nDev) && j < nnCnt[s] && !checked[id]) f = cost[size + s] + nnDist[id] < cost[nnIdx[id]];
    flag[id] = f ? 1 : 0;
}

__device__ __forceinline__ unsigned long long starCostKey(double c) {  // costs are non-negative: the bit pattern orders them
    return (unsigned long long)__double_as_longlong(c + 0.0);
}
template <typename S>
__global__ void starRewireMinKernel(const uint32_t* __restrict__ ids, const uint8_t* __restrict__ ok, uint32_t rows, const uint32_t* __restrict__ nDev, uint32_t k, uint32_t size,
                                    const uint32_t* __restrict__ nnIdx, const S* __restrict__ nnDist, const S* __restrict__ cost,
                                    unsigned long long* __restrict__ bestKey) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= starCount(rows, nDev) || !ok[e]) return;
    const uint32_t id = ids[e];
    atomicMin(bestKey + nnIdx[id], starCostKey((double)(cost[size + id / k] + nnDist[id])));
}
template <typename S>
__global__ void starRewirePickKernel(const uint32_t* __restrict__ ids, const uint8_t* __restrict__ ok, uint32_t rows, const uint32_t* __restrict__ nDev, uint32_t k, uint32_t size,
                                     const uint32_t* __restrict__ nnIdx, const S* __restrict__ nnDist, const S* __restrict__ cost,
                                     const unsigned long long* __restrict__ bestKey, uint32_t* __restrict__ bestCand) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= starCount(rows, nDev) || !ok[e]) return;
    const uint32_t id = ids[e], nb = nnIdx[id];
    if (starCostKey((double)(cost[size + id / k] + nnDist[id])) == bestKey[nb]) atomicMin(bestCand + nb, e);
}
template <typename S>
__global__ void starRewireApplyKernel(const uint32_t* __restrict__ ids, const uint8_t* __restrict__ ok, uint32_t rows, const uint32_t* __restrict__ nDev, uint32_t k, uint32_t size,
                                      const uint32_t* __restrict__ nnIdx, const S* __restrict__ nnDist, const S* __restrict__ cost,
                                      const uint32_t* __restrict__ bestCand, uint32_t* __restrict__ parent, S* __restrict__ delta,
                                      uint32_t* __restrict__ counters) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= starCount(rows, nDev) || !ok[e]) return;
    const uint32_t id = ids[e], nb = nnIdx[id];
    if (bestCand[nb] != e) return;
    const uint32_t from = size + id / k;
    parent[nb] = from;
    delta[nb] = cost[nb] - (cost[from] + nnDist[id]);  // :647 (cost[] itself is updated by the push pass below)
    atomicAdd(counters, 1u);
}
// nonConcurrentPushUpdate (:664-688) for all re-parented nodes at once
template <typename S>
__global__ void starPushKernel(uint32_t nBefore, uint32_t rows, const uint32_t* __restrict__ nDev, const uint32_t* __restrict__ parent,
                               const S* __restrict__ delta, S* __restrict__ cost, const uint32_t* __restrict__ applied,
                               unsigned long long* __restrict__ bestKey, uint32_t* __restrict__ bestCand) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = nBefore + starCount(rows, nDev);  // the tree after this wave's appends
    if (i >= n) return;
    if (*applied == 0u) return;  // no node was re-parented in this wave: every decrease is zero, nothing to push
    // the offers of this wave are spent (an offer exists only where something was applied): leave the per-node slots of the
    // old nodes as the next wave expects them, so that a queued wave needs no memset for them
    if (i < nBefore) bestKey[i] = ~0ull, bestCand[i] = MPTG_NO_INDEX;
    S c = cost[i];
    bool changed = false;
    uint32_t steps = 0;  // a path has fewer than n nodes; the bound keeps a corrupted tree from hanging the device
    for (uint32_t a = i; a != MPTG_NO_INDEX && steps < n; a = parent[a], ++steps) {
        const S d = delta[a];
        if (d > S(0)) c = c - d, changed = true;
    }
    if (changed) cost[i] = c;
}
// best goal node: smallest cost, then smallest index
template <typename S>
__global__ void starGoalKernel(const uint32_t* __restrict__ goalList, const S* __restrict__ cost, uint32_t* __restrict__ result) {
    __shared__ unsigned long long best;
    if (threadIdx.x == 0) best = ~0ull;
    __syncthreads();
    const uint32_t n = min(goalList[0], 65535u);
    unsigned long long mine = ~0ull;
    uint32_t mineIdx = MPTG_NO_INDEX;
    for (uint32_t g = threadIdx.x; g < n; g += blockDim.x) {
        const uint32_t id = goalList[1 + g];
        const unsigned long long key = starCostKey((double)cost[id]);
        if (key < mine || (key == mine && id < mineIdx)) mine = key, mineIdx = id;
    }
    atomicMin(&best, mine);
    __syncthreads();
    if (mineIdx != MPTG_NO_INDEX && mine == best) atomicMin(result, mineIdx);
}

// The rewiring tail of a queued wave on a young tree as ONE CTA: clear the decreases, best offer per old node (the three
// passes of starRewireMin/Pick/ApplyKernel), push the decreases down the subtrees, restore the offer slots, best goal node.
// Same arithmetic and the same atomics as the separate kernels -- the barriers stand where the launches were; six
// operations of a wave whose cost is its number of operations.  result[0] / result[1] are preset by the
// wave's first compaction.
constexpr int TAIL_THREADS = 1024;
template <typename S>
__global__ void __launch_bounds__(TAIL_THREADS) starRewireTailKernel(const uint32_t* __restrict__ ids, const uint8_t* __restrict__ ok, uint32_t rows,
                                                                    const uint32_t* __restrict__ nRDev, uint32_t waveRows, const uint32_t* __restrict__ nSDev,
                                                                    uint32_t k, uint32_t size, const uint32_t* __restrict__ nnIdx, const S* __restrict__ nnDist,
                                                                    S* cost, unsigned long long* bestKey, uint32_t* bestCand, uint32_t* parent, S* delta,
                                                                    const uint32_t* goalList, uint32_t* result) {
    __shared__ unsigned long long best;
    __shared__ uint32_t applied;
    const uint32_t nR = starCount(rows, nRDev), n = size + starCount(waveRows, nSDev);
    if (threadIdx.x == 0) best = ~0ull, applied = 0u;
    for (uint32_t i = threadIdx.x; i < n; i += TAIL_THREADS) delta[i] = S(0);
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < nR; e += TAIL_THREADS) {  // starRewireMinKernel
        if (!ok[e]) continue;
        const uint32_t id = ids[e];
        atomicMin(bestKey + nnIdx[id], starCostKey((double)(cost[size + id / k] + nnDist[id])));
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < nR; e += TAIL_THREADS) {  // starRewirePickKernel
        if (!ok[e]) continue;
        const uint32_t id = ids[e], nb = nnIdx[id];
        if (starCostKey((double)(cost[size + id / k] + nnDist[id])) == *(volatile unsigned long long*)(bestKey + nb)) atomicMin(bestCand + nb, e);
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < nR; e += TAIL_THREADS) {  // starRewireApplyKernel
        if (!ok[e]) continue;
        const uint32_t id = ids[e], nb = nnIdx[id];
        if (*(volatile uint32_t*)(bestCand + nb) != e) continue;
        const uint32_t from = size + id / k;
        parent[nb] = from;
        delta[nb] = cost[nb] - (cost[from] + nnDist[id]);  // :647
        atomicAdd(&applied, 1u);
    }
    __syncthreads();
    if (applied != 0u) {  // starPushKernel
        if (threadIdx.x == 0) result[1] = applied;
        // two passes: every node reads the OLD costs of nobody but itself, yet the decreases must all be read before the
        // offer slots are restored -- they are different arrays, so one loop does both
        for (uint32_t i = threadIdx.x; i < n; i += TAIL_THREADS) {
            S c = cost[i];
            bool changed = false;
            uint32_t steps = 0;
            for (uint32_t a = i; a != MPTG_NO_INDEX && steps < n; a = *(volatile uint32_t*)(parent + a), ++steps) {
                const S d = *(volatile S*)(delta + a);
                if (d > S(0)) c = c - d, changed = true;
            }
            if (changed) cost[i] = c;
            if (i < size) bestKey[i] = ~0ull, bestCand[i] = MPTG_NO_INDEX;
        }
    }
    __syncthreads();
    if (goalList != nullptr) {  // starGoalKernel
        const uint32_t ng = min(goalList[0], 65535u);
        unsigned long long mine = ~0ull;
        uint32_t mineIdx = MPTG_NO_INDEX;
        for (uint32_t g = threadIdx.x; g < ng; g += TAIL_THREADS) {
            const uint32_t id = goalList[1 + g];
            const unsigned long long key = starCostKey((double)*(volatile S*)(cost + id));
            if (key < mine || (key == mine && id < mineIdx)) mine = key, mineIdx = id;
        }
        atomicMin(&best, mine);
        __syncthreads();
        if (mineIdx != MPTG_NO_INDEX && mine == best) atomicMin(result, mineIdx);
    }
}

}  // namespace mptg

struct mptg_prrtstar {
    mptg_ctx* ctx = nullptr;
    mptg_geom* geom = nullptr;
    mptg_knn* knn = nullptr;  // owned
    mptg_space_desc space{};
    int D = 0, scalar = MPTG_F32, dims = 0;
    double range = 0, goalBias = 0, goalRadius = 0, linkStep = 0, rewireFactor = 1.1, rRRG = 0;  // rRRG > 0: radius rewiring
    bool hasGoal = false;
    uint64_t seed = 0, drawn = 0, waves = 0, rewires = 0;
    uint32_t capacity = 0, size = 0, maxWave = 0, stride = 0, goalNode = MPTG_NO_INDEX;
    // tree
    void *bounds = nullptr, *goal = nullptr, *nodes = nullptr, *cost = nullptr, *delta = nullptr;
    uint32_t *parent = nullptr, *bestCand = nullptr, *goalList = nullptr;
    unsigned long long* bestKey = nullptr;
    // wave
    void *samples = nullptr, *from = nullptr, *to = nullptr, *fresh = nullptr, *nearDist = nullptr, *dNew = nullptr, *dOf = nullptr, *defCost = nullptr;
    void *nnDist = nullptr, *eFrom = nullptr, *eTo = nullptr;
    uint32_t *nearIdx = nullptr, *nearCnt = nullptr, *sel = nullptr, *nSel = nullptr, *nearOf = nullptr, *nnIdx = nullptr, *nnCnt = nullptr,
             *limit = nullptr, *nearRank = nullptr, *ids = nullptr, *inv = nullptr, *result = nullptr;
    uint8_t *alive = nullptr, *okValid = nullptr, *okLink = nullptr, *keep = nullptr, *order = nullptr, *flag = nullptr, *checked = nullptr, *okEdge = nullptr;
    void* selTemp = nullptr;
    size_t selBytes = 0;
    uint32_t* host = nullptr;  // pinned: [0] select count, [1] best goal node, [2] rewires applied, [3] survivors of a queued wave
    uint32_t* cnt = nullptr;   // device: survivors, candidate-parent edges, rewiring edges of a queued wave
    uint32_t queuedMax = 1024; // waves up to this size run without host synchronisation between their steps (starWaveQueuedT)
    bool timing = false;       // MPTG_STAR_TIMING=1: host time spent issuing the queued waves / waiting for them, printed at destroy
    double issueUs = 0, waitUs = 0;
    uint64_t queuedWaves = 0, queuedLaunches = 0;
};

namespace {

void starFree(mptg_prrtstar* p) {
    if (!p) return;
    if (p->knn) mptg_knn_destroy(p->knn);
    for (void* q : {p->bounds, p->goal, p->nodes, p->cost, p->delta, (void*)p->parent, (void*)p->bestCand, (void*)p->goalList, (void*)p->bestKey,
                    p->samples, p->from, p->to, p->fresh, p->nearDist, p->dNew, p->dOf, p->defCost, p->nnDist, p->eFrom, p->eTo, (void*)p->nearIdx,
                    (void*)p->nearCnt, (void*)p->sel, (void*)p->nSel, (void*)p->nearOf, (void*)p->nnIdx, (void*)p->nnCnt, (void*)p->limit, (void*)p->nearRank,
                    (void*)p->ids, (void*)p->inv, (void*)p->result, (void*)p->alive, (void*)p->okValid, (void*)p->okLink, (void*)p->keep,
                    (void*)p->order, (void*)p->flag, (void*)p->checked, (void*)p->okEdge, p->selTemp, (void*)p->cnt})
        cudaFree(q);
    if (p->host) cudaFreeHost(p->host);
    delete p;
}

// k = ceil(kRRG ln(n + 1)), kRRG = rewireFactor e (1 + 1/d) (rrg_rewire_neighbors.hpp:53-61), in the space's scalar type
template <typename S>
uint32_t starK(double rewireFactor, int dims, uint32_t n) {
    const S e = (S)2.718281828459045235360287471352662498L;
    const S kRRG = (S)rewireFactor * e * (S(1) + S(1) / (S)dims);
    const int k = (int)std::ceil(kRRG * std::log((S)(n + 1.0)));
    return (uint32_t)(k < 1 ? 1 : k);
}

int starSelect(mptg_prrtstar* p, const uint8_t* flags, size_t n, uint32_t* out, uint32_t* countHost) {
    mptg_ctx* ctx = p->ctx;
    size_t bytes = p->selBytes;
    MPTG_CUDA(ctx, cub::DeviceSelect::Flagged(p->selTemp, bytes, thrust::counting_iterator<uint32_t>(0), flags, out, p->nSel, (int)n, ctx->stream));
    MPTG_LAUNCHED(ctx);
    MPTG_CUDA(ctx, cudaMemcpyAsync(p->host, p->nSel, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *countHost = p->host[0];
    return MPTG_OK;
}

// the buffers whose size depends on the neighbour-row stride (k-nearest: k at capacity; radius rewiring: MPTG_MAX_K)
int starAllocNeighbourBuffers(mptg_prrtstar* p) {
    mptg_ctx* ctx = p->ctx;
    for (void** q : {&p->nnDist, &p->eFrom, &p->eTo, (void**)&p->nnIdx, (void**)&p->ids, (void**)&p->inv, (void**)&p->order, (void**)&p->flag,
                     (void**)&p->checked, (void**)&p->okEdge, &p->selTemp}) {
        if (*q) cudaFree(*q);
        *q = nullptr;
    }
    const size_t sb = (size_t)p->D * p->scalar, W = p->maxWave, K = p->stride, sc = p->scalar;
    int rc = MPTG_OK;
    auto alloc = [&](auto** q, size_t bytes) {
        if (rc) return;
        cudaError_t e = cudaMalloc((void**)q, bytes ? bytes : 16);
        if (e != cudaSuccess) rc = fail(ctx, MPTG_ERR_OOM, "mptg_prrtstar: %s", cudaGetErrorString(e));
    };
    alloc(&p->nnIdx, W * K * 4), alloc(&p->nnDist, W * K * sc);
    alloc(&p->order, W * K), alloc(&p->flag, W * K), alloc(&p->checked, W * K), alloc(&p->okEdge, W * K);
    alloc(&p->ids, W * K * 4), alloc(&p->inv, W * K * 4), alloc(&p->eFrom, W * K * sb), alloc(&p->eTo, W * K * sb);
    if (!rc) {
        cub::DeviceSelect::Flagged(nullptr, p->selBytes, thrust::counting_iterator<uint32_t>(0), p->flag, p->ids, p->nSel, (int)(W * K));
        alloc(&p->selTemp, p->selBytes);
    }
    return rc;
}

// A wave WITHOUT host synchronisation between its steps, for the small waves of a young tree: the
// three compaction counts (survivors, candidate-parent edges, rewiring edges) stay on the device, every step is launched
// over an upper bound of its items -- W survivors, W k edges -- and reads the count the step before it left (starCount);
// rows past a count hold defined filler whose results nobody reads.  Same kernels, same arithmetic, same tree as
// starWaveT (tests/test_gpu_parity.py compares the two on the same seed); one synchronisation at the end, for the new size.
// Launching the bounds costs nothing while W k is a few ten thousand edges; larger waves keep the exact launches.
template <typename S>
int starWaveQueuedT(mptg_prrtstar* p, uint32_t W) {
    mptg_ctx* ctx = p->ctx;
    cudaStream_t st = ctx->stream;
    const auto t0 = std::chrono::steady_clock::now();
    const uint64_t launches0 = ctx->launches;
    const DevSpace<S> sp = makeDevSpace<S>(p->space);
    const int D = p->D;
    const uint32_t grid = (W + 127) / 128;
    const bool biased = p->hasGoal && p->goalBias > 0 && p->goalNode == MPTG_NO_INDEX;  // :468-481
    sampleKernel<S><<<grid, 128, 0, st>>>(sp, (const S*)p->bounds, (const S*)p->bounds + D, p->seed, p->drawn, W, biased ? (const S*)p->goal : nullptr,
                                          (S)p->goalBias, (S*)p->samples);
    MPTG_LAUNCHED(ctx);
    p->drawn += W;
    if (int rc = mptg_knn_query_dev(p->knn, p->samples, W, 1, -1.0, p->nearIdx, p->nearDist, p->nearCnt)) return rc;  // :517
    starSteerKernel<S><<<grid, 128, 0, st>>>(sp, (const S*)p->nodes, (const S*)p->samples, p->nearIdx, (const S*)p->nearDist, p->nearCnt, W,
                                             (S)p->range, (S*)p->from, (S*)p->to, (S*)p->dNew, p->alive);
    MPTG_LAUNCHED(ctx);
    if (int rc = mptg_valid_batch_dev(p->geom, p->to, W, p->okValid, nullptr)) return rc;                              // :539
    if (int rc = mptg_link_batch_dev(p->geom, &p->space, p->from, p->to, W, p->linkStep, p->okLink, nullptr)) return rc;  // :545
    // result[0] best goal node, [1] rewires applied, [2] survivors: one copy to the host at the end
    uint32_t *nS = p->result + 2, *nE = p->cnt, *nR = p->cnt + 1;
    auto select = [&](const uint8_t* f0, const uint8_t* f1, const uint8_t* f2, uint32_t n, uint32_t* out, uint32_t* count, uint32_t* init) {
        starCompactKernel<<<1, COMPACT_THREADS, 0, st>>>(f0, f1, f2, n, out, count, init);
        MPTG_LAUNCHED(ctx);
    };
    select(p->alive, p->okValid, p->okLink, W, p->sel, nS, p->result);
    ++p->waves;
    starGatherKernel<S><<<grid, 128, 0, st>>>(p->sel, W, nS, D, (const S*)p->to, p->nearIdx, (const S*)p->dNew, (S*)p->fresh, p->nearOf, (S*)p->dOf,
                                              (const S*)p->nodes);
    MPTG_LAUNCHED(ctx);
    // neighbourhoods (:559-562)
    uint32_t k = starK<S>(p->rewireFactor, p->dims, p->size);
    double radius = -1.0;
    if (p->rRRG > 0) {
        const S n1 = (S)(p->size + 1.0);
        radius = (double)((S)p->rRRG * std::pow(std::log(n1) / n1, S(1) / (S)p->dims));
        k = p->stride;
    }
    if (k > p->stride) k = p->stride;
    if (int rc = mptg_knn_query_dev(p->knn, p->fresh, W, k, radius, p->nnIdx, p->nnDist, p->nnCnt)) return rc;
    starRankKernel<S><<<(W + 3) / 4, 128, 0, st>>>(W, nS, k, p->nnIdx, (const S*)p->nnDist, p->nnCnt, p->nearOf, (const S*)p->dOf, (const S*)p->cost, p->order,
                                                  p->flag, p->checked, p->limit, p->nearRank, (S*)p->defCost);
    MPTG_LAUNCHED(ctx);
    // candidate parents, one link batch (:580-605)
    const uint32_t WK = W * k, gridE = (WK + 127) / 128;
    select(p->flag, nullptr, nullptr, WK, p->ids, nE, nullptr);
    starEdgeKernel<S><<<gridE, 128, 0, st>>>(p->ids, WK, nE, k, D, true, (const S*)p->fresh, (const S*)p->nodes, p->nnIdx, (S*)p->eFrom, (S*)p->eTo, p->inv);
    MPTG_LAUNCHED(ctx);
    if (int rc = mptg_link_batch_dev(p->geom, &p->space, p->eFrom, p->eTo, WK, p->linkStep, p->okEdge, nullptr)) return rc;
    starAppendKernel<S><<<grid, 128, 0, st>>>(sp, W, nS, k, (const S*)p->fresh, p->nnIdx, (const S*)p->nnDist, p->order, p->limit, p->nearRank, p->checked,
                                              p->inv, p->okEdge, p->nearOf, (const S*)p->defCost, p->size, p->hasGoal ? (const S*)p->goal : nullptr,
                                              (S)p->goalRadius, (S*)p->nodes, p->parent, (S*)p->cost, p->goalList, p->nnCnt, p->flag);
    MPTG_LAUNCHED(ctx);
    // rewire (:626-656): the offers were flagged by the append kernel, row by row
    select(p->flag, nullptr, nullptr, WK, p->ids, nR, nullptr);
    starEdgeKernel<S><<<gridE, 128, 0, st>>>(p->ids, WK, nR, k, D, false, (const S*)p->fresh, (const S*)p->nodes, p->nnIdx, (S*)p->eFrom, (S*)p->eTo, nullptr);
    MPTG_LAUNCHED(ctx);
    if (int rc = mptg_link_batch_dev(p->geom, &p->space, p->eFrom, p->eTo, WK, p->linkStep, p->okEdge, nullptr)) return rc;
    const uint32_t bound = p->size + W;  // <= capacity (checked by the caller)
    if (bound <= (1u << 15)) {  // young tree: the whole tail in one CTA
        starRewireTailKernel<S><<<1, TAIL_THREADS, 0, st>>>(p->ids, p->okEdge, WK, nR, W, nS, k, p->size, p->nnIdx, (const S*)p->nnDist, (S*)p->cost, p->bestKey,
                                                           p->bestCand, p->parent, (S*)p->delta, p->hasGoal ? p->goalList : nullptr, p->result);
        MPTG_LAUNCHED(ctx);
    } else {
    MPTG_CUDA(ctx, cudaMemsetAsync(p->delta, 0, (size_t)bound * sizeof(S), st));  // (bestKey / bestCand: left clean by every push pass)
    starRewireMinKernel<S><<<gridE, 128, 0, st>>>(p->ids, p->okEdge, WK, nR, k, p->size, p->nnIdx, (const S*)p->nnDist, (const S*)p->cost, p->bestKey);
    MPTG_LAUNCHED(ctx);
    starRewirePickKernel<S><<<gridE, 128, 0, st>>>(p->ids, p->okEdge, WK, nR, k, p->size, p->nnIdx, (const S*)p->nnDist, (const S*)p->cost, p->bestKey,
                                                  p->bestCand);
    MPTG_LAUNCHED(ctx);
    starRewireApplyKernel<S><<<gridE, 128, 0, st>>>(p->ids, p->okEdge, WK, nR, k, p->size, p->nnIdx, (const S*)p->nnDist, (const S*)p->cost, p->bestCand,
                                                   p->parent, (S*)p->delta, p->result + 1);
    MPTG_LAUNCHED(ctx);
    starPushKernel<S><<<(bound + 127) / 128, 128, 0, st>>>(p->size, W, nS, p->parent, (const S*)p->delta, (S*)p->cost, p->result + 1, p->bestKey, p->bestCand);
    MPTG_LAUNCHED(ctx);
    if (p->hasGoal) {
        starGoalKernel<S><<<1, 256, 0, st>>>(p->goalList, (const S*)p->cost, p->result);
        MPTG_LAUNCHED(ctx);
    }
    }
    MPTG_CUDA(ctx, cudaMemcpyAsync(p->host + 1, p->result, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    const auto t1 = std::chrono::steady_clock::now();
    MPTG_CUDA(ctx, cudaStreamSynchronize(st));
    if (p->timing) {
        const auto t2 = std::chrono::steady_clock::now();
        p->issueUs += std::chrono::duration<double, std::micro>(t1 - t0).count();
        p->waitUs += std::chrono::duration<double, std::micro>(t2 - t1).count();
        p->queuedLaunches += ctx->launches - launches0;
        ++p->queuedWaves;
    }
    const uint32_t added = p->host[3];
    p->goalNode = p->host[1];
    p->rewires += p->host[2];
    if (added) {
        uint32_t first = 0;
        if (int rc = mptg_knn_insert_dev(p->knn, p->fresh, added, &first)) return rc;  // :619 (asynchronous: the next wave's search follows it on the stream)
        if (first != p->size) return fail(ctx, MPTG_ERR_CUDA, "mptg_prrtstar_wave: node numbering out of step");
    }
    p->size += added;
    return MPTG_OK;
}

template <typename S>
int starWaveT(mptg_prrtstar* p, uint32_t W) {
    if (W <= p->queuedMax && (uint64_t)p->size + W <= p->capacity && (uint64_t)W * p->stride <= (1u << 16)) return starWaveQueuedT<S>(p, W);
    mptg_ctx* ctx = p->ctx;
    cudaStream_t st = ctx->stream;
    const DevSpace<S> sp = makeDevSpace<S>(p->space);
    const int D = p->D;
    const uint32_t grid = (W + 127) / 128;
    const bool biased = p->hasGoal && p->goalBias > 0 && p->goalNode == MPTG_NO_INDEX;  // :468-481
    sampleKernel<S><<<grid, 128, 0, st>>>(sp, (const S*)p->bounds, (const S*)p->bounds + D, p->seed, p->drawn, W, biased ? (const S*)p->goal : nullptr,
                                          (S)p->goalBias, (S*)p->samples);
    MPTG_LAUNCHED(ctx);
    p->drawn += W;
    if (int rc = mptg_knn_query_dev(p->knn, p->samples, W, 1, -1.0, p->nearIdx, p->nearDist, p->nearCnt)) return rc;  // :517
    starSteerKernel<S><<<grid, 128, 0, st>>>(sp, (const S*)p->nodes, (const S*)p->samples, p->nearIdx, (const S*)p->nearDist, p->nearCnt, W,
                                             (S)p->range, (S*)p->from, (S*)p->to, (S*)p->dNew, p->alive);
    MPTG_LAUNCHED(ctx);
    if (int rc = mptg_valid_batch_dev(p->geom, p->to, W, p->okValid, nullptr)) return rc;                              // :539
    if (int rc = mptg_link_batch_dev(p->geom, &p->space, p->from, p->to, W, p->linkStep, p->okLink, nullptr)) return rc;  // :545
    prrtFlagKernel<<<grid, 128, 0, st>>>(p->alive, p->okValid, p->okLink, W, p->keep);
    MPTG_LAUNCHED(ctx);
    uint32_t nS = 0;
    if (int rc = starSelect(p, p->keep, W, p->sel, &nS)) return rc;
    if (nS > p->capacity - p->size) nS = p->capacity - p->size;
    ++p->waves;
    if (nS == 0) return MPTG_OK;
    starGatherKernel<S><<<(nS + 127) / 128, 128, 0, st>>>(p->sel, nS, nullptr, D, (const S*)p->to, p->nearIdx, (const S*)p->dNew, (S*)p->fresh, p->nearOf,
                                                         (S*)p->dOf, nullptr);
    MPTG_LAUNCHED(ctx);
    // neighbourhoods (:559-562)
    uint32_t k = starK<S>(p->rewireFactor, p->dims, p->size);
    double radius = -1.0;
    if (p->rRRG > 0) {  // rewire_r_nearest (rrg_rewire_neighbors.hpp:102-128): everything within r(n), at most MPTG_MAX_K
        const S n1 = (S)(p->size + 1.0);
        radius = (double)((S)p->rRRG * std::pow(std::log(n1) / n1, S(1) / (S)p->dims));
        k = p->stride;
    }
    if (k > p->stride) k = p->stride;
    if (int rc = mptg_knn_query_dev(p->knn, p->fresh, nS, k, radius, p->nnIdx, p->nnDist, p->nnCnt)) return rc;
    starRankKernel<S><<<(nS + 3) / 4, 128, 0, st>>>(nS, nullptr, k, p->nnIdx, (const S*)p->nnDist, p->nnCnt, p->nearOf, (const S*)p->dOf, (const S*)p->cost,
                                                   p->order, p->flag, p->checked, p->limit, p->nearRank, (S*)p->defCost);
    MPTG_LAUNCHED(ctx);
    // candidate parents, one link batch (:580-605)
    uint32_t nE = 0;
    if (int rc = starSelect(p, p->flag, (size_t)nS * k, p->ids, &nE)) return rc;
    if (nE) {
        starEdgeKernel<S><<<(nE + 127) / 128, 128, 0, st>>>(p->ids, nE, nullptr, k, D, true, (const S*)p->fresh, (const S*)p->nodes, p->nnIdx, (S*)p->eFrom,
                                                           (S*)p->eTo, p->inv);
        MPTG_LAUNCHED(ctx);
        if (int rc = mptg_link_batch_dev(p->geom, &p->space, p->eFrom, p->eTo, nE, p->linkStep, p->okEdge, nullptr)) return rc;
    }
    starAppendKernel<S><<<(nS + 127) / 128, 128, 0, st>>>(sp, nS, nullptr, k, (const S*)p->fresh, p->nnIdx, (const S*)p->nnDist, p->order, p->limit, p->nearRank,
                                                         p->checked, p->inv, p->okEdge, p->nearOf, (const S*)p->defCost, p->size, p->hasGoal ? (const S*)p->goal : nullptr,
                                                         (S)p->goalRadius, (S*)p->nodes, p->parent, (S*)p->cost, p->goalList, nullptr, nullptr);
    MPTG_LAUNCHED(ctx);
    // rewire (:626-656)
    const size_t nSK = (size_t)nS * k;
    starRewireFlagKernel<S><<<(unsigned)((nSK + 127) / 128), 128, 0, st>>>(nS, nullptr, k, p->size, p->nnIdx, (const S*)p->nnDist, p->nnCnt, p->checked,
                                                                          (const S*)p->cost, p->flag);
    MPTG_LAUNCHED(ctx);
    uint32_t nR = 0;
    if (int rc = starSelect(p, p->flag, nSK, p->ids, &nR)) return rc;
    const uint32_t total = p->size + nS;
    if (nR) {
        starEdgeKernel<S><<<(nR + 127) / 128, 128, 0, st>>>(p->ids, nR, nullptr, k, D, false, (const S*)p->fresh, (const S*)p->nodes, p->nnIdx, (S*)p->eFrom,
                                                           (S*)p->eTo, nullptr);
        MPTG_LAUNCHED(ctx);
        if (int rc = mptg_link_batch_dev(p->geom, &p->space, p->eFrom, p->eTo, nR, p->linkStep, p->okEdge, nullptr)) return rc;
        // (bestKey / bestCand are in their "no offer" state: set at creation, restored by every push pass that follows offers -- ADVICE r1)
        MPTG_CUDA(ctx, cudaMemsetAsync(p->delta, 0, (size_t)total * sizeof(S), st));
        MPTG_CUDA(ctx, cudaMemsetAsync(p->result + 1, 0, 4, st));
        const uint32_t gr = (nR + 127) / 128;
        starRewireMinKernel<S><<<gr, 128, 0, st>>>(p->ids, p->okEdge, nR, nullptr, k, p->size, p->nnIdx, (const S*)p->nnDist, (const S*)p->cost, p->bestKey);
        MPTG_LAUNCHED(ctx);
        starRewirePickKernel<S><<<gr, 128, 0, st>>>(p->ids, p->okEdge, nR, nullptr, k, p->size, p->nnIdx, (const S*)p->nnDist, (const S*)p->cost, p->bestKey,
                                                    p->bestCand);
        MPTG_LAUNCHED(ctx);
        starRewireApplyKernel<S><<<gr, 128, 0, st>>>(p->ids, p->okEdge, nR, nullptr, k, p->size, p->nnIdx, (const S*)p->nnDist, (const S*)p->cost, p->bestCand,
                                                     p->parent, (S*)p->delta, p->result + 1);
        MPTG_LAUNCHED(ctx);
        starPushKernel<S><<<(total + 127) / 128, 128, 0, st>>>(p->size, nS, nullptr, p->parent, (const S*)p->delta, (S*)p->cost, p->result + 1, p->bestKey,
                                                               p->bestCand);
        MPTG_LAUNCHED(ctx);
    } else {
        MPTG_CUDA(ctx, cudaMemsetAsync(p->result + 1, 0, 4, st));
    }
    MPTG_CUDA(ctx, cudaMemsetAsync(p->result, 0xFF, 4, st));
    if (p->hasGoal) {
        starGoalKernel<S><<<1, 256, 0, st>>>(p->goalList, (const S*)p->cost, p->result);
        MPTG_LAUNCHED(ctx);
    }
    MPTG_CUDA(ctx, cudaMemcpyAsync(p->host + 1, p->result, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    uint32_t first = 0;
    if (int rc = mptg_knn_insert_dev(p->knn, p->fresh, nS, &first)) return rc;  // :619
    if (first != p->size) return fail(ctx, MPTG_ERR_CUDA, "mptg_prrtstar_wave: node numbering out of step");
    MPTG_CUDA(ctx, cudaStreamSynchronize(st));
    p->goalNode = p->host[1];
    p->rewires += p->host[2];
    p->size += nS;
    return MPTG_OK;
}

}  // namespace

extern "C" {

int mptg_prrtstar_create(mptg_ctx* ctx, mptg_geom* geom, const mptg_prrt_params* prm, double rewire_factor, mptg_prrtstar** out) {
    if (!ctx || !geom || !prm || !out || !spaceOk(prm->space) || prm->capacity == 0 || prm->max_wave == 0 || !(prm->range > 0) || !(rewire_factor > 0))
        return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_prrtstar_create: bad argument");
    if (geom->ctx != ctx || geom->scalar != prm->space->scalar || geom->D != spaceScalars(prm->space))
        return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_prrtstar_create: the geometry's states are not states of this space");
    if (geom->kind == MPTG_GEOM_MESH && !(prm->link_step > 0)) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_prrtstar_create: mesh geometries need link_step > 0");
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    auto* p = new mptg_prrtstar();
    p->ctx = ctx, p->geom = geom, p->space = *prm->space;
    p->D = spaceScalars(prm->space), p->scalar = prm->space->scalar, p->dims = spaceDimensions(*prm->space);
    p->range = prm->range, p->goalBias = prm->goal_bias, p->goalRadius = prm->goal_radius, p->linkStep = prm->link_step, p->rewireFactor = rewire_factor;
    p->seed = prm->seed, p->capacity = prm->capacity, p->maxWave = prm->max_wave, p->hasGoal = prm->goal_state != nullptr;
    const uint32_t kFull = p->scalar == MPTG_F32 ? starK<float>(rewire_factor, p->dims, p->capacity) : starK<double>(rewire_factor, p->dims, p->capacity);
    p->stride = kFull < MPTG_MAX_K ? kFull : MPTG_MAX_K;
    const size_t sb = (size_t)p->D * p->scalar, W = p->maxWave, sc = p->scalar;
    int rc = mptg_knn_create(ctx, prm->space, p->capacity, &p->knn);
    if (!rc) rc = uploadBounds(ctx, prm->space, prm->lo, prm->hi, &p->bounds);
    auto alloc = [&](auto** q, size_t bytes) {
        if (rc) return;
        cudaError_t e = cudaMalloc((void**)q, bytes ? bytes : 16);
        if (e != cudaSuccess) rc = fail(ctx, MPTG_ERR_OOM, "mptg_prrtstar_create: %s", cudaGetErrorString(e));
    };
    alloc(&p->nodes, (size_t)p->capacity * sb), alloc(&p->parent, (size_t)p->capacity * 4), alloc(&p->cost, (size_t)p->capacity * sc);
    alloc(&p->delta, (size_t)p->capacity * sc), alloc(&p->bestKey, (size_t)p->capacity * 8), alloc(&p->bestCand, (size_t)p->capacity * 4);
    alloc(&p->goalList, 65536 * 4);
    alloc(&p->samples, W * sb), alloc(&p->from, W * sb), alloc(&p->to, W * sb), alloc(&p->fresh, W * sb);
    alloc(&p->nearDist, W * sc), alloc(&p->dNew, W * sc), alloc(&p->dOf, W * sc), alloc(&p->defCost, W * sc);
    alloc(&p->nearIdx, W * 4), alloc(&p->nearCnt, W * 4), alloc(&p->sel, W * 4), alloc(&p->nearOf, W * 4), alloc(&p->limit, W * 4), alloc(&p->nearRank, W * 4);
    alloc(&p->nSel, 4), alloc(&p->result, 16), alloc(&p->cnt, 16);
    alloc(&p->alive, W), alloc(&p->okValid, W), alloc(&p->okLink, W), alloc(&p->keep, W);
    alloc(&p->nnCnt, W * 4);
    p->timing = getenv("MPTG_STAR_TIMING") != nullptr;
    if (const char* e = getenv("MPTG_STAR_QUEUED_MAX")) p->queuedMax = (uint32_t)atoi(e);  // 0: every wave with host synchronisation (comparisons, tests)
    if (!rc) rc = starAllocNeighbourBuffers(p);
    if (!rc && p->hasGoal) {
        alloc(&p->goal, sb);
        if (!rc) rc = uploadSync(ctx, p->goal, prm->goal_state, sb);
    }
    if (!rc) rc = memsetSync(ctx, p->goalList, 0, 65536 * 4);
    if (!rc) rc = memsetSync(ctx, p->bestKey, 0xFF, (size_t)p->capacity * 8);  // "no offer": the state every push pass restores
    if (!rc) rc = memsetSync(ctx, p->bestCand, 0xFF, (size_t)p->capacity * 4);
    if (!rc && cudaMallocHost((void**)&p->host, 4 * sizeof(uint32_t)) != cudaSuccess) rc = fail(ctx, MPTG_ERR_OOM, "mptg_prrtstar_create: pinned allocation failed");
    if (rc) {
        starFree(p);
        return rc;
    }
    *out = p;
    return MPTG_OK;
}

int mptg_prrtstar_set_rewire_radius(mptg_prrtstar* p, double r_rrg) {
    if (!p || !(r_rrg > 0)) return fail(p ? p->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_prrtstar_set_rewire_radius: bad argument");
    if (p->waves != 0) return fail(p->ctx, MPTG_ERR_BAD_ARG, "mptg_prrtstar_set_rewire_radius: must be called before the first wave");
    MPTG_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
    p->rRRG = r_rrg;
    if (p->stride != MPTG_MAX_K) {
        p->stride = MPTG_MAX_K;
        return starAllocNeighbourBuffers(p);
    }
    return MPTG_OK;
}

int mptg_prrtstar_destroy(mptg_prrtstar* p) {
    if (!p) return MPTG_OK;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    if (p->timing && p->queuedWaves)
        fprintf(stderr, "[mptg prrt*] %llu queued waves: %.1f us issuing + %.1f us waiting per wave, %.1f kernel launches per wave\n",
                (unsigned long long)p->queuedWaves, p->issueUs / p->queuedWaves, p->waitUs / p->queuedWaves, (double)p->queuedLaunches / p->queuedWaves);
    starFree(p);
    return MPTG_OK;
}

int mptg_prrtstar_add_start(mptg_prrtstar* p, const void* state) {
    if (!p || !state) return fail(p ? p->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_prrtstar_add_start: bad argument");
    if (p->size >= p->capacity) return fail(p->ctx, MPTG_ERR_CAPACITY, "mptg_prrtstar_add_start: tree is full");
    mptg_ctx* ctx = p->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t sb = (size_t)p->D * p->scalar;
    const uint32_t none = MPTG_NO_INDEX;
    const double zero = 0.0;  // all-zero bits in either scalar type
    if (int rc = uploadSync(ctx, (char*)p->nodes + (size_t)p->size * sb, state, sb)) return rc;
    if (int rc = uploadSync(ctx, p->parent + p->size, &none, sizeof none)) return rc;
    if (int rc = uploadSync(ctx, (char*)p->cost + (size_t)p->size * p->scalar, &zero, p->scalar)) return rc;
    uint32_t first = 0;
    if (int rc = mptg_knn_insert(p->knn, state, 1, &first)) return rc;
    ++p->size;
    return MPTG_OK;
}

int mptg_prrtstar_wave(mptg_prrtstar* p, uint32_t n_samples, uint32_t* size_out, uint32_t* goal_node_out) {
    if (!p || n_samples == 0 || n_samples > p->maxWave) return fail(p ? p->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_prrtstar_wave: bad argument");
    if (p->size == 0) return fail(p->ctx, MPTG_ERR_BAD_ARG, "mptg_prrtstar_wave: there are no valid initial states");
    MPTG_CUDA(p->ctx, cudaSetDevice(p->ctx->device));
    const int rc = p->scalar == MPTG_F32 ? starWaveT<float>(p, n_samples) : starWaveT<double>(p, n_samples);
    if (size_out) *size_out = p->size;
    if (goal_node_out) *goal_node_out = p->goalNode;
    return rc;
}

uint32_t mptg_prrtstar_size(const mptg_prrtstar* p) { return p ? p->size : 0; }
uint64_t mptg_prrtstar_samples_drawn(const mptg_prrtstar* p) { return p ? p->drawn : 0; }
uint64_t mptg_prrtstar_rewires(const mptg_prrtstar* p) { return p ? p->rewires : 0; }

int mptg_prrtstar_get_tree(mptg_prrtstar* p, uint32_t first, uint32_t count, void* states_out, uint32_t* parents_out, void* costs_out) {
    if (!p || (uint64_t)first + count > p->size) return fail(p ? p->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_prrtstar_get_tree: bad range");
    if (count == 0) return MPTG_OK;
    mptg_ctx* ctx = p->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t sb = (size_t)p->D * p->scalar;
    if (states_out) MPTG_CUDA(ctx, cudaMemcpyAsync(states_out, (char*)p->nodes + first * sb, count * sb, cudaMemcpyDeviceToHost, ctx->stream));
    if (parents_out) MPTG_CUDA(ctx, cudaMemcpyAsync(parents_out, p->parent + first, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (costs_out)
        MPTG_CUDA(ctx, cudaMemcpyAsync(costs_out, (char*)p->cost + (size_t)first * p->scalar, (size_t)count * p->scalar, cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPTG_OK;
}

}  // extern "C"
