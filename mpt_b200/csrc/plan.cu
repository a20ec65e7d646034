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
    if (int rc = mptg_valid_batch_dev(p->geom, p->to, W, p->okValid)) return rc;
    if (int rc = mptg_link_batch_dev(p->geom, &p->space, p->from, p->to, W, p->linkStep, p->okLink)) return rc;
    prrtFlagKernel<<<grid, 128, 0, ctx->stream>>>(p->alive, p->okValid, p->okLink, W, p->keep);
    MPTG_LAUNCHED(ctx);
    size_t bytes = p->selBytes;
    MPTG_CUDA(ctx, cub::DeviceSelect::Flagged(p->selTemp, bytes, thrust::counting_iterator<uint32_t>(0), p->keep, p->sel, p->nSel, (int)W,
                                              ctx->stream));
    MPTG_LAUNCHED(ctx);
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
    const uint32_t init[2] = {0u, MPTG_NO_INDEX};
    if (int rc = uploadSync(p->ctx, p->result, init, sizeof init)) return rc;
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
