// metric.cu -- batched Space::distance, interpolate and steer (SURVEY.md section 8 rows a4, a5 and
// the steer stage of Appendix C).  One thread per item; these are memory-light helper stages, the
// arithmetic itself is in space.cuh and is shared with the kNN and edge kernels.
#include "common.cuh"
#include "../../include/mptg/mptg_space.h"

namespace mptg {

template <typename S>
__global__ void distanceKernel(DevSpace<S> sp, const S* a, const S* b, uint32_t n, S* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const S* pa = a + (size_t)i * sp.D;
    const S* pb = b + (size_t)i * sp.D;
    out[i] = dev::distance<S>(sp, [&](int c) { return pa[c]; }, [&](int c) { return pb[c]; });
}

template <typename S>
__global__ void interpolateKernel(DevSpace<S> sp, const S* a, const S* b, const S* t, uint32_t n, S* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dev::interpolate<S>(sp, a + (size_t)i * sp.D, b + (size_t)i * sp.D, t[i], out + (size_t)i * sp.D);
}

// impl/prrt/prrt.hpp:430-434 and impl/prrt_star/prrt_star.hpp:529-536
template <typename S>
__global__ void steerKernel(DevSpace<S> sp, const S* near, const S* sample, const S* d, uint32_t n, S range, S* out,
                            S* distOut) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const S* pn = near + (size_t)i * sp.D;
    const S* ps = sample + (size_t)i * sp.D;
    S* po = out + (size_t)i * sp.D;
    if (d[i] > range) {
        dev::interpolate<S>(sp, pn, ps, fp::div_(range, d[i]), po);
    } else {
        for (int c = 0; c < sp.D; ++c) po[c] = ps[c];
    }
    if (distOut) distOut[i] = dev::distance<S>(sp, [&](int c) { return pn[c]; }, [&](int c) { return po[c]; });
}

}  // namespace mptg

using namespace mptg;

namespace {

// copy `count` host blocks to consecutive regions of scratch slot 0; returns device pointers
struct Staged {
    void* p[4];
};

int stageIn(mptg_ctx* ctx, int n, const void* const* host, const size_t* bytes, size_t outBytes, Staged* st, void** outDev) {
    size_t total = 0;
    for (int i = 0; i < n; ++i) total += (bytes[i] + 255) & ~(size_t)255;
    void* base;
    int rc = scratch(ctx, 0, total, &base);
    if (rc) return rc;
    rc = scratch(ctx, 1, outBytes, outDev);
    if (rc) return rc;
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
        st->p[i] = (char*)base + off;
        if (bytes[i]) MPTG_CUDA(ctx, cudaMemcpyAsync(st->p[i], host[i], bytes[i], cudaMemcpyHostToDevice, ctx->stream));
        off += (bytes[i] + 255) & ~(size_t)255;
    }
    return MPTG_OK;
}

}  // namespace

extern "C" {

int mptg_distance_batch(mptg_ctx* ctx, const mptg_space_desc* space, const void* a, const void* b, uint32_t n,
                        void* out) {
    const int D = spaceScalars(space);
    if (!ctx || D <= 0 || (n && (!a || !b || !out))) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_distance_batch: bad argument");
    if (n == 0) return MPTG_OK;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t sb = (size_t)n * D * space->scalar;
    const void* host[2] = {a, b};
    const size_t bytes[2] = {sb, sb};
    Staged st;
    void* dOut;
    int rc = stageIn(ctx, 2, host, bytes, (size_t)n * space->scalar, &st, &dOut);
    if (rc) return rc;
    const dim3 grid((n + 127) / 128), block(128);
    if (space->scalar == MPTG_F32)
        distanceKernel<float><<<grid, block, 0, ctx->stream>>>(makeDevSpace<float>(*space), (const float*)st.p[0], (const float*)st.p[1], n, (float*)dOut);
    else
        distanceKernel<double><<<grid, block, 0, ctx->stream>>>(makeDevSpace<double>(*space), (const double*)st.p[0], (const double*)st.p[1], n, (double*)dOut);
    MPTG_LAUNCHED(ctx);
    MPTG_CUDA(ctx, cudaMemcpyAsync(out, dOut, (size_t)n * space->scalar, cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPTG_OK;
}

int mptg_interpolate_batch(mptg_ctx* ctx, const mptg_space_desc* space, const void* a, const void* b, const void* t,
                           uint32_t n, void* out) {
    const int D = spaceScalars(space);
    if (!ctx || D <= 0 || (n && (!a || !b || !t || !out))) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_interpolate_batch: bad argument");
    if (n == 0) return MPTG_OK;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t sb = (size_t)n * D * space->scalar;
    const void* host[3] = {a, b, t};
    const size_t bytes[3] = {sb, sb, (size_t)n * space->scalar};
    Staged st;
    void* dOut;
    int rc = stageIn(ctx, 3, host, bytes, sb, &st, &dOut);
    if (rc) return rc;
    const dim3 grid((n + 127) / 128), block(128);
    if (space->scalar == MPTG_F32)
        interpolateKernel<float><<<grid, block, 0, ctx->stream>>>(makeDevSpace<float>(*space), (const float*)st.p[0], (const float*)st.p[1], (const float*)st.p[2], n, (float*)dOut);
    else
        interpolateKernel<double><<<grid, block, 0, ctx->stream>>>(makeDevSpace<double>(*space), (const double*)st.p[0], (const double*)st.p[1], (const double*)st.p[2], n, (double*)dOut);
    MPTG_LAUNCHED(ctx);
    MPTG_CUDA(ctx, cudaMemcpyAsync(out, dOut, sb, cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPTG_OK;
}

int mptg_steer_batch(mptg_ctx* ctx, const mptg_space_desc* space, const void* near, const void* sample, const void* d,
                     uint32_t n, double range, void* out, void* distOut) {
    const int D = spaceScalars(space);
    if (!ctx || D <= 0 || (n && (!near || !sample || !d || !out))) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_steer_batch: bad argument");
    if (n == 0) return MPTG_OK;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t sb = (size_t)n * D * space->scalar, db = (size_t)n * space->scalar;
    const void* host[3] = {near, sample, d};
    const size_t bytes[3] = {sb, sb, db};
    Staged st;
    void* dOut;
    int rc = stageIn(ctx, 3, host, bytes, sb + db, &st, &dOut);
    if (rc) return rc;
    void* dDist = (char*)dOut + sb;
    const dim3 grid((n + 127) / 128), block(128);
    if (space->scalar == MPTG_F32)
        steerKernel<float><<<grid, block, 0, ctx->stream>>>(makeDevSpace<float>(*space), (const float*)st.p[0], (const float*)st.p[1], (const float*)st.p[2], n, (float)range, (float*)dOut, distOut ? (float*)dDist : nullptr);
    else
        steerKernel<double><<<grid, block, 0, ctx->stream>>>(makeDevSpace<double>(*space), (const double*)st.p[0], (const double*)st.p[1], (const double*)st.p[2], n, range, (double*)dOut, distOut ? (double*)dDist : nullptr);
    MPTG_LAUNCHED(ctx);
    MPTG_CUDA(ctx, cudaMemcpyAsync(out, dOut, sb, cudaMemcpyDeviceToHost, ctx->stream));
    if (distOut) MPTG_CUDA(ctx, cudaMemcpyAsync(distOut, dDist, db, cudaMemcpyDeviceToHost, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPTG_OK;
}
}
