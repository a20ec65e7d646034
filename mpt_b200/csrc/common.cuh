// common.cuh -- context, error handling and device buffers shared by the libmptg translation units.
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mptg/mptg.h"

struct mptg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    int smCount = 148;
    // pinned staging for small host<->device transfers (counts, flags)
    void* pinned = nullptr;
    size_t pinnedBytes = 0;
    // device scratch reused by the host-pointer entry points (grown on demand)
    void* scratch[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t scratchBytes[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
};

namespace mptg {

extern thread_local std::string g_lastError;  // for failures without a context

inline int fail(mptg_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    g_lastError = buf;
    return code;
}

#define MPTG_CUDA(ctx, expr)                                                                             \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return ::mptg::fail((ctx), e_ == cudaErrorMemoryAllocation ? MPTG_ERR_OOM : MPTG_ERR_CUDA,   \
                                "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// after a kernel launch: count it and surface launch errors
#define MPTG_LAUNCHED(ctx)                                                                               \
    do {                                                                                                 \
        ++(ctx)->launches;                                                                               \
        cudaError_t e_ = cudaGetLastError();                                                             \
        if (e_ != cudaSuccess)                                                                           \
            return ::mptg::fail((ctx), MPTG_ERR_CUDA, "kernel launch failed: %s (%s:%d)",               \
                                cudaGetErrorString(e_), __FILE__, __LINE__);                             \
    } while (0)

// Grow-only device scratch slot.
inline int scratch(mptg_ctx* ctx, int slot, size_t bytes, void** out) {
    if (ctx->scratchBytes[slot] < bytes) {
        if (ctx->scratch[slot]) {
            MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            MPTG_CUDA(ctx, cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr;
            ctx->scratchBytes[slot] = 0;
        }
        size_t want = bytes + bytes / 4 + 256;
        MPTG_CUDA(ctx, cudaMalloc(&ctx->scratch[slot], want));
        ctx->scratchBytes[slot] = want;
    }
    *out = ctx->scratch[slot];
    return MPTG_OK;
}

// Host -> device upload that is complete when it returns.  cudaMemcpy from pageable memory may return
// while the DMA is still in flight, and the context stream is non-blocking (not ordered after the
// legacy default stream), so every upload goes through the context stream and is synchronised.
inline int uploadSync(mptg_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return MPTG_OK;
    MPTG_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPTG_OK;
}
inline int memsetSync(mptg_ctx* ctx, void* dst, int value, size_t bytes) {
    MPTG_CUDA(ctx, cudaMemsetAsync(dst, value, bytes, ctx->stream));
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPTG_OK;
}

inline int spaceScalars(const mptg_space_desc* s) {
    if (!s || s->n_parts < 1 || s->n_parts > MPTG_MAX_PARTS) return -1;
    if (s->scalar != MPTG_F32 && s->scalar != MPTG_F64) return -1;
    int n = 0;
    for (int i = 0; i < s->n_parts; ++i) {
        const mptg_space_part& p = s->part[i];
        if (p.kind == MPTG_PART_SO3) n += 4;
        else if ((p.kind == MPTG_PART_LP || p.kind == MPTG_PART_SO2) && p.dim >= 1 && (p.p == 0 || p.p == 1 || p.p == 2)) n += p.dim;
        else return -1;
    }
    return n <= MPTG_MAX_SCALARS ? n : -1;
}

}  // namespace mptg
