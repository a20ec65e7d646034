// ctx.cu -- context lifetime and small queries of the C ABI.
#include "common.cuh"

namespace mptg {
thread_local std::string g_lastError;
}
using namespace mptg;

extern "C" {

int mptg_abi_version(void) { return MPTG_ABI_VERSION; }

int mptg_ctx_create(int device, mptg_ctx** out) {
    if (!out) return fail(nullptr, MPTG_ERR_BAD_ARG, "mptg_ctx_create: out is NULL");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, MPTG_ERR_CUDA, "mptg_ctx_create: no CUDA device (%s); libmptg has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0) {
        e = cudaGetDevice(&device);
        if (e != cudaSuccess) return fail(nullptr, MPTG_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
    }
    if (device >= count) return fail(nullptr, MPTG_ERR_BAD_ARG, "mptg_ctx_create: device %d of %d", device, count);
    auto* ctx = new mptg_ctx();
    ctx->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, MPTG_ERR_CUDA, "mptg_ctx_create: %s", cudaGetErrorString(e));
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        ctx->smCount = prop.multiProcessorCount;
        if (prop.major < 10) {
            cudaStreamDestroy(ctx->stream);
            delete ctx;
            return fail(nullptr, MPTG_ERR_CUDA, "mptg_ctx_create: device %d is sm_%d%d; libmptg is built for sm_100a only", device,
                        prop.major, prop.minor);
        }
    }
    ctx->pinnedBytes = 1 << 16;
    if ((e = cudaMallocHost(&ctx->pinned, ctx->pinnedBytes)) != cudaSuccess) {
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return fail(nullptr, MPTG_ERR_OOM, "mptg_ctx_create: cudaMallocHost: %s", cudaGetErrorString(e));
    }
    *out = ctx;
    return MPTG_OK;
}

int mptg_ctx_destroy(mptg_ctx* ctx) {
    if (!ctx) return MPTG_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (void* p : ctx->scratch)
        if (p) cudaFree(p);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return MPTG_OK;
}

int mptg_sync(mptg_ctx* ctx) {
    if (!ctx) return fail(nullptr, MPTG_ERR_BAD_ARG, "mptg_sync: null context");
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPTG_OK;
}

const char* mptg_last_error(const mptg_ctx* ctx) { return ctx ? ctx->err.c_str() : g_lastError.c_str(); }
void* mptg_ctx_stream(mptg_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t mptg_ctx_launch_count(const mptg_ctx* ctx) { return ctx ? ctx->launches : 0; }
int mptg_space_scalars(const mptg_space_desc* space) { return spaceScalars(space); }
int mptg_space_dimensions(const mptg_space_desc* space) {
    if (spaceScalars(space) < 0) return -1;
    int n = 0;
    for (int i = 0; i < space->n_parts; ++i) n += space->part[i].kind == MPTG_PART_SO3 ? 3 : space->part[i].dim;
    return n;
}
}
