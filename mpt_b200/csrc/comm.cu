// comm.cu -- tree-sharded nearest-neighbour search across the GPUs of one box (north_star: "kNN
// shards tree points across GPUs with a local top-k merged by NCCL over NVLink").  One process per GPU; every rank
// holds a SPATIAL shard of the tree points (mptg_knn_insert_ids: stored with their global indices) and the whole query
// wave; rank r owns the results of a contiguous slice of the wave.
//
// After inserting, all ranks call mptg_knn_shard_sync: every shard indexes what it stores and the bounding boxes of the
// (up to 32) children of each shard's top node are all-gathered -- a few KB that let every rank bound the distance of a
// query to EVERY shard by itself.
//
// Protocol of one wave (everything is enqueued on the context stream; no host synchronisation between the steps):
//   1. root bounds     lb[g][q] = lower bound of query q to the bounding box of shard g, for every g, computed locally;
//                      home(q) = argmin_g lb[g][q]
//   2. home search     the home rank searches q to the end: its k-th distance is an upper bound B[q] of the k-th
//                      distance over the union (a tight one: most of a query's neighbours live at home)  all-reduce(min), 4 Q bytes
//   3. bounded search  every other rank g with lb[g][q] <= B[q] searches q with radius B[q] (inclusive: ties are
//                      decided by the global index in the merge); all other (q, g) pairs are provably empty
//   4. exchange        the [k] candidate rows of every query go to the query's owner                     send/recv, 8 k Q / G bytes per pair
//   5. merge           k best of the G rows by (distance, global index): identical to the single-GPU result.
// r1 dealt the points round-robin, which made every GPU search every query (1.5x at 8 GPUs); here a query is searched
// in its home shard and in the few shards its k-th ball reaches.  (Tried and dropped: a quick bound from the first k
// candidates of the home shard instead of the full home search -- it removes the latency-bound home pass (8,192 searches
// per GPU at 8 GPUs), but its bound is ~1.5x the true radius, remote shards hold fewer than k points inside it and
// therefore never tighten it, and the bounded pass tripled: 0.78 ms per wave against 0.64.)
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy the process already has -- torch's -- or the system
// one), so libmptg.so keeps loading on machines without it; every failure returns MPTG_ERR_NCCL.
#include <dlfcn.h>
#include <nccl.h>

#include "knn_shard.cuh"

namespace mptg {
namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

const NcclApi* ncclApi(std::string& why) {
    static NcclApi api;
    static bool tried = false;
    static std::string err;
    if (!tried) {
        tried = true;
        const char* env = getenv("MPTG_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) {
            err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "");
        } else {
#define MPTG_SYM(field, name)                                      \
    api.field = (decltype(api.field))dlsym(api.lib, name);        \
    if (!api.field) err = std::string("libnccl: missing ") + name;
            MPTG_SYM(GetUniqueId, "ncclGetUniqueId")
            MPTG_SYM(CommInitRank, "ncclCommInitRank")
            MPTG_SYM(CommDestroy, "ncclCommDestroy")
            MPTG_SYM(AllGather, "ncclAllGather")
            MPTG_SYM(AllReduce, "ncclAllReduce")
            MPTG_SYM(Send, "ncclSend")
            MPTG_SYM(Recv, "ncclRecv")
            MPTG_SYM(GroupStart, "ncclGroupStart")
            MPTG_SYM(GroupEnd, "ncclGroupEnd")
            MPTG_SYM(GetErrorString, "ncclGetErrorString")
#undef MPTG_SYM
        }
    }
    why = err;
    return err.empty() ? &api : nullptr;
}

#define MPTG_NCCL(ctx, api, expr)                                                                                              \
    do {                                                                                                                       \
        ncclResult_t r_ = (expr);                                                                                              \
        if (r_ != ncclSuccess) return fail((ctx), MPTG_ERR_NCCL, "%s failed: %s (%s:%d)", #expr, (api)->GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

// ---- small kernels of the protocol
template <typename S>
__global__ void shardFillKernel(uint32_t* idx, S* dist, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = MPTG_NO_INDEX, dist[i] = (S)INFINITY;
}
// B[q] = k-th distance of the home search (rows of the other queries still hold +inf)
template <typename S>
__global__ void shardBoundKernel(const S* __restrict__ dist, uint32_t Q, uint32_t k, S* __restrict__ bound) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < Q) bound[q] = dist[(size_t)q * k + (k - 1)];
}
// second pass: the queries of other homes whose ball B[q] reaches this shard, with radius B[q]
template <typename S>
__global__ void shardCapKernel(const float* __restrict__ lbMine, const uint8_t* __restrict__ home, const S* __restrict__ bound, int rank, uint32_t Q,
                               S* __restrict__ cap) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const S b = bound[q];
    cap[q] = ((int)home[q] != rank && (double)lbMine[q] <= (double)b) ? b : S(-1);
}

inline void sliceOf(uint32_t n, int world, int rank, uint32_t& first, uint32_t& count) {
    const uint32_t base = n / (uint32_t)world, rem = n % (uint32_t)world;
    first = (uint32_t)rank * base + ((uint32_t)rank < rem ? (uint32_t)rank : rem);
    count = base + ((uint32_t)rank < rem ? 1u : 0u);
}

}  // namespace
}  // namespace mptg

using namespace mptg;

struct mptg_comm {
    mptg_ctx* ctx = nullptr;
    const NcclApi* api = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    void* buf = nullptr;  // work space of one wave
    size_t bufBytes = 0;
    uint64_t waves = 0;
    // synchronised top boxes of every shard (mptg_knn_shard_sync)
    void* peerBox = nullptr;       // [world][2 D][32] scalars
    void* shardBox = nullptr;      // [world][2 D] scalars: one box per shard
    uint32_t* peerN = nullptr;     // [world]
    size_t peerBlockBytes = 0;     // bytes of one shard's block the three arrays above are sized for
    mptg_knn* syncedShard = nullptr;
    uint32_t syncedSize = 0;
    uint64_t syncedBuilds = 0;
    bool timing = false;  // MPTG_COMM_TIMING=1
    cudaEvent_t ev[2][7];
    double sumMs[6] = {0, 0, 0, 0, 0, 0};
    uint64_t timedWaves = 0;
};

extern "C" {

int mptg_comm_unique_id(void* idOut) {
    if (!idOut) return fail(nullptr, MPTG_ERR_BAD_ARG, "mptg_comm_unique_id: null output");
    std::string why;
    const NcclApi* api = ncclApi(why);
    if (!api) return fail(nullptr, MPTG_ERR_NCCL, "mptg_comm_unique_id: %s", why.c_str());
    static_assert(sizeof(ncclUniqueId) == MPTG_UNIQUE_ID_BYTES, "MPTG_UNIQUE_ID_BYTES must equal sizeof(ncclUniqueId)");
    ncclUniqueId id;
    MPTG_NCCL(nullptr, api, api->GetUniqueId(&id));
    memcpy(idOut, &id, sizeof id);
    return MPTG_OK;
}

int mptg_comm_init(mptg_ctx* ctx, const void* uniqueId, int rank, int world, mptg_comm** out) {
    if (!ctx || !uniqueId || !out || world < 1 || world > 255 || rank < 0 || rank >= world)
        return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_comm_init: bad argument (1 <= world <= 255, 0 <= rank < world)");
    std::string why;
    const NcclApi* api = ncclApi(why);
    if (!api) return fail(ctx, MPTG_ERR_NCCL, "mptg_comm_init: %s", why.c_str());
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, uniqueId, sizeof id);
    auto* c = new mptg_comm();
    c->ctx = ctx, c->api = api, c->rank = rank, c->world = world;
    ncclResult_t r = api->CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        delete c;
        return fail(ctx, MPTG_ERR_NCCL, "ncclCommInitRank(rank %d of %d) failed: %s", rank, world, api->GetErrorString(r));
    }
    if (const char* e = getenv("MPTG_COMM_TIMING")) c->timing = e[0] == '1';
    if (c->timing)
        for (int b = 0; b < 2; ++b)
            for (int i = 0; i < 7; ++i) cudaEventCreate(&c->ev[b][i]);
    *out = c;
    return MPTG_OK;
}

int mptg_comm_destroy(mptg_comm* c) {
    if (!c) return MPTG_OK;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    if (c->timing && c->timedWaves)
        fprintf(stderr, "[mptg comm] rank %d of %d, %llu waves, ms per wave: root bounds + fill %.3f | home search %.3f | bound all-reduce %.3f | "
                        "bounded search %.3f | exchange %.3f | merge %.3f\n",
                c->rank, c->world, (unsigned long long)c->timedWaves, c->sumMs[0] / c->timedWaves, c->sumMs[1] / c->timedWaves, c->sumMs[2] / c->timedWaves,
                c->sumMs[3] / c->timedWaves, c->sumMs[4] / c->timedWaves, c->sumMs[5] / c->timedWaves);
    if (c->comm) c->api->CommDestroy(c->comm);
    cudaFree(c->buf);
    cudaFree(c->peerBox);
    cudaFree(c->shardBox);
    cudaFree(c->peerN);
    delete c;
    return MPTG_OK;
}

int mptg_comm_rank(const mptg_comm* c) { return c ? c->rank : -1; }
int mptg_comm_world(const mptg_comm* c) { return c ? c->world : 0; }

int mptg_comm_slice(const mptg_comm* c, uint32_t n, uint32_t* firstOut, uint32_t* countOut) {
    if (!c || !firstOut || !countOut) return fail(c ? c->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_comm_slice: bad argument");
    sliceOf(n, c->world, c->rank, *firstOut, *countOut);
    return MPTG_OK;
}

}  // extern "C"

namespace {

template <typename S>
int queryShardedT(mptg_comm* c, mptg_knn* shard, const S* queries, uint32_t Q, uint32_t k, double radius, uint32_t* idxOut, S* distOut,
                  uint32_t* countOut) {
    mptg_ctx* ctx = c->ctx;
    const NcclApi* api = c->api;
    cudaStream_t st = ctx->stream;
    const int G = c->world, me = c->rank;
    uint32_t myFirst, myCount;
    sliceOf(Q, G, me, myFirst, myCount);
    const uint32_t maxSlice = Q / (uint32_t)G + 1u;
    // work space
    size_t bytes = 0;
    auto take = [&](size_t b) {
        const size_t o = bytes;
        bytes += (b + 255) & ~(size_t)255;
        return o;
    };
    const size_t oLbMine = take((size_t)Q * 4), oHome = take(Q), oCap = take((size_t)Q * sizeof(S)), oBound = take((size_t)Q * sizeof(S)),
                 oIdx = take((size_t)Q * k * 4), oDist = take((size_t)Q * k * sizeof(S)), oRecvI = take((size_t)G * maxSlice * k * 4),
                 oRecvD = take((size_t)G * maxSlice * k * sizeof(S));
    if (c->bufBytes < bytes) {
        MPTG_CUDA(ctx, cudaStreamSynchronize(st));
        if (c->buf) MPTG_CUDA(ctx, cudaFree(c->buf));
        c->buf = nullptr, c->bufBytes = 0;
        MPTG_CUDA(ctx, cudaMalloc(&c->buf, bytes + bytes / 4));
        c->bufBytes = bytes + bytes / 4;
    }
    char* W = (char*)c->buf;
    float* lbMine = (float*)(W + oLbMine);
    uint8_t* home = (uint8_t*)(W + oHome);
    S* cap = (S*)(W + oCap);
    S* bound = (S*)(W + oBound);
    uint32_t* idx = (uint32_t*)(W + oIdx);
    S* dist = (S*)(W + oDist);
    uint32_t* recvI = (uint32_t*)(W + oRecvI);
    S* recvD = (S*)(W + oRecvD);
    const ncclDataType_t ncclS = sizeof(S) == 4 ? ncclFloat32 : ncclFloat64;
    const uint32_t g256 = (Q + 255) / 256;

    // MPTG_COMM_TIMING=1: CUDA events between the steps, averaged over the waves and printed when the communicator is destroyed
#define MARK(i)                                                        \
    do {                                                               \
        if (c->timing) cudaEventRecord(c->ev[(c->waves & 1)][i], st); \
    } while (0)
    if (c->timing && c->waves >= 12) {  // fold the events of the wave before last (long finished) into the sums; the first waves set up connections
        cudaEvent_t* e = c->ev[c->waves & 1];
        if (cudaEventSynchronize(e[6]) == cudaSuccess) {
            for (int i = 0; i < 6; ++i) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, e[i], e[i + 1]) == cudaSuccess) c->sumMs[i] += ms;
            }
            ++c->timedWaves;
        }
    }
    MARK(0);
    // 1. bounds of every query to every shard, home shards, output rows preset to "nothing found"
    if (int rc = knnShardRootAll(shard, c->shardBox, c->peerN, G, me, queries, Q, lbMine, home, cap)) return rc;
    shardFillKernel<S><<<(unsigned)(((size_t)Q * k + 255) / 256), 256, 0, st>>>(idx, dist, (size_t)Q * k);
    MPTG_LAUNCHED(ctx);
    MARK(1);
    // 2. home search
    if (int rc = knnShardQuery(shard, queries, Q, k, radius, cap, idx, dist)) return rc;
    MARK(2);
    if (G > 1) {
        shardBoundKernel<S><<<g256, 256, 0, st>>>(dist, Q, k, bound);
        MPTG_LAUNCHED(ctx);
        MPTG_NCCL(ctx, api, api->AllReduce(bound, bound, Q, ncclS, ncclMin, c->comm, st));
    }
    MARK(3);
    if (G > 1) {
        // 3. bounded search of the queries of other homes that reach this shard
        shardCapKernel<S><<<g256, 256, 0, st>>>(lbMine, home, bound, me, Q, cap);
        MPTG_LAUNCHED(ctx);
        if (int rc = knnShardQuery(shard, queries, Q, k, radius, cap, idx, dist)) return rc;
    }
    MARK(4);
    if (G > 1) {
        // 4. candidate rows to the owners of the queries
        MPTG_NCCL(ctx, api, api->GroupStart());
        for (int p = 0; p < G; ++p) {
            uint32_t f, n;
            sliceOf(Q, G, p, f, n);
            if (p == me) continue;
            if (n) {
                MPTG_NCCL(ctx, api, api->Send(idx + (size_t)f * k, (size_t)n * k, ncclUint32, p, c->comm, st));
                MPTG_NCCL(ctx, api, api->Send(dist + (size_t)f * k, (size_t)n * k, ncclS, p, c->comm, st));
            }
            if (myCount) {
                MPTG_NCCL(ctx, api, api->Recv(recvI + (size_t)p * myCount * k, (size_t)myCount * k, ncclUint32, p, c->comm, st));
                MPTG_NCCL(ctx, api, api->Recv(recvD + (size_t)p * myCount * k, (size_t)myCount * k, ncclS, p, c->comm, st));
            }
        }
        MPTG_NCCL(ctx, api, api->GroupEnd());
    }
    MARK(5);
    if (myCount) {
        MPTG_CUDA(ctx, cudaMemcpyAsync(recvI + (size_t)me * myCount * k, idx + (size_t)myFirst * k, (size_t)myCount * k * 4, cudaMemcpyDeviceToDevice, st));
        MPTG_CUDA(ctx, cudaMemcpyAsync(recvD + (size_t)me * myCount * k, dist + (size_t)myFirst * k, (size_t)myCount * k * sizeof(S), cudaMemcpyDeviceToDevice, st));
        // 5. merge by (distance, global index)
        if (int rc = mptg_knn_merge_dev(ctx, sizeof(S) == 4 ? MPTG_F32 : MPTG_F64, (uint32_t)G, myCount, k, recvI, recvD, idxOut, distOut, countOut)) return rc;
    }
    MARK(6);
    ++c->waves;
    return MPTG_OK;
}
#undef MARK

}  // namespace

extern "C" {

int mptg_knn_shard_sync(mptg_comm* c, mptg_knn* shard) {
    if (!c || !shard) return fail(c ? c->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_knn_shard_sync: bad argument");
    mptg_ctx* ctx = c->ctx;
    if (knnShardCtx(shard) != ctx) return fail(ctx, MPTG_ERR_BAD_ARG, "mptg_knn_shard_sync: the shard belongs to another context");
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* topBox = nullptr;
    uint32_t nTop = 0;
    if (int rc = knnShardIndexAll(shard, &topBox, &nTop)) return rc;
    const size_t blockBytes = (size_t)2 * knnShardScalars(shard) * 32 * knnShardScalar(shard);
    if (c->peerBlockBytes < blockBytes) {  // a communicator may serve structures of different spaces, one after the other
        MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(c->peerBox), cudaFree(c->shardBox), cudaFree(c->peerN);
        c->peerBox = c->shardBox = nullptr, c->peerN = nullptr, c->peerBlockBytes = 0;
        MPTG_CUDA(ctx, cudaMalloc(&c->peerBox, blockBytes * c->world));
        MPTG_CUDA(ctx, cudaMalloc(&c->shardBox, blockBytes / 32 * c->world));
        MPTG_CUDA(ctx, cudaMalloc(&c->peerN, sizeof(uint32_t) * c->world));
        c->peerBlockBytes = blockBytes;
    }
    char* mine = (char*)c->peerBox + blockBytes * c->rank;
    if (nTop) MPTG_CUDA(ctx, cudaMemcpyAsync(mine, topBox, blockBytes, cudaMemcpyDeviceToDevice, ctx->stream));
    MPTG_CUDA(ctx, cudaMemcpyAsync(c->peerN + c->rank, &nTop, sizeof nTop, cudaMemcpyHostToDevice, ctx->stream));
    if (c->world > 1) {  // in place: every rank's block already sits at its own offset
        MPTG_NCCL(ctx, c->api, c->api->AllGather(mine, c->peerBox, blockBytes, ncclUint8, c->comm, ctx->stream));
        MPTG_NCCL(ctx, c->api, c->api->AllGather(c->peerN + c->rank, c->peerN, 1, ncclUint32, c->comm, ctx->stream));
    }
    if (int rc = knnShardUnion(shard, c->peerBox, c->peerN, c->world, c->shardBox)) return rc;
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // nTop lives on this frame
    c->syncedShard = shard;
    c->syncedSize = knnShardSize(shard);
    c->syncedBuilds = knnShardBuilds(shard);
    return MPTG_OK;
}

int mptg_knn_query_sharded_dev(mptg_comm* c, mptg_knn* shard, const void* queriesDev, uint32_t Q, uint32_t k, double radius,
                               uint32_t* idxOutDev, void* distOutDev, uint32_t* countOutDev) {
    if (!c || !shard || (!queriesDev && Q) || !idxOutDev || !distOutDev)
        return fail(c ? c->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_knn_query_sharded: bad argument");
    if (knnShardCtx(shard) != c->ctx) return fail(c->ctx, MPTG_ERR_BAD_ARG, "mptg_knn_query_sharded: the shard belongs to another context");
    if (k == 0 || k > MPTG_MAX_K) return fail(c->ctx, MPTG_ERR_BAD_ARG, "mptg_knn_query_sharded: k=%u out of range 1..%d", k, MPTG_MAX_K);
    if (c->syncedShard != shard || c->syncedSize != knnShardSize(shard) || c->syncedBuilds != knnShardBuilds(shard))
        return fail(c->ctx, MPTG_ERR_BAD_ARG, "mptg_knn_query_sharded: the shard has changed since the last mptg_knn_shard_sync (call it on every rank after inserting)");
    if (Q == 0) return MPTG_OK;
    MPTG_CUDA(c->ctx, cudaSetDevice(c->ctx->device));
    return knnShardScalar(shard) == MPTG_F32
               ? queryShardedT<float>(c, shard, (const float*)queriesDev, Q, k, radius, idxOutDev, (float*)distOutDev, countOutDev)
               : queryShardedT<double>(c, shard, (const double*)queriesDev, Q, k, radius, idxOutDev, (double*)distOutDev, countOutDev);
}

int mptg_knn_query_sharded(mptg_comm* c, mptg_knn* shard, const void* queries, uint32_t Q, uint32_t k, double radius, uint32_t* idxOut,
                           void* distOut, uint32_t* countOut) {
    if (!c || !shard || (!queries && Q) || !idxOut || !distOut) return fail(c ? c->ctx : nullptr, MPTG_ERR_BAD_ARG, "mptg_knn_query_sharded: bad argument");
    if (Q == 0) return MPTG_OK;
    mptg_ctx* ctx = c->ctx;
    MPTG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int scalar = knnShardScalar(shard);
    uint32_t first, count;
    sliceOf(Q, c->world, c->rank, first, count);
    const size_t qBytes = (size_t)Q * knnShardScalars(shard) * scalar;
    const size_t iBytes = (size_t)count * k * 4, dBytes = (size_t)count * k * scalar, cBytes = (size_t)count * 4;
    void* dq;
    void* dout;
    int rc = scratch(ctx, 0, qBytes, &dq);
    if (rc) return rc;
    rc = scratch(ctx, 1, iBytes + dBytes + cBytes + 64, &dout);
    if (rc) return rc;
    void* dDist = dout;
    uint32_t* dIdx = (uint32_t*)((char*)dout + ((dBytes + 7) & ~(size_t)7));
    uint32_t* dCnt = dIdx + (size_t)count * k;
    MPTG_CUDA(ctx, cudaMemcpyAsync(dq, queries, qBytes, cudaMemcpyHostToDevice, ctx->stream));
    rc = mptg_knn_query_sharded_dev(c, shard, dq, Q, k, radius, dIdx, dDist, dCnt);
    if (rc) return rc;
    if (count) {
        MPTG_CUDA(ctx, cudaMemcpyAsync(distOut, dDist, dBytes, cudaMemcpyDeviceToHost, ctx->stream));
        MPTG_CUDA(ctx, cudaMemcpyAsync(idxOut, dIdx, iBytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (countOut) MPTG_CUDA(ctx, cudaMemcpyAsync(countOut, dCnt, cBytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    MPTG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPTG_OK;
}

}  // extern "C"
