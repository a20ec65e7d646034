// mock_mptg.cpp -- TEST-ONLY stand-in for libmptg.so: the same C ABI (include/mptg/mptg.h) answered by the
// CPU oracle, so the host-side wave planners (include/mptg/planner.hpp) can be exercised on a
// machine without a GPU.  Built into tests/cpp/_build/libmptg_mock.so by tests/test_host_cpp.py;
// never shipped, never loaded by the product (the product library has no CPU fallback).
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../oracle/oracle.hpp"
#include "../../oracle/oracle_nao.hpp"

using namespace oracle;

struct mptg_ctx {
    std::string err;
    uint64_t launches = 0;
};
struct mptg_knn {
    mptg_ctx* ctx;
    mptg_space_desc sp;
    int D;
    uint32_t capacity;
    std::vector<unsigned char> pts;
    uint32_t size = 0;
};
struct mptg_geom {
    mptg_ctx* ctx;
    int kind, scalar, D;
    Grid<float> gridF;
    Grid<double> gridD;
    Shapes<float> shapesF;
    Shapes<double> shapesD;
    LinkArm<float> armF;
    LinkArm<double> armD;
    MeshPair<float> meshF;
    NaoCup<float> naoF;
    NaoCup<double> naoD;
};

static thread_local std::string g_err;
static int fail(mptg_ctx* c, int code, const char* msg) {
    if (c) c->err = msg;
    g_err = msg;
    return code;
}

template <typename S>
static bool validOne(mptg_geom* g, const S* q);
template <>
bool validOne<float>(mptg_geom* g, const float* q) {
    switch (g->kind) {
        case MPTG_GEOM_GRID: return g->gridF.valid(q);
        case MPTG_GEOM_SHAPES: return g->shapesF.valid(q);
        case MPTG_GEOM_LINKARM: return g->armF.valid(q);
        case MPTG_GEOM_NAOCUP: return g->naoF.valid(q);
        default: return g->meshF.valid(q);
    }
}
template <>
bool validOne<double>(mptg_geom* g, const double* q) {
    switch (g->kind) {
        case MPTG_GEOM_GRID: return g->gridD.valid(q);
        case MPTG_GEOM_SHAPES: return g->shapesD.valid(q);
        case MPTG_GEOM_NAOCUP: return g->naoD.valid(q);
        default: return g->armD.valid(q);
    }
}

extern "C" {
int mptg_abi_version(void) { return MPTG_ABI_VERSION; }
int mptg_ctx_create(int, mptg_ctx** out) {
    *out = new mptg_ctx();
    return MPTG_OK;
}
int mptg_ctx_destroy(mptg_ctx* c) {
    delete c;
    return MPTG_OK;
}
int mptg_sync(mptg_ctx*) { return MPTG_OK; }
const char* mptg_last_error(const mptg_ctx* c) { return c ? c->err.c_str() : g_err.c_str(); }
void* mptg_ctx_stream(mptg_ctx*) { return nullptr; }
uint64_t mptg_ctx_launch_count(const mptg_ctx* c) { return c->launches; }
int mptg_ctx_sm_count(const mptg_ctx*) { return 1; }
int mptg_space_scalars(const mptg_space_desc* s) { return spaceScalars(*s); }
int mptg_space_dimensions(const mptg_space_desc* s) { return spaceDimensions(*s); }

int mptg_knn_create(mptg_ctx* ctx, const mptg_space_desc* sp, uint32_t capacity, mptg_knn** out) {
    auto* k = new mptg_knn{ctx, *sp, spaceScalars(*sp), capacity, {}, 0};
    *out = k;
    return MPTG_OK;
}
int mptg_knn_destroy(mptg_knn* k) {
    delete k;
    return MPTG_OK;
}
int mptg_knn_set_strategy(mptg_knn*, int) { return MPTG_OK; }
int mptg_knn_set_index_map(mptg_knn*, uint32_t, uint32_t) { return MPTG_OK; }
uint32_t mptg_knn_size(const mptg_knn* k) { return k->size; }
int mptg_knn_insert(mptg_knn* k, const void* states, uint32_t count, uint32_t* first) {
    if ((uint64_t)k->size + count > k->capacity) return fail(k->ctx, MPTG_ERR_CAPACITY, "mock: capacity exceeded");
    if (first) *first = k->size;
    const size_t bytes = (size_t)count * k->D * k->sp.scalar;
    const unsigned char* p = (const unsigned char*)states;
    k->pts.insert(k->pts.end(), p, p + bytes);
    k->size += count;
    ++k->ctx->launches;
    return MPTG_OK;
}
int mptg_knn_query(mptg_knn* k, const void* q, uint32_t Q, uint32_t kk, double radius, uint32_t* idx, void* dist, uint32_t* cnt) {
    if (kk == 0 || kk > MPTG_MAX_K) return fail(k->ctx, MPTG_ERR_BAD_ARG, "mock: k out of range");
    if (k->sp.scalar == MPTG_F32)
        knnBrute<float>(k->sp, (const float*)k->pts.data(), k->size, (const float*)q, Q, kk, radius, idx, (float*)dist, cnt);
    else
        knnBrute<double>(k->sp, (const double*)k->pts.data(), k->size, (const double*)q, Q, kk, radius, idx, (double*)dist, cnt);
    ++k->ctx->launches;
    return MPTG_OK;
}

int mptg_grid_create(mptg_ctx* ctx, int scalar, int32_t w, int32_t h, const uint8_t* occ, mptg_geom** out) {
    auto* g = new mptg_geom();
    g->ctx = ctx, g->kind = MPTG_GEOM_GRID, g->scalar = scalar, g->D = 2;
    g->gridF.width = g->gridD.width = w;
    g->gridF.height = g->gridD.height = h;
    g->gridF.occ.assign(occ, occ + (size_t)w * h);
    g->gridD.occ = g->gridF.occ;
    *out = g;
    return MPTG_OK;
}
int mptg_shapes_create(mptg_ctx* ctx, int scalar, int32_t dim, int32_t nb, const double* c, const double* r, int32_t nr,
                       const double* rects, mptg_geom** out) {
    auto* g = new mptg_geom();
    g->ctx = ctx, g->kind = MPTG_GEOM_SHAPES, g->scalar = scalar, g->D = dim;
    g->shapesF.dim = g->shapesD.dim = dim;
    for (int i = 0; i < nb * dim; ++i) g->shapesF.centres.push_back((float)c[i]), g->shapesD.centres.push_back(c[i]);
    for (int i = 0; i < nb; ++i) g->shapesF.radii.push_back((float)r[i]), g->shapesD.radii.push_back(r[i]);
    for (int i = 0; i < nr * 4; ++i) g->shapesF.rects.push_back((float)rects[i]), g->shapesD.rects.push_back(rects[i]);
    *out = g;
    return MPTG_OK;
}
int mptg_linkarm_create(mptg_ctx* ctx, int scalar, int32_t n, const double* len, double radius, int32_t nc, const double* c,
                        mptg_geom** out) {
    auto* g = new mptg_geom();
    g->ctx = ctx, g->kind = MPTG_GEOM_LINKARM, g->scalar = scalar, g->D = n;
    g->armF.nLinks = g->armD.nLinks = n;
    g->armF.linkRadius = (float)radius, g->armD.linkRadius = radius;
    for (int i = 0; i < n; ++i) g->armF.lengths.push_back((float)len[i]), g->armD.lengths.push_back(len[i]);
    for (int i = 0; i < nc * 3; ++i) g->armF.circles.push_back((float)c[i]), g->armD.circles.push_back(c[i]);
    *out = g;
    return MPTG_OK;
}
int mptg_mesh_pair_create(mptg_ctx* ctx, int scalar, uint32_t nr, const float* rt, uint32_t ne, const float* et, mptg_geom** out) {
    if (scalar != MPTG_F32) return fail(ctx, MPTG_ERR_UNSUPPORTED, "mock: f32 meshes only");
    auto* g = new mptg_geom();
    g->ctx = ctx, g->kind = MPTG_GEOM_MESH, g->scalar = scalar, g->D = 7;
    g->meshF.set(rt, nr, et, ne);
    *out = g;
    return MPTG_OK;
}
int mptg_naocup_create(mptg_ctx* ctx, int scalar, mptg_geom** out) {
    auto* g = new mptg_geom();
    g->ctx = ctx, g->kind = MPTG_GEOM_NAOCUP, g->scalar = scalar, g->D = 10;
    *out = g;
    return MPTG_OK;
}
int mptg_naocup_configs(int scalar, double* start, double* goal, double* lo, double* hi) {  // naocup.hpp:196-301
    static const double startC[10] = {1.125998, -0.691876, 1.888312, 0.776246, 0.245398, 1.259372, 0.279146, -1.587732, -0.510780, -1.823800};
    static const double goalC[10] = {0.258284303377494,  -0.2699099199363406, -0.01113121187052224, 1.2053012757652763,  1.2716626717484503,
                                     -0.9826967097045605, 0.07355836822937814, 0.25450053440459897,  -0.9512909033938429, -0.5297424293532234};
    static const double loDeg[10] = {-119.5, -94.5, -119.5, 0.5, -104.5, -119.5, 0.5, -119.5, -89.5, -104.5};
    static const double hiDeg[10] = {119.5, -0.5, 119.5, 89.5, 104.5, 119.5, 94.5, 119.5, -0.5, 104.5};
    for (int i = 0; i < 10; ++i) {
        const double kd = 3.14159265358979323846 / 180.0;
        const float kf = float(float(3.14159265358979323846) / float(180.0));
        if (start) start[i] = scalar == MPTG_F32 ? (double)(float)startC[i] : startC[i];
        if (goal) goal[i] = scalar == MPTG_F32 ? (double)(float)goalC[i] : goalC[i];
        if (lo) lo[i] = scalar == MPTG_F32 ? (double)(float(loDeg[i]) * kf) : loDeg[i] * kd;
        if (hi) hi[i] = scalar == MPTG_F32 ? (double)(float(hiDeg[i]) * kf) : hiDeg[i] * kd;
    }
    return MPTG_OK;
}
int mptg_geom_destroy(mptg_geom* g) {
    delete g;
    return MPTG_OK;
}
int mptg_geom_kind(const mptg_geom* g) { return g->kind; }

int mptg_valid_batch(mptg_geom* g, const void* st, uint32_t n, uint8_t* ok, uint8_t* near) {
    if (near) memset(near, 0, n);
    for (uint32_t i = 0; i < n; ++i)
        ok[i] = g->scalar == MPTG_F32 ? validOne<float>(g, (const float*)st + (size_t)i * g->D) : validOne<double>(g, (const double*)st + (size_t)i * g->D);
    ++g->ctx->launches;
    return MPTG_OK;
}
int mptg_link_batch(mptg_geom* g, const mptg_space_desc* sp, const void* from, const void* to, uint32_t n, double step, uint8_t* ok, uint8_t* near) {
    if (near) memset(near, 0, n);
    for (uint32_t i = 0; i < n; ++i) {
        if (g->scalar == MPTG_F32) {
            const float* a = (const float*)from + (size_t)i * g->D;
            const float* b = (const float*)to + (size_t)i * g->D;
            switch (g->kind) {
                case MPTG_GEOM_GRID: ok[i] = g->gridF.link(a, b); break;
                case MPTG_GEOM_SHAPES: ok[i] = g->shapesF.link(a, b); break;
                case MPTG_GEOM_LINKARM: ok[i] = g->armF.link(a, b); break;
                case MPTG_GEOM_NAOCUP: ok[i] = g->naoF.link(a, b); break;
                default: ok[i] = discreteMotionValid<float>(*sp, (float)step, a, b, [&](const float* q) { return g->meshF.valid(q); }); break;
            }
        } else {
            const double* a = (const double*)from + (size_t)i * g->D;
            const double* b = (const double*)to + (size_t)i * g->D;
            switch (g->kind) {
                case MPTG_GEOM_GRID: ok[i] = g->gridD.link(a, b); break;
                case MPTG_GEOM_SHAPES: ok[i] = g->shapesD.link(a, b); break;
                case MPTG_GEOM_NAOCUP: ok[i] = g->naoD.link(a, b); break;
                default: ok[i] = g->armD.link(a, b); break;
            }
        }
    }
    ++g->ctx->launches;
    return MPTG_OK;
}
int mptg_steer_batch(mptg_ctx* ctx, const mptg_space_desc* sp, const void* near, const void* sample, const void* d, uint32_t n,
                     double range, void* out, void* distOut) {
    const int D = spaceScalars(*sp);
    auto run = [&](auto tag) {
        using S = decltype(tag);
        for (uint32_t i = 0; i < n; ++i) {
            const S* nr = (const S*)near + (size_t)i * D;
            const S* sm = (const S*)sample + (size_t)i * D;
            S* o = (S*)out + (size_t)i * D;
            if (((const S*)d)[i] > S(range)) interpolate<S>(*sp, nr, sm, fp::div_(S(range), ((const S*)d)[i]), o);
            else std::memcpy(o, sm, sizeof(S) * D);
            if (distOut) ((S*)distOut)[i] = distance<S>(*sp, nr, o);
        }
    };
    if (sp->scalar == MPTG_F32) run(float{});
    else run(double{});
    ++ctx->launches;
    return MPTG_OK;
}
}
