// oracle_capi.cpp -- C entry points of the CPU oracle (liboracle.so), for ctypes.
// TEST INFRASTRUCTURE ONLY (see oracle.hpp header).  OpenMP over independent items; this is also
// what bench.py times as the CPU baseline ("port" of the reference path, all host threads).
#include <omp.h>

#include <cstring>
#include <memory>

#include "oracle.hpp"
#include "oracle_nao.hpp"
#include "oracle_sample.hpp"
#include "oracle_tree.hpp"

using namespace oracle;

namespace {
struct Geom {
    int kind = 0, scalar = MPTG_F32;
    Grid<float> gridF;
    Grid<double> gridD;
    Shapes<float> shapesF;
    Shapes<double> shapesD;
    LinkArm<float> armF;
    LinkArm<double> armD;
    MeshPair<float> meshF;
    MeshPair<double> meshD;
    NaoCup<float> naoF;
    NaoCup<double> naoD;
    uint64_t stats[4] = {0, 0, 0, 0};
};

template <typename S>
struct Pick;
template <>
struct Pick<float> {
    static Grid<float>& grid(Geom& g) { return g.gridF; }
    static Shapes<float>& shapes(Geom& g) { return g.shapesF; }
    static LinkArm<float>& arm(Geom& g) { return g.armF; }
    static MeshPair<float>& mesh(Geom& g) { return g.meshF; }
    static NaoCup<float>& nao(Geom& g) { return g.naoF; }
};
template <>
struct Pick<double> {
    static Grid<double>& grid(Geom& g) { return g.gridD; }
    static Shapes<double>& shapes(Geom& g) { return g.shapesD; }
    static LinkArm<double>& arm(Geom& g) { return g.armD; }
    static MeshPair<double>& mesh(Geom& g) { return g.meshD; }
    static NaoCup<double>& nao(Geom& g) { return g.naoD; }
};

template <typename S>
int validBatch(Geom& g, const S* st, uint32_t n, uint8_t* ok, double* margin) {
    int D = g.kind == MPTG_GEOM_MESH ? 7 : g.kind == MPTG_GEOM_NAOCUP ? 10 : g.kind == MPTG_GEOM_LINKARM ? Pick<S>::arm(g).nLinks
                                       : g.kind == MPTG_GEOM_SHAPES  ? Pick<S>::shapes(g).dim
                                                                     : 2;
    uint64_t bv = 0, tt = 0;
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : bv, tt)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        const S* q = st + (size_t)i * D;
        typename MeshPair<S>::Counters C;
        bool v = false;
        double m = std::numeric_limits<double>::quiet_NaN();
        switch (g.kind) {
            case MPTG_GEOM_GRID: v = Pick<S>::grid(g).valid(q); break;
            case MPTG_GEOM_SHAPES: v = Pick<S>::shapes(g).valid(q); break;
            case MPTG_GEOM_LINKARM: v = Pick<S>::arm(g).valid(q); break;
            case MPTG_GEOM_NAOCUP: {
                uint64_t pairs = 0;
                v = margin ? Pick<S>::nao(g).valid(q, nullptr, &m, &pairs) : Pick<S>::nao(g).valid(q, nullptr, nullptr, &pairs);
                tt += pairs;
                break;
            }
            case MPTG_GEOM_MESH:
                if (margin) v = Pick<S>::mesh(g).valid(q, &m, 1e-6 * Pick<S>::mesh(g).scale, &C);
                else v = Pick<S>::mesh(g).valid(q, nullptr, 0, &C);
                break;
        }
        ok[i] = v ? 1 : 0;
        if (margin) margin[i] = m;
        bv += C.bvTests;
        tt += C.triTests;
    }
    g.stats[0] = n, g.stats[1] = bv, g.stats[2] = tt, g.stats[3] = n;
    return 0;
}

template <typename S>
int linkBatch(Geom& g, const mptg_space_desc* sp, const S* from, const S* to, uint32_t n, double step,
              uint8_t* ok, uint8_t* nearContact, double tolRel, uint64_t* statesOut) {
    int D = g.kind == MPTG_GEOM_MESH ? 7 : g.kind == MPTG_GEOM_NAOCUP ? 10 : g.kind == MPTG_GEOM_LINKARM ? Pick<S>::arm(g).nLinks
                                       : g.kind == MPTG_GEOM_SHAPES  ? Pick<S>::shapes(g).dim
                                                                     : 2;
    uint64_t totalStates = 0, bv = 0, tt = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : totalStates, bv, tt)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        typename MeshPair<S>::Counters C;
        const S* a = from + (size_t)i * D;
        const S* b = to + (size_t)i * D;
        bool v = false;
        bool nc = false;
        switch (g.kind) {
            case MPTG_GEOM_GRID: v = Pick<S>::grid(g).link(a, b); break;
            case MPTG_GEOM_SHAPES: v = Pick<S>::shapes(g).link(a, b); break;
            case MPTG_GEOM_LINKARM: v = Pick<S>::arm(g).link(a, b); break;
            case MPTG_GEOM_NAOCUP: {
                uint64_t cnt = 0;
                v = Pick<S>::nao(g).link(a, b, &cnt);
                totalStates += cnt;
                if (nearContact) {  // some midpoint of the edge's whole recursion decides within tolRel of a threshold
                    double m = std::numeric_limits<double>::infinity();
                    Pick<S>::nao(g).linkMargin(a, b, &m);
                    nc = m < tolRel;
                }
                break;
            }
            case MPTG_GEOM_MESH: {
                auto& mesh = Pick<S>::mesh(g);
                uint64_t cnt = 0;
                v = discreteMotionValid<S>(
                    *sp, S(step), a, b, [&](const S* q) { return mesh.valid(q, nullptr, 0, &C); }, &cnt);
                totalStates += cnt;
                bv += C.bvTests;
                tt += C.triTests;
                if (nearContact) {
                    // exhaustive pass (no early out) to find states within the contact band
                    double tol = tolRel * mesh.scale;
                    uint64_t c2 = 0;
                    discreteMotionValid<S>(
                        *sp, S(step), a, b,
                        [&](const S* q) {
                            double m;
                            mesh.valid(q, &m, tol);
                            if (std::fabs(m) < tol) nc = true;
                            return true;
                        },
                        &c2);
                }
                break;
            }
        }
        ok[i] = v ? 1 : 0;
        if (nearContact) nearContact[i] = nc ? 1 : 0;
    }
    if (statesOut) *statesOut = totalStates;
    g.stats[0] = totalStates, g.stats[1] = bv, g.stats[2] = tt, g.stats[3] = n;
    return 0;
}
}  // namespace

extern "C" {

int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }

int orc_distance_batch(const mptg_space_desc* sp, const void* a, const void* b, uint32_t n, void* out) {
    int D = spaceScalars(*sp);
    if (sp->scalar == MPTG_F32) {
        for (uint32_t i = 0; i < n; ++i)
            ((float*)out)[i] = distance<float>(*sp, (const float*)a + (size_t)i * D, (const float*)b + (size_t)i * D);
    } else {
        for (uint32_t i = 0; i < n; ++i)
            ((double*)out)[i] = distance<double>(*sp, (const double*)a + (size_t)i * D, (const double*)b + (size_t)i * D);
    }
    return 0;
}

int orc_interpolate_batch(const mptg_space_desc* sp, const void* a, const void* b, const void* t, uint32_t n,
                          void* out) {
    int D = spaceScalars(*sp);
    if (sp->scalar == MPTG_F32) {
        for (uint32_t i = 0; i < n; ++i)
            interpolate<float>(*sp, (const float*)a + (size_t)i * D, (const float*)b + (size_t)i * D,
                               ((const float*)t)[i], (float*)out + (size_t)i * D);
    } else {
        for (uint32_t i = 0; i < n; ++i)
            interpolate<double>(*sp, (const double*)a + (size_t)i * D, (const double*)b + (size_t)i * D,
                                ((const double*)t)[i], (double*)out + (size_t)i * D);
    }
    return 0;
}

// steer (impl/prrt/prrt.hpp:430-434, impl/prrt_star/prrt_star.hpp:529-536)
int orc_steer_batch(const mptg_space_desc* sp, const void* near, const void* sample, const void* d, uint32_t n,
                    double range, void* out, void* distOut) {
    int D = spaceScalars(*sp);
    auto run = [&](auto tag) {
        using S = decltype(tag);
        const S* nr = (const S*)near;
        const S* sm = (const S*)sample;
        const S* dd = (const S*)d;
        S* o = (S*)out;
        for (uint32_t i = 0; i < n; ++i) {
            if (dd[i] > S(range)) {
                interpolate<S>(*sp, nr + (size_t)i * D, sm + (size_t)i * D, fp::div_(S(range), dd[i]), o + (size_t)i * D);
            } else {
                std::memcpy(o + (size_t)i * D, sm + (size_t)i * D, sizeof(S) * D);
            }
            if (distOut) ((S*)distOut)[i] = distance<S>(*sp, nr + (size_t)i * D, o + (size_t)i * D);
        }
    };
    if (sp->scalar == MPTG_F32) run(float{});
    else run(double{});
    return 0;
}

int orc_knn(const mptg_space_desc* sp, const void* pts, uint32_t n, const void* queries, uint32_t Q, uint32_t k,
            double radius, uint32_t* idx, void* dist, uint32_t* count) {
    if (sp->scalar == MPTG_F32)
        knnBrute<float>(*sp, (const float*)pts, n, (const float*)queries, Q, k, radius, idx, (float*)dist, count);
    else
        knnBrute<double>(*sp, (const double*)pts, n, (const double*)queries, Q, k, radius, idx, (double*)dist, count);
    return 0;
}

// Tree-accelerated exact kNN (same results as orc_knn); the CPU baseline for large trees.
void* orc_tree_create(const mptg_space_desc* sp, const void* pts, uint32_t n) {
    if (sp->scalar != MPTG_F32) return nullptr;
    auto* t = new BoxTree<float>();
    t->build(*sp, (const float*)pts, n);
    return t;
}
void orc_tree_destroy(void* t) { delete (BoxTree<float>*)t; }
int orc_tree_knn(void* tree, const void* queries, uint32_t Q, uint32_t k, double radius, uint32_t* idx, void* dist,
                 uint32_t* count, uint64_t* distEvals) {
    auto* t = (BoxTree<float>*)tree;
    uint64_t ev = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : ev)
    for (int64_t qi = 0; qi < (int64_t)Q; ++qi) {
        ev += t->query((const float*)queries + (size_t)qi * t->D, k, radius, idx + (size_t)qi * k,
                       (float*)dist + (size_t)qi * k, count ? count + qi : nullptr);
    }
    if (distEvals) *distEvals = ev;
    return 0;
}

void* orc_grid_create(int scalar, int w, int h, const uint8_t* occ) {
    auto* g = new Geom();
    g->kind = MPTG_GEOM_GRID;
    g->scalar = scalar;
    g->gridF.width = g->gridD.width = w;
    g->gridF.height = g->gridD.height = h;
    g->gridF.occ.assign(occ, occ + (size_t)w * h);
    for (auto& c : g->gridF.occ) c = c ? 1 : 0;
    g->gridD.occ = g->gridF.occ;
    return g;
}

void* orc_shapes_create(int scalar, int dim, int nBalls, const double* centres, const double* radii, int nRects,
                        const double* rects) {
    auto* g = new Geom();
    g->kind = MPTG_GEOM_SHAPES;
    g->scalar = scalar;
    g->shapesF.dim = g->shapesD.dim = dim;
    for (int i = 0; i < nBalls * dim; ++i) {
        g->shapesF.centres.push_back((float)centres[i]);
        g->shapesD.centres.push_back(centres[i]);
    }
    for (int i = 0; i < nBalls; ++i) {
        g->shapesF.radii.push_back((float)radii[i]);
        g->shapesD.radii.push_back(radii[i]);
    }
    for (int i = 0; i < nRects * 4; ++i) {
        g->shapesF.rects.push_back((float)rects[i]);
        g->shapesD.rects.push_back(rects[i]);
    }
    return g;
}

void* orc_linkarm_create(int scalar, int nLinks, const double* lengths, double linkRadius, int nCircles,
                         const double* cxcyr) {
    auto* g = new Geom();
    g->kind = MPTG_GEOM_LINKARM;
    g->scalar = scalar;
    g->armF.nLinks = g->armD.nLinks = nLinks;
    g->armF.linkRadius = (float)linkRadius;
    g->armD.linkRadius = linkRadius;
    for (int i = 0; i < nLinks; ++i) {
        g->armF.lengths.push_back((float)lengths[i]);
        g->armD.lengths.push_back(lengths[i]);
    }
    for (int i = 0; i < nCircles * 3; ++i) {
        g->armF.circles.push_back((float)cxcyr[i]);
        g->armD.circles.push_back(cxcyr[i]);
    }
    return g;
}

void* orc_naocup_create(int scalar) {
    auto* g = new Geom();
    g->kind = MPTG_GEOM_NAOCUP;
    g->scalar = scalar;
    return g;
}

void* orc_mesh_pair_create(int scalar, uint32_t nr, const float* robotTris, uint32_t ne, const float* envTris) {
    auto* g = new Geom();
    g->kind = MPTG_GEOM_MESH;
    g->scalar = scalar;
    if (scalar == MPTG_F32) g->meshF.set(robotTris, nr, envTris, ne);
    else g->meshD.set(robotTris, nr, envTris, ne);
    return g;
}

void orc_geom_destroy(void* g) { delete (Geom*)g; }

// decide triangle pairs of mesh states with the orientation-predicate formulation instead of the SAT (contact tests)
void orc_mesh_use_predicates(void* geom, int on) {
    Geom& g = *(Geom*)geom;
    g.meshF.predicates = g.meshD.predicates = on != 0;
}

// n triangle pairs (9 doubles each, world coordinates): the 17-axis SAT with its margin, and the predicate formulation
int orc_tri_pairs(uint32_t n, const double* P, const double* Q, uint8_t* satOut, double* marginOut, uint8_t* predOut) {
    for (uint32_t i = 0; i < n; ++i) {
        double p[3][3], q[3][3];
        for (int v = 0; v < 3; ++v)
            for (int c = 0; c < 3; ++c) p[v][c] = P[(size_t)i * 9 + v * 3 + c], q[v][c] = Q[(size_t)i * 9 + v * 3 + c];
        double m;
        satOut[i] = triTriIntersect<double>(p, q, &m) ? 1 : 0;
        marginOut[i] = m;
        predOut[i] = triTriPredicates(p, q) ? 1 : 0;
    }
    return 0;
}

int orc_valid_batch(void* geom, const void* states, uint32_t n, uint8_t* ok, double* margin) {
    Geom& g = *(Geom*)geom;
    return g.scalar == MPTG_F32 ? validBatch<float>(g, (const float*)states, n, ok, margin)
                                : validBatch<double>(g, (const double*)states, n, ok, margin);
}

int orc_link_batch(void* geom, const mptg_space_desc* sp, const void* from, const void* to, uint32_t n, double step,
                   uint8_t* ok, uint8_t* nearContact, double tolRel, uint64_t* statesOut) {
    Geom& g = *(Geom*)geom;
    return g.scalar == MPTG_F32
               ? linkBatch<float>(g, sp, (const float*)from, (const float*)to, n, step, ok, nearContact, tolRel, statesOut)
               : linkBatch<double>(g, sp, (const double*)from, (const double*)to, n, step, ok, nearContact, tolRel,
                                   statesOut);
}

// counters of the last valid/link batch: [0]=states checked (mesh), [1]=BV pair tests, [2]=triangle
// pair tests, [3]=items.  These feed the algorithmic-flop figure of SURVEY.md section 8(d).
int orc_geom_counters(void* geom, uint64_t out[4]) {
    Geom& g = *(Geom*)geom;
    for (int i = 0; i < 4; ++i) out[i] = g.stats[i];
    return 0;
}

// Exhaustive / sampled accuracy figures of mptg_fpmath.h against libm (see fpmath_check.cpp)
// one Philox4x32-10 block with the counter layout of the sample streams: (g_lo, g_hi, blk, 0), key = seed
void orc_philox_block(uint64_t seed, uint64_t g, uint32_t blk, uint32_t out[4]) { Philox::block(seed, g, blk, out); }

// samples first .. first+n-1 of stream `seed`; goal (or NULL) with its bias
int orc_sample_batch(const mptg_space_desc* sp, const double* lo, const double* hi, uint64_t seed, uint64_t first, uint32_t n,
                     const void* goal, double goalBias, void* out) {
    int D = 0;
    for (int i = 0; i < sp->n_parts; ++i) D += sp->part[i].kind == MPTG_PART_SO3 ? 4 : sp->part[i].dim;
    for (uint32_t i = 0; i < n; ++i) {
        if (sp->scalar == MPTG_F32) sampleState<float>(*sp, lo, hi, seed, first + i, (const float*)goal, goalBias, (float*)out + (size_t)i * D);
        else sampleState<double>(*sp, lo, hi, seed, first + i, (const double*)goal, goalBias, (double*)out + (size_t)i * D);
    }
    return 0;
}
// the raw uniforms of samples first .. first+n-1: positions 0 .. perState-1 of each sample's stream (position 0 is the
// goal-bias draw); float or double per `scalar`
int orc_sample_uniforms(int scalar, uint64_t seed, uint64_t first, uint32_t n, int perState, void* out) {
    for (uint32_t i = 0; i < n; ++i) {
        if (scalar == MPTG_F32) {
            SampleStream<float> st(seed, first + i);
            for (int j = 0; j < perState; ++j) ((float*)out)[(size_t)i * perState + j] = st.next();
        } else {
            SampleStream<double> st(seed, first + i);
            for (int j = 0; j < perState; ++j) ((double*)out)[(size_t)i * perState + j] = st.next();
        }
    }
    return 0;
}
int orc_sample_from_uniforms(const mptg_space_desc* sp, const double* lo, const double* hi, const void* uniforms, int perState, uint32_t n,
                             void* out) {
    int D = 0;
    for (int i = 0; i < sp->n_parts; ++i) D += sp->part[i].kind == MPTG_PART_SO3 ? 4 : sp->part[i].dim;
    for (uint32_t i = 0; i < n; ++i) {
        int j = 0;
        if (sp->scalar == MPTG_F32) {
            const float* u = (const float*)uniforms + (size_t)i * perState;
            sampleFromUniforms<float>(*sp, lo, hi, [&]() { return u[j++]; }, (float*)out + (size_t)i * D);
        } else {
            const double* u = (const double*)uniforms + (size_t)i * perState;
            sampleFromUniforms<double>(*sp, lo, hi, [&]() { return u[j++]; }, (double*)out + (size_t)i * D);
        }
    }
    return 0;
}

double orc_acos01f(float x) { return fp::acos01(x); }
double orc_acos01d(double x) { return fp::acos01(x); }
void orc_sincosd(double x, double* s, double* c) { fp::sincos_(x, s, c); }
}
