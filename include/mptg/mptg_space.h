// mptg_space.h -- metric spaces for device and host code: distance (a4) and interpolate (a5).
//
// Operation order is part of the contract (bit-exact against the CPU oracle, DESIGN.md "arithmetic"):
//   LP p=2   sqrt(fma chain of squared differences, coordinate 0 first)      [test/lp_space_test.cpp:49]
//   LP p=1   left-to-right sum of |d|;  p=inf  max |d|
//   SO2      per coordinate |a-b|, reflected at pi (2pi - d), then the LP norm [test/so2_space_test.cpp:46-64]
//   SO3      acos01(min(1,|fma chain dot|)): half the rotation angle           [test/so3_space_test.cpp:53-55]
//   product  sum over parts, in order, of d*weight (weight skipped when 1)     [test/se3_space_test.cpp:70-71]
// interpolate follows src/mpt/{lp,so2,so3,scaled,cartesian}_space.hpp with plain (unfused) arithmetic.
// Compile with --fmad=false: only the explicit fma_ calls may fuse.
#pragma once

#include "mptg.h"
#include "mptg_fpmath.h"

namespace mptg {

template <typename S>
struct DevSpace {
    int nParts;
    int D;  // scalars per state
    int kind[MPTG_MAX_PARTS];
    int p[MPTG_MAX_PARTS];
    int dim[MPTG_MAX_PARTS];  // scalars of the part
    int off[MPTG_MAX_PARTS];
    S weight[MPTG_MAX_PARTS];
    int weighted[MPTG_MAX_PARTS];
};

template <typename S>
inline DevSpace<S> makeDevSpace(const mptg_space_desc& s) {
    DevSpace<S> d{};
    d.nParts = s.n_parts;
    int off = 0;
    for (int i = 0; i < s.n_parts; ++i) {
        d.kind[i] = s.part[i].kind;
        d.p[i] = s.part[i].p;
        d.dim[i] = s.part[i].kind == MPTG_PART_SO3 ? 4 : s.part[i].dim;
        d.off[i] = off;
        d.weight[i] = (S)s.part[i].weight;
        d.weighted[i] = s.part[i].weight != 1.0;
        off += d.dim[i];
    }
    d.D = off;
    return d;
}

// compile-time shapes with a dedicated code path; everything else interprets DevSpace at run time
enum SpaceShape { SHAPE_GENERIC = 0, SHAPE_SE3 = 1, SHAPE_L2_2 = 2, SHAPE_L2_3 = 3, SHAPE_L1 = 4 };  // L1: one unweighted L1 part of any dimension (N-link arms)

inline SpaceShape classifySpace(const mptg_space_desc& s) {
    if (s.n_parts == 2 && s.part[0].kind == MPTG_PART_SO3 && s.part[1].kind == MPTG_PART_LP && s.part[1].p == 2 &&
        s.part[1].dim == 3)
        return SHAPE_SE3;
    if (s.n_parts == 1 && s.part[0].kind == MPTG_PART_LP && s.part[0].p == 2 && s.part[0].weight == 1.0) {
        if (s.part[0].dim == 2) return SHAPE_L2_2;
        if (s.part[0].dim == 3) return SHAPE_L2_3;
    }
    if (s.n_parts == 1 && s.part[0].kind == MPTG_PART_LP && s.part[0].p == 1 && s.part[0].weight == 1.0) return SHAPE_L1;
    return SHAPE_GENERIC;
}

namespace dev {

namespace fp = ::mptg::fp;

template <typename S>
MPTG_HD S so3Dist(S a0, S a1, S a2, S a3, S b0, S b1, S b2, S b3) {
    S dot = a0 * b0;
    dot = fp::fma_(a1, b1, dot);
    dot = fp::fma_(a2, b2, dot);
    dot = fp::fma_(a3, b3, dot);
    S ad = fp::abs_(dot);
    if (ad > S(1)) ad = S(1);
    return fp::acos01(ad);
}

template <typename S>
MPTG_HD S l2Dist3(S a0, S a1, S a2, S b0, S b1, S b2) {
    S d0 = a0 - b0, d1 = a1 - b1, d2 = a2 - b2;
    S acc = d0 * d0;
    acc = fp::fma_(d1, d1, acc);
    acc = fp::fma_(d2, d2, acc);
    return fp::sqrt_(acc);
}

// One part; A and B are callables i -> scalar (i relative to the part).
template <typename S, typename A, typename B>
MPTG_HD S partDistance(int kind, int p, int dim, A a, B b) {
    if (kind == MPTG_PART_SO3) return so3Dist<S>(a(0), a(1), a(2), a(3), b(0), b(1), b(2), b(3));
    const S pi = fp::consts<S>::pi();
    S acc = S(0);
    for (int i = 0; i < dim; ++i) {
        S d = a(i) - b(i);
        if (kind == MPTG_PART_SO2) {
            d = fp::abs_(d);
            if (d > pi) d = S(2) * pi - d;
        }
        if (p == 2) acc = (i == 0) ? d * d : fp::fma_(d, d, acc);
        else if (p == 1) acc = (i == 0) ? fp::abs_(d) : acc + fp::abs_(d);
        else acc = (i == 0) ? fp::abs_(d) : (fp::abs_(d) > acc ? fp::abs_(d) : acc);
    }
    return p == 2 ? fp::sqrt_(acc) : acc;
}

template <typename S, typename A, typename B>
MPTG_HD S distance(const DevSpace<S>& sp, A a, B b) {
    S total = S(0);
    for (int i = 0; i < sp.nParts; ++i) {
        const int off = sp.off[i];
        S d = partDistance<S>(
            sp.kind[i], sp.p[i], sp.dim[i], [&](int j) { return a(off + j); }, [&](int j) { return b(off + j); });
        if (sp.weighted[i]) d = d * sp.weight[i];
        total = (i == 0) ? d : total + d;
    }
    return total;
}

// SE(3) fast path: same arithmetic as distance() on {SO3 w0, LP2(3) w1}
template <typename S>
MPTG_HD S se3Distance(S w0, bool weighted0, S w1, bool weighted1, const S* a, const S* b) {
    S dr = so3Dist<S>(a[0], a[1], a[2], a[3], b[0], b[1], b[2], b[3]);
    if (weighted0) dr = dr * w0;
    S dt = l2Dist3<S>(a[4], a[5], a[6], b[4], b[5], b[6]);
    if (weighted1) dt = dt * w1;
    return dr + dt;
}

template <typename S>
MPTG_HD S so2Bound(S x) {
    const S pi = fp::consts<S>::pi();
    while (x > pi) x = x - S(2) * pi;
    while (x < -pi) x = x + S(2) * pi;
    return x;
}

// interpolate(space, a, b, t) -> q  (a, b, q: AoS pointers of one state)
template <typename S>
MPTG_HD void interpolate(const DevSpace<S>& sp, const S* a, const S* b, S t, S* q) {
    for (int i = 0; i < sp.nParts; ++i) {
        const int off = sp.off[i];
        const S* pa = a + off;
        const S* pb = b + off;
        S* pq = q + off;
        if (sp.kind[i] == MPTG_PART_LP) {
            for (int j = 0; j < sp.dim[i]; ++j) pq[j] = (pb[j] - pa[j]) * t + pa[j];  // lp_space.hpp:51-52
        } else if (sp.kind[i] == MPTG_PART_SO2) {                                      // so2_space.hpp:53-62
            const S pi = fp::consts<S>::pi();
            for (int j = 0; j < sp.dim[i]; ++j) {
                S ccw = pb[j] - pa[j];
                if (ccw < S(0)) ccw = ccw + S(2) * pi;
                if (ccw < pi) {
                    pq[j] = so2Bound(pa[j] + ccw * t);
                } else {
                    S cw = S(2) * pi - ccw;
                    pq[j] = so2Bound(pa[j] - cw * t);
                }
            }
        } else {  // so3_space.hpp:54-80
            S d = pa[0] * pb[0] + pa[1] * pb[1] + pa[2] * pb[2] + pa[3] * pb[3];
            S ad = fp::abs_(d);
            S s0, s1;
            if (d >= S(1) - fp::consts<S>::eps()) {
                s0 = S(1) - t;
                s1 = t;
            } else {
                S theta = fp::acos01(ad > S(1) ? S(1) : ad);
                S sinTheta = fp::sin_(theta);
                s0 = fp::div_(fp::sin_((S(1) - t) * theta), sinTheta);
                s1 = fp::div_(fp::sin_(t * theta), sinTheta);
            }
            if (d < S(0)) s1 = -s1;
            for (int j = 0; j < 4; ++j) pq[j] = s0 * pa[j] + s1 * pb[j];
        }
    }
}

}  // namespace dev
}  // namespace mptg
