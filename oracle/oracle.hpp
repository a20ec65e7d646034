// oracle.hpp -- CPU restatement of the MPT planning hot path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, link or execute anything under oracle/.  The product (libmptg.so, include/mptg/*.hpp,
// mpt_b200/) never includes or calls it.
//
// Parity status:
//   * metric distance: PINNED by the reference's own known-answer tests
//     (test/{lp,scaled,so2,so3,se2,se3}_space_test.cpp), re-run in oracle/kat_main.cpp.
//   * interpolate, DiscreteMotionValidator, grid / shapes / link-arm checks, GoalState: arithmetic follows
//     the in-tree reference sources line by line (cited per function) and is PINNED against the
//     reference's own code: those headers are compiled from /root/reference against stand-in
//     Eigen/Nigh headers (oracle/ref_driver.cpp -> oracle/_ref/libref.so) and their outputs on seeded
//     inputs are committed as tests/golden/reference_golden.npz (tests/test_reference_parity.py).
//     Decisions are identical; interpolated quaternions agree to libm tolerance (the reference calls
//     libm, we call mptg_fpmath.h).
//   * planner loops (PRRT, PRRT*, PPRM -- src/mpt/impl/{prrt,prrt_star,pprm}): PINNED against the reference's own
//     planner classes, compiled from /root/reference against the same stand-ins plus an exhaustive-scan Nigh
//     (oracle/ref_planner.cpp, tests/cpp/reference_planner_parity.cpp): same random stream in, identical trees /
//     roadmaps / solution paths / PRRT* costs out at one sample per wave (golden trees in reference_golden.npz).
//   * kNN result order and mesh-mesh collision: PARITY UNPINNED -- the arithmetic lives in Nigh and
//     FCL, which are not in /root/reference and not on this machine.  The oracle defines them:
//     kNN = exact total order by (distance, insertion index); mesh = AABB-overlap && 17-axis
//     separating-axis triangle test (published PQP/FCL TriContact scheme), touching == colliding.
//
// Scalar arithmetic has a written order; compile with -ffp-contract=off.  acos/sin/cos come from
// include/mptg/mptg_fpmath.h so the CUDA kernels can reproduce them bit for bit.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <utility>
#include <vector>

#include "../include/mptg/mptg.h"
#include "../include/mptg/mptg_fpmath.h"

namespace oracle {

namespace fp = mptg::fp;

// ------------------------------------------------------------------ space helpers
inline int partScalars(const mptg_space_part& p) { return p.kind == MPTG_PART_SO3 ? 4 : p.dim; }
inline int spaceScalars(const mptg_space_desc& s) {
    int n = 0;
    for (int i = 0; i < s.n_parts; ++i) n += partScalars(s.part[i]);
    return n;
}
// reference: space.dimensions() (LP: dim; SO2: dim; SO3: 3; Cartesian: sum) -- SURVEY appendix A
inline int spaceDimensions(const mptg_space_desc& s) {
    int n = 0;
    for (int i = 0; i < s.n_parts; ++i) n += s.part[i].kind == MPTG_PART_SO3 ? 3 : s.part[i].dim;
    return n;
}

// ------------------------------------------------------------------ distance (a4)
// L_p over `dim` differences that are already non-negative magnitudes or signed differences.
// p=2: sqrt(fma chain of squares, left to right); p=1: left-to-right sum of |d|; p=0: max |d|.
// Pinned by test/lp_space_test.cpp:49 (== sqrt(20)), so2_space_test.cpp:64.
template <typename S>
S lpNorm(const S* d, int dim, int p) {
    if (p == 2) {
        S acc = d[0] * d[0];
        for (int i = 1; i < dim; ++i) acc = fp::fma_(d[i], d[i], acc);
        return fp::sqrt_(acc);
    } else if (p == 1) {
        S acc = fp::abs_(d[0]);
        for (int i = 1; i < dim; ++i) acc = acc + fp::abs_(d[i]);
        return acc;
    } else {
        S acc = fp::abs_(d[0]);
        for (int i = 1; i < dim; ++i) acc = std::max(acc, fp::abs_(d[i]));
        return acc;
    }
}

template <typename S>
S partDistance(const mptg_space_part& part, const S* a, const S* b) {
    switch (part.kind) {
        case MPTG_PART_LP: {
            S d[MPTG_MAX_SCALARS];
            for (int i = 0; i < part.dim; ++i) d[i] = a[i] - b[i];
            return lpNorm(d, part.dim, part.p);
        }
        case MPTG_PART_SO2: {
            // per coordinate shortest arc: delta=|a-b|, if delta > pi use 2pi - delta
            // pinned by test/so2_space_test.cpp:46-50 (d(-1,3) == 2*M_PI - 4 exactly)
            const S pi = fp::consts<S>::pi();
            S d[MPTG_MAX_SCALARS];
            for (int i = 0; i < part.dim; ++i) {
                S delta = fp::abs_(a[i] - b[i]);
                if (delta > pi) delta = S(2) * pi - delta;
                d[i] = delta;
            }
            return lpNorm(d, part.dim, part.p);
        }
        case MPTG_PART_SO3: {
            // acos(|a.b|): HALF the rotation angle, pinned by test/so3_space_test.cpp:53-55
            S dot = a[0] * b[0];
            dot = fp::fma_(a[1], b[1], dot);
            dot = fp::fma_(a[2], b[2], dot);
            dot = fp::fma_(a[3], b[3], dot);
            S ad = fp::abs_(dot);
            if (ad > S(1)) ad = S(1);
            return fp::acos01(ad);
        }
    }
    return std::numeric_limits<S>::quiet_NaN();
}

// Cartesian: sum over parts of weight*d in tuple order (test/se3_space_test.cpp:70-71,
// se2_space_test.cpp:66); Scaled: d*weight (test/scaled_space_test.cpp:51).
template <typename S>
S distance(const mptg_space_desc& sp, const S* a, const S* b) {
    S total = 0;
    int off = 0;
    for (int i = 0; i < sp.n_parts; ++i) {
        const mptg_space_part& part = sp.part[i];
        S d = partDistance(part, a + off, b + off);
        if (part.weight != 1.0) d = d * S(part.weight);
        total = (i == 0) ? d : total + d;
        off += partScalars(part);
    }
    return total;
}

// ------------------------------------------------------------------ interpolate (a5)
template <typename S>
S so2Bound(S x) {  // wrap by repeated +-2pi (a fmod-style bound fails so2_space_test.cpp:80)
    const S pi = fp::consts<S>::pi();
    while (x > pi) x = x - S(2) * pi;
    while (x < -pi) x = x + S(2) * pi;
    return x;
}

template <typename S>
void partInterpolate(const mptg_space_part& part, const S* a, const S* b, S t, S* q) {
    switch (part.kind) {
        case MPTG_PART_LP:
            // src/mpt/lp_space.hpp:51-52: (b - a) * d + a
            for (int i = 0; i < part.dim; ++i) q[i] = (b[i] - a[i]) * t + a[i];
            break;
        case MPTG_PART_SO2: {
            // src/mpt/so2_space.hpp:53-62
            const S pi = fp::consts<S>::pi();
            for (int i = 0; i < part.dim; ++i) {
                S ccw = b[i] - a[i];
                if (ccw < S(0)) ccw = ccw + S(2) * pi;
                if (ccw < pi) {
                    q[i] = so2Bound(a[i] + ccw * t);
                } else {
                    S cw = S(2) * pi - ccw;
                    q[i] = so2Bound(a[i] - cw * t);
                }
            }
            break;
        }
        case MPTG_PART_SO3: {
            // src/mpt/so3_space.hpp:54-80 (including the signed-dot quirk at :61)
            S d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
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
            for (int i = 0; i < 4; ++i) q[i] = s0 * a[i] + s1 * b[i];
            break;
        }
    }
}

// src/mpt/cartesian_space.hpp:44-67 (same t for every component), scaled_space.hpp:44-51 (pass-through)
template <typename S>
void interpolate(const mptg_space_desc& sp, const S* a, const S* b, S t, S* q) {
    int off = 0;
    for (int i = 0; i < sp.n_parts; ++i) {
        partInterpolate(sp.part[i], a + off, b + off, t, q + off);
        off += partScalars(sp.part[i]);
    }
}

// ------------------------------------------------------------------ kNN (a1, a2)
// Exact brute force with the total order (distance, index).  Rows ascending; radius < 0 = unbounded,
// otherwise keeps d <= radius (SURVEY appendix A "kNN result").
template <typename S>
void knnBrute(const mptg_space_desc& sp, const S* pts, uint32_t n, const S* queries, uint32_t Q, uint32_t k,
              double radius, uint32_t* idxOut, S* distOut, uint32_t* countOut) {
    const int D = spaceScalars(sp);
    const bool bounded = radius >= 0 && std::isfinite(radius);
    const S r = S(radius);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t qi = 0; qi < (int64_t)Q; ++qi) {
        std::vector<std::pair<S, uint32_t>> best;  // ascending, size <= k
        best.reserve(k + 1);
        const S* q = queries + (size_t)qi * D;
        for (uint32_t i = 0; i < n; ++i) {
            S d = distance(sp, pts + (size_t)i * D, q);
            if (d != d) continue;  // NaN (e.g. a NaN query) is never a neighbour
            if (bounded && !(d <= r)) continue;
            if (best.size() == k && !(d < best.back().first)) continue;  // index ascending: ties keep older
            auto it = std::upper_bound(best.begin(), best.end(), std::make_pair(d, i));
            best.insert(it, std::make_pair(d, i));
            if (best.size() > k) best.pop_back();
        }
        for (uint32_t j = 0; j < k; ++j) {
            bool have = j < best.size();
            idxOut[(size_t)qi * k + j] = have ? best[j].second : MPTG_NO_INDEX;
            distOut[(size_t)qi * k + j] = have ? best[j].first : std::numeric_limits<S>::infinity();
        }
        if (countOut) countOut[qi] = (uint32_t)best.size();
    }
}

// ------------------------------------------------------------------ DiscreteMotionValidator (a6)
// src/mpt/discrete_motion_validator.hpp:71-130.  `valid` is the state validator; `from` is assumed
// valid and not checked (:72-73).  Returns the decision; *statesChecked counts validator calls.
template <typename S, typename Valid>
bool discreteMotionValid(const mptg_space_desc& sp, S step, const S* from, const S* to, Valid&& valid,
                         uint64_t* statesChecked = nullptr) {
    constexpr std::size_t Q = 256;  // fixedBisectQueueSize_ (:54)
    uint64_t dummy = 0;
    uint64_t& cnt = statesChecked ? *statesChecked : dummy;
    S tmp[MPTG_MAX_SCALARS];
    ++cnt;
    if (!valid(to)) return false;
    const S invStep = fp::div_(S(1), step);                               // :64
    std::size_t steps = (std::size_t)std::ceil(distance(sp, from, to) * invStep);  // :78
    if (steps < 2) return true;
    const S delta = fp::div_(S(1), S(steps));                             // :82
    std::array<std::pair<std::size_t, std::size_t>, Q> queue;
    queue[0] = {1, steps - 1};
    std::size_t qStart = 0, qEnd = 1;
    auto check = [&](std::size_t i) {
        interpolate(sp, from, to, S(i) * delta, tmp);
        ++cnt;
        return valid(tmp);
    };
    while (qStart != qEnd) {
        auto [mn, mx] = queue[qStart++ % Q];
        if (mn == mx) {
            if (!check(mn)) return false;
        } else if (qEnd + 2 < qStart + Q) {
            std::size_t mid = (mn + mx) / 2;
            if (!check(mid)) return false;
            if (mn < mid) queue[qEnd++ % Q] = {mn, mid - 1};
            if (mid < mx) queue[qEnd++ % Q] = {mid + 1, mx};
        } else {
            for (std::size_t i = mn; i <= mx; ++i)
                if (!check(i)) return false;
        }
    }
    return true;
}

// ------------------------------------------------------------------ occupancy grid (a8)
// demo/png_2d_scenario.hpp:104-117,152-165
template <typename S>
struct Grid {
    int width = 0, height = 0;
    std::vector<uint8_t> occ;  // 1 = obstacle

    bool valid(const S* q) const {
        int x = (int)(q[0] + S(0.5));  // :106-107 (truncation toward zero)
        int y = (int)(q[1] + S(0.5));
        long long idx = (long long)width * y + x;  // :109
        if (idx < 0 || idx >= (long long)width * height) return false;  // reference reads out of bounds (UB): obstacle
        return !occ[(size_t)idx];
    }
    bool validSegment(const S* a, const S* b) const {  // :152-165
        S mid[2] = {(a[0] + b[0]) / S(2), (a[1] + b[1]) / S(2)};
        S dx = b[0] - a[0], dy = b[1] - a[1];
        S distSquared = dx * dx + dy * dy;
        if (distSquared < S(1)) return true;
        if (!valid(mid)) return false;
        if (!validSegment(a, mid)) return false;
        return validSegment(mid, b);
    }
    bool link(const S* a, const S* b) const {  // :112-117
        if (!valid(a) || !valid(b)) return false;
        return validSegment(a, b);
    }
};

// ------------------------------------------------------------------ balls & rects (a10)
// demo/shape_hierarchy.hpp:168-273, demo/holonomic_2d_point_scenario.hpp:95-113,
// test/planner_integration_test.hpp:143-149 (N-D sphere, closed form)
template <typename S>
struct Shapes {
    int dim = 2;
    std::vector<S> centres, radii;  // balls
    std::vector<S> rects;           // x0,y0,x1,y1

    // shape_hierarchy.hpp:259-270, generalised from 2 to `dim` coordinates (left-to-right sums)
    S distPointSegmentSquared(const S* pt, const S* s0, const S* s1) const {
        S c1 = 0, c2 = 0, ww = 0;
        for (int i = 0; i < dim; ++i) {
            S v = s1[i] - s0[i], w = pt[i] - s0[i];
            c1 = (i == 0) ? v * w : c1 + v * w;
            c2 = (i == 0) ? v * v : c2 + v * v;
            ww = (i == 0) ? w * w : ww + w * w;
        }
        if (c1 <= S(0)) return ww;
        if (c2 <= c1) {
            S acc = 0;
            for (int i = 0; i < dim; ++i) {
                S e = pt[i] - s1[i];
                acc = (i == 0) ? e * e : acc + e * e;
            }
            return acc;
        }
        S f = c1 / c2;
        S acc = 0;
        for (int i = 0; i < dim; ++i) {
            S v = s1[i] - s0[i];
            S e = s0[i] - pt[i] + v * f;
            acc = (i == 0) ? e * e : acc + e * e;
        }
        return acc;
    }
    bool ballPointValid(int j, const S* p) const {  // :222-226
        S acc = 0;
        for (int i = 0; i < dim; ++i) {
            S e = p[i] - centres[(size_t)j * dim + i];
            acc = (i == 0) ? e * e : acc + e * e;
        }
        return acc > radii[j] * radii[j];
    }
    bool ballSegmentValid(int j, const S* a, const S* b) const {  // :228-231
        return distPointSegmentSquared(&centres[(size_t)j * dim], a, b) > radii[j] * radii[j];
    }
    bool rectPointValid(int j, const S* p) const {  // :177-182
        const S* r = &rects[(size_t)j * 4];
        return !(p[0] >= r[0] && p[0] <= r[2] && p[1] >= r[1] && p[1] <= r[3]);
    }
    bool rectBisect(int j, const S* a, const S* b) const {  // :191-203
        S mid[2] = {(a[0] + b[0]) / S(2), (a[1] + b[1]) / S(2)};
        S dx = b[0] - a[0], dy = b[1] - a[1];
        S distSquared = dx * dx + dy * dy;
        if (distSquared < S(1)) return true;
        if (!rectPointValid(j, mid)) return false;
        if (!rectBisect(j, a, mid)) return false;
        return rectBisect(j, mid, b);
    }
    bool rectSegmentValid(int j, const S* a, const S* b) const {  // :184-189
        if (!rectPointValid(j, a) || !rectPointValid(j, b)) return false;
        return rectBisect(j, a, b);
    }
    bool valid(const S* q) const {  // holonomic_2d_point_scenario.hpp:95-103
        for (size_t j = 0; j < radii.size(); ++j)
            if (!ballPointValid((int)j, q)) return false;
        for (size_t j = 0; j < rects.size() / 4; ++j)
            if (!rectPointValid((int)j, q)) return false;
        return true;
    }
    bool link(const S* a, const S* b) const {  // :105-113
        for (size_t j = 0; j < radii.size(); ++j)
            if (!ballSegmentValid((int)j, a, b)) return false;
        for (size_t j = 0; j < rects.size() / 4; ++j)
            if (!rectSegmentValid((int)j, a, b)) return false;
        return true;
    }
};

// ------------------------------------------------------------------ N-link planar arm (a9)
// demo/link_manipulator_scenario.hpp:99-138; Circle::segmentIsValid(a,b,r) shape_hierarchy.hpp:234-237
template <typename S>
struct LinkArm {
    int nLinks = 0;
    std::vector<S> lengths;
    S linkRadius = 0;
    std::vector<S> circles;  // cx, cy, r

    static S distPointSegmentSquared2(const S* pt, const S* s0, const S* s1) {  // shape_hierarchy.hpp:259-270
        S vx = s1[0] - s0[0], vy = s1[1] - s0[1];
        S wx = pt[0] - s0[0], wy = pt[1] - s0[1];
        S c1 = vx * wx + vy * wy;
        if (c1 <= S(0)) return wx * wx + wy * wy;
        S c2 = vx * vx + vy * vy;
        if (c2 <= c1) {
            S ex = pt[0] - s1[0], ey = pt[1] - s1[1];
            return ex * ex + ey * ey;
        }
        S f = c1 / c2;
        S ex = s0[0] - pt[0] + vx * f, ey = s0[1] - pt[1] + vy * f;
        return ex * ex + ey * ey;
    }
    bool valid(const S* q) const {  // :99-116
        S from[2] = {0, 0}, to[2];
        S angle = 0;
        for (int i = 0; i < nLinks; ++i) {
            angle = angle + q[i];
            S sn, cs;
            fp::sincos_(angle, &sn, &cs);
            to[0] = from[0] + lengths[i] * cs;
            to[1] = from[1] + lengths[i] * sn;
            for (size_t c = 0; c < circles.size() / 3; ++c) {
                S rr = circles[c * 3 + 2] + linkRadius;
                if (!(distPointSegmentSquared2(&circles[c * 3], from, to) > rr * rr)) return false;
            }
            from[0] = to[0];
            from[1] = to[1];
        }
        return true;
    }
    bool bisectLink(const S* a, const S* b) const {  // :125-138
        S maxDiff = 0;
        for (int i = 0; i < nLinks; ++i) maxDiff = std::max(maxDiff, fp::abs_(a[i] - b[i]));
        if (maxDiff < S(0.02)) return true;
        S mid[MPTG_MAX_SCALARS];
        for (int i = 0; i < nLinks; ++i) mid[i] = (a[i] + b[i]) / S(2);
        if (!valid(mid)) return false;
        if (!bisectLink(a, mid)) return false;
        return bisectLink(mid, b);
    }
    bool link(const S* a, const S* b) const {  // :118-123
        if (!valid(a) || !valid(b)) return false;
        return bisectLink(a, b);
    }
};

// ------------------------------------------------------------------ mesh vs mesh (a7)
// Stands in for fcl::collide(robotBVH, T(q), envBVH, I) with the default request
// (demo/se3_rigid_body_scenario.hpp:282-296): boolean "any triangle pair intersects".
// PARITY UNPINNED (FCL absent).  Definition used here and by the kernels:
//   collide(state) = exists (i,j): aabb(T(q) tri_i) overlaps aabb(env_j)  (closed intervals)
//                                  && no separating axis among the 17 PQP/FCL TriContact axes
//   (strict separation only: touching counts as contact).
// The answer does not depend on BVH shape because the leaf test includes the per-triangle AABB test
// and all node bounds are conservative.
struct V3 {
    double x, y, z;
};
template <typename S>
struct Tri {
    S v[3][3];
};

template <typename S>
inline void quatToRot(const S* q /*x y z w*/, S R[9]) {
    // Eigen::Quaternion::toRotationMatrix operation order (Eigen/src/Geometry/Quaternion.h):
    // tx=2x, ty=2y, tz=2z, twx=tx*w ... ; R(0,0)=1-(tyy+tzz) ...
    const S x = q[0], y = q[1], z = q[2], w = q[3];
    const S tx = S(2) * x, ty = S(2) * y, tz = S(2) * z;
    const S twx = tx * w, twy = ty * w, twz = tz * w;
    const S txx = tx * x, txy = ty * x, txz = tz * x;
    const S tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = S(1) - (tyy + tzz);
    R[1] = txy - twz;
    R[2] = txz + twy;
    R[3] = txy + twz;
    R[4] = S(1) - (txx + tzz);
    R[5] = tyz - twx;
    R[6] = txz - twy;
    R[7] = tyz + twx;
    R[8] = S(1) - (txx + tyy);
}

// T(q) v = R v + t, rows left to right: (R0*x + R1*y) + R2*z, then + t
template <typename S>
inline void xformPoint(const S R[9], const S t[3], const S v[3], S out[3]) {
    out[0] = ((R[0] * v[0] + R[1] * v[1]) + R[2] * v[2]) + t[0];
    out[1] = ((R[3] * v[0] + R[4] * v[1]) + R[5] * v[2]) + t[1];
    out[2] = ((R[6] * v[0] + R[7] * v[1]) + R[8] * v[2]) + t[2];
}

template <typename S>
inline void cross3(const S a[3], const S b[3], S o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
template <typename S>
inline S dot3(const S a[3], const S b[3]) {
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}

// project6: true when the projections of the two triangles on `ax` overlap or touch.
// gapOut (optional) = signed separation along the axis divided by |ax| (>0 separated).
template <typename S>
inline bool project6(const S ax[3], const S p[3][3], const S q[3][3], double* gapOut) {
    S P0 = dot3(ax, p[0]), P1 = dot3(ax, p[1]), P2 = dot3(ax, p[2]);
    S Q0 = dot3(ax, q[0]), Q1 = dot3(ax, q[1]), Q2 = dot3(ax, q[2]);
    S mx1 = std::max(P0, std::max(P1, P2)), mn1 = std::min(P0, std::min(P1, P2));
    S mx2 = std::max(Q0, std::max(Q1, Q2)), mn2 = std::min(Q0, std::min(Q1, Q2));
    if (gapOut) {
        double len = std::sqrt((double)ax[0] * ax[0] + (double)ax[1] * ax[1] + (double)ax[2] * ax[2]);
        double g = std::max((double)mn1 - (double)mx2, (double)mn2 - (double)mx1);
        *gapOut = len > 0 ? g / len : -std::numeric_limits<double>::infinity();
    }
    if (mn1 > mx2) return false;
    if (mn2 > mx1) return false;
    return true;
}

// 17-axis separating axis test (PQP TriContact / FCL Intersect::intersect_Triangle scheme):
// both normals, 9 edge x edge, 3 normal x edge per triangle; coordinates relative to P[0].
// Returns true when the triangles intersect or touch.  marginOut = max over axes of normalised gap.
template <typename S>
bool triTriIntersect(const S P[3][3], const S Qt[3][3], double* marginOut = nullptr) {
    S p[3][3], q[3][3];
    for (int i = 0; i < 3; ++i)
        for (int c = 0; c < 3; ++c) {
            p[i][c] = P[i][c] - P[0][c];
            q[i][c] = Qt[i][c] - P[0][c];
        }
    S e[3][3], f[3][3];
    for (int c = 0; c < 3; ++c) {
        e[0][c] = p[1][c] - p[0][c];
        e[1][c] = p[2][c] - p[1][c];
        e[2][c] = p[0][c] - p[2][c];
        f[0][c] = q[1][c] - q[0][c];
        f[1][c] = q[2][c] - q[1][c];
        f[2][c] = q[0][c] - q[2][c];
    }
    S n1[3], m1[3];
    cross3(e[0], e[1], n1);
    cross3(f[0], f[1], m1);
    bool hit = true;
    double margin = -std::numeric_limits<double>::infinity();
    auto test = [&](const S ax[3]) {
        double g;
        bool ov = project6(ax, p, q, marginOut ? &g : nullptr);
        if (marginOut) margin = std::max(margin, g);
        if (!ov) hit = false;
        return ov;
    };
    S ax[3];
    if (!test(n1) && !marginOut) return false;
    if (!test(m1) && !marginOut) return false;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            cross3(e[i], f[j], ax);
            if (!test(ax) && !marginOut) return false;
        }
    for (int i = 0; i < 3; ++i) {
        cross3(e[i], n1, ax);
        if (!test(ax) && !marginOut) return false;
    }
    for (int i = 0; i < 3; ++i) {
        cross3(f[i], m1, ax);
        if (!test(ax) && !marginOut) return false;
    }
    if (marginOut) *marginOut = margin;
    return hit;
}

// ---- second, independent triangle-triangle formulation (test infrastructure for the contact tests).
// FCL is absent, so the 17-axis SAT above cannot be pinned to it; what can be checked is that a formulation built from
// different arithmetic -- orientation predicates (the determinants Guigue & Devillers' test is built from), evaluated in
// double -- gives the same answer wherever the configuration is not within the contact band.  Closed sets: touching
// counts as contact, as above.  Two triangles meet iff an edge of one meets the other triangle.
inline double orient3d(const double* a, const double* b, const double* c, const double* d) {
    const double adx = a[0] - d[0], ady = a[1] - d[1], adz = a[2] - d[2];
    const double bdx = b[0] - d[0], bdy = b[1] - d[1], bdz = b[2] - d[2];
    const double cdx = c[0] - d[0], cdy = c[1] - d[1], cdz = c[2] - d[2];
    return adx * (bdy * cdz - bdz * cdy) - ady * (bdx * cdz - bdz * cdx) + adz * (bdx * cdy - bdy * cdx);
}
inline double orient2d(const double* a, const double* b, const double* c) {
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0]);
}
inline bool onSegment2d(const double* a, const double* b, const double* c) {  // c collinear with a,b: inside [a,b]?
    return std::min(a[0], b[0]) <= c[0] && c[0] <= std::max(a[0], b[0]) && std::min(a[1], b[1]) <= c[1] && c[1] <= std::max(a[1], b[1]);
}
inline bool segSeg2d(const double* a, const double* b, const double* c, const double* d) {
    const double o1 = orient2d(a, b, c), o2 = orient2d(a, b, d), o3 = orient2d(c, d, a), o4 = orient2d(c, d, b);
    if (((o1 > 0 && o2 < 0) || (o1 < 0 && o2 > 0)) && ((o3 > 0 && o4 < 0) || (o3 < 0 && o4 > 0))) return true;
    if (o1 == 0 && onSegment2d(a, b, c)) return true;
    if (o2 == 0 && onSegment2d(a, b, d)) return true;
    if (o3 == 0 && onSegment2d(c, d, a)) return true;
    if (o4 == 0 && onSegment2d(c, d, b)) return true;
    return false;
}
inline bool pointInTri2d(const double* x, const double* p, const double* q, const double* r) {
    const double o1 = orient2d(p, q, x), o2 = orient2d(q, r, x), o3 = orient2d(r, p, x);
    return (o1 >= 0 && o2 >= 0 && o3 >= 0) || (o1 <= 0 && o2 <= 0 && o3 <= 0);
}
// closed segment [a,b] against closed triangle (p,q,r)
inline bool segTriPredicates(const double* a, const double* b, const double* p, const double* q, const double* r) {
    const double sa = orient3d(p, q, r, a), sb = orient3d(p, q, r, b);
    if ((sa > 0 && sb > 0) || (sa < 0 && sb < 0)) return false;
    if (sa == 0 && sb == 0) {  // the segment lies in the triangle's plane: decide in 2-D, dropping the normal's largest component
        const double e0[3] = {q[0] - p[0], q[1] - p[1], q[2] - p[2]}, e1[3] = {r[0] - p[0], r[1] - p[1], r[2] - p[2]};
        const double n[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0]};
        int drop = 0;
        if (std::fabs(n[1]) > std::fabs(n[drop])) drop = 1;
        if (std::fabs(n[2]) > std::fabs(n[drop])) drop = 2;
        const int u = (drop + 1) % 3, v = (drop + 2) % 3;
        const double A[2] = {a[u], a[v]}, B[2] = {b[u], b[v]}, P[2] = {p[u], p[v]}, Q[2] = {q[u], q[v]}, R[2] = {r[u], r[v]};
        return pointInTri2d(A, P, Q, R) || pointInTri2d(B, P, Q, R) || segSeg2d(A, B, P, Q) || segSeg2d(A, B, Q, R) || segSeg2d(A, B, R, P);
    }
    // the segment meets the plane in one point (possibly an end point): that point is inside the triangle iff the
    // line (a,b) has the three edges on one side
    const double o1 = orient3d(a, b, p, q), o2 = orient3d(a, b, q, r), o3 = orient3d(a, b, r, p);
    return (o1 >= 0 && o2 >= 0 && o3 >= 0) || (o1 <= 0 && o2 <= 0 && o3 <= 0);
}
inline bool triTriPredicates(const double P[3][3], const double Q[3][3]) {
    for (int i = 0; i < 3; ++i) {
        if (segTriPredicates(P[i], P[(i + 1) % 3], Q[0], Q[1], Q[2])) return true;
        if (segTriPredicates(Q[i], Q[(i + 1) % 3], P[0], P[1], P[2])) return true;
    }
    return false;
}

template <typename S>
struct Aabb {
    S lo[3], hi[3];
    void reset() {
        for (int c = 0; c < 3; ++c) {
            lo[c] = std::numeric_limits<S>::infinity();
            hi[c] = -std::numeric_limits<S>::infinity();
        }
    }
    void add(const S p[3]) {
        for (int c = 0; c < 3; ++c) {
            lo[c] = std::min(lo[c], p[c]);
            hi[c] = std::max(hi[c], p[c]);
        }
    }
    void add(const Aabb& o) {
        add(o.lo);
        add(o.hi);
    }
    bool overlaps(const Aabb& o, S pad = 0) const {
        for (int c = 0; c < 3; ++c)
            if (lo[c] - pad > o.hi[c] || o.lo[c] - pad > hi[c]) return false;
        return true;
    }
};

template <typename S>
struct Bvh {  // binary median-split AABB tree, one triangle per leaf
    struct Node {
        Aabb<S> box;
        int left = -1, right = -1, tri = -1;
    };
    std::vector<Node> nodes;
    std::vector<Tri<S>> tris;

    static Aabb<S> triBox(const Tri<S>& t) {
        Aabb<S> b;
        b.reset();
        for (int i = 0; i < 3; ++i) b.add(t.v[i]);
        return b;
    }
    int build(std::vector<int>& ids, int lo, int hi) {
        Node n;
        n.box.reset();
        for (int i = lo; i < hi; ++i) n.box.add(triBox(tris[ids[i]]));
        int me = (int)nodes.size();
        nodes.push_back(n);
        if (hi - lo == 1) {
            nodes[me].tri = ids[lo];
            return me;
        }
        int axis = 0;
        S ext = -1;
        for (int c = 0; c < 3; ++c)
            if (n.box.hi[c] - n.box.lo[c] > ext) ext = n.box.hi[c] - n.box.lo[c], axis = c;
        int mid = (lo + hi) / 2;
        std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int a, int b) {
            S ca = tris[a].v[0][axis] + tris[a].v[1][axis] + tris[a].v[2][axis];
            S cb = tris[b].v[0][axis] + tris[b].v[1][axis] + tris[b].v[2][axis];
            return ca < cb || (ca == cb && a < b);
        });
        int l = build(ids, lo, mid);
        int r = build(ids, mid, hi);
        nodes[me].left = l;
        nodes[me].right = r;
        return me;
    }
    void set(const float* tris9, uint32_t n) {
        tris.resize(n);
        for (uint32_t i = 0; i < n; ++i)
            for (int v = 0; v < 3; ++v)
                for (int c = 0; c < 3; ++c) tris[i].v[v][c] = S(tris9[(size_t)i * 9 + v * 3 + c]);
        nodes.clear();
        if (n == 0) return;
        std::vector<int> ids(n);
        for (uint32_t i = 0; i < n; ++i) ids[i] = (int)i;
        nodes.reserve(2 * n);
        build(ids, 0, (int)n);
    }
};

template <typename S>
struct MeshPair {
    Bvh<S> robot, env;  // robot in its local frame, env in world frame
    double scale = 1;   // env AABB diagonal, used for the relative near-contact band
    bool predicates = false;  // decide triangle pairs with triTriPredicates (on the same transformed vertices, in double)
    struct Counters {
        uint64_t bvTests = 0, triTests = 0, states = 0;
    };

    void set(const float* robotTris, uint32_t nr, const float* envTris, uint32_t ne) {
        robot.set(robotTris, nr);
        env.set(envTris, ne);
        if (!env.nodes.empty()) {
            const auto& b = env.nodes[0].box;
            double dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
            scale = std::sqrt(dx * dx + dy * dy + dz * dz);
        }
    }

    // world AABB of a robot-local box under (R,t): centre/half-extent form with |R|, padded
    static Aabb<S> worldBox(const Aabb<S>& b, const S R[9], const S t[3]) {
        S c[3], h[3];
        for (int i = 0; i < 3; ++i) {
            c[i] = (b.lo[i] + b.hi[i]) * S(0.5);
            h[i] = (b.hi[i] - b.lo[i]) * S(0.5);
        }
        Aabb<S> w;
        S mag = 0;
        S cw[3], hw[3];
        for (int r = 0; r < 3; ++r) {
            cw[r] = R[r * 3] * c[0] + R[r * 3 + 1] * c[1] + R[r * 3 + 2] * c[2] + t[r];
            hw[r] = fp::abs_(R[r * 3]) * h[0] + fp::abs_(R[r * 3 + 1]) * h[1] + fp::abs_(R[r * 3 + 2]) * h[2];
            mag += fp::abs_(cw[r]) + hw[r];
        }
        S pad = S(64) * fp::consts<S>::eps() * mag;
        for (int r = 0; r < 3; ++r) {
            w.lo[r] = cw[r] - hw[r] - pad;
            w.hi[r] = cw[r] + hw[r] + pad;
        }
        return w;
    }

    // Returns true when the state is collision free.  If marginOut != nullptr, does not stop at the
    // first hit and reports min over candidate pairs (AABBs within `band`) of the SAT margin.
    bool valid(const S* state /*qx qy qz qw tx ty tz*/, double* marginOut = nullptr, double band = 0,
               Counters* cnt = nullptr) const {
        Counters local;
        Counters& C = cnt ? *cnt : local;
        ++C.states;
        if (robot.nodes.empty() || env.nodes.empty()) {
            if (marginOut) *marginOut = std::numeric_limits<double>::infinity();
            return true;
        }
        S R[9];
        quatToRot(state, R);
        const S* t = state + 4;
        bool collide = false;
        double minMargin = std::numeric_limits<double>::infinity();
        std::vector<std::pair<int, int>> stack;
        stack.push_back({0, 0});
        const S pad = S(band);
        while (!stack.empty()) {
            auto [rn, en] = stack.back();
            stack.pop_back();
            const auto& a = robot.nodes[rn];
            const auto& b = env.nodes[en];
            ++C.bvTests;
            Aabb<S> wb = worldBox(a.box, R, t);
            if (!wb.overlaps(b.box, pad)) continue;
            if (a.tri >= 0 && b.tri >= 0) {
                S P[3][3];
                for (int v = 0; v < 3; ++v) xformPoint(R, t, robot.tris[a.tri].v[v], P[v]);
                Aabb<S> pb;
                pb.reset();
                for (int v = 0; v < 3; ++v) pb.add(P[v]);
                if (!pb.overlaps(b.box, pad)) continue;
                ++C.triTests;
                if (marginOut) {
                    double m;
                    bool hit = triTriIntersect<S>(P, env.tris[b.tri].v, &m);
                    if (hit && pb.overlaps(b.box, 0)) collide = true;
                    minMargin = std::min(minMargin, m);
                } else if (predicates) {
                    double Pd[3][3], Qd[3][3];
                    for (int v = 0; v < 3; ++v)
                        for (int c = 0; c < 3; ++c) Pd[v][c] = (double)P[v][c], Qd[v][c] = (double)env.tris[b.tri].v[v][c];
                    if (triTriPredicates(Pd, Qd)) return false;
                } else if (triTriIntersect<S>(P, env.tris[b.tri].v)) {
                    return false;
                }
                continue;
            }
            // descend the node with the larger box (robot extents are rotation invariant enough)
            bool descendRobot;
            if (a.tri >= 0) descendRobot = false;
            else if (b.tri >= 0) descendRobot = true;
            else {
                S ea = 0, eb = 0;
                for (int c = 0; c < 3; ++c) {
                    ea = std::max(ea, a.box.hi[c] - a.box.lo[c]);
                    eb = std::max(eb, b.box.hi[c] - b.box.lo[c]);
                }
                descendRobot = ea > eb;
            }
            if (descendRobot) {
                stack.push_back({a.left, en});
                stack.push_back({a.right, en});
            } else {
                stack.push_back({rn, b.left});
                stack.push_back({rn, b.right});
            }
        }
        if (marginOut) *marginOut = minMargin;
        return !collide;
    }
};

}  // namespace oracle
