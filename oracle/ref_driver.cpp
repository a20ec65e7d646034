// ref_driver.cpp -- the REFERENCE'S OWN hot-path code, compiled from where it lies under
// /root/reference, behind a small C interface.  TEST INFRASTRUCTURE ONLY (built into oracle/_ref/).
//
// Compiled from the reference tree (no copies): src/mpt/{lp,so2,so3,scaled,cartesian,se3}_space.hpp
// (interpolate), src/mpt/discrete_motion_validator.hpp, src/mpt/goal_state.hpp,
// demo/{shape_hierarchy,holonomic_2d_point_scenario,png_2d_scenario,link_manipulator_scenario}.hpp.
// Their external dependencies are absent here, so they are satisfied by stand-ins of OURS under
// oracle/shim/: Eigen/Dense (value types + element-wise operators), nigh/*.hpp (metric spaces -- the
// distance arithmetic in there is ours, pinned by the reference's KATs), png.h (declarations).
// mpt/log.hpp (needs real Eigen I/O) and mpt/box_bounds.hpp (Eigen block expressions) are replaced by
// trivial stand-ins through their include guards; neither is on the hot path.
// So: interpolate(), DiscreteMotionValidator::operator(), PNG2dScenario::valid/link,
// Holonomic2DPointScenario::valid/link, LinkManipulatorScenario::valid/link and the shape tests below
// are executed exactly as the reference wrote them (libm sin/cos/acos, -ffp-contract=off).
// headers the reference's log.hpp would have pulled in
#include <cassert>
#include <cmath>
#include <cstdint>
#include <iostream>
#include <memory>
#include <vector>

// ---- stand-ins selected through the reference's include guards
#define MPT_LOG_HPP_
#define MPT_LOG(level) \
    if (true) {        \
    } else             \
        std::clog
#define MPT_BOX_BOUNDS_HPP
namespace unc::robotics::mpt {
template <typename S, int dim>
class BoxBounds {
public:
    BoxBounds() {}
    template <typename A, typename B>
    BoxBounds(const A&, const B&) {}
};
}  // namespace unc::robotics::mpt

#include <mpt/discrete_motion_validator.hpp>
#include <mpt/goal_state.hpp>
#include <mpt/lp_space.hpp>
#include <mpt/se3_space.hpp>
#include <mpt/so2_space.hpp>
#include <mpt/so3_space.hpp>

#include <holonomic_2d_point_scenario.hpp>
#include <link_manipulator_scenario.hpp>
#include <png_2d_scenario.hpp>

using namespace unc::robotics;

namespace {
template <int N>
Eigen::Matrix<double, N, 1> vec(const double* p) {
    Eigen::Matrix<double, N, 1> v;
    for (int i = 0; i < N; ++i) v[i] = p[i];
    return v;
}
template <typename S>
mpt::SE3State<S> se3(const S* p) {  // ABI order: qx qy qz qw tx ty tz
    mpt::SE3State<S> q;
    q.rotation() = Eigen::Quaternion<S>(p[3], p[0], p[1], p[2]);
    q.translation() = Eigen::Matrix<S, 3, 1>(p[4], p[5], p[6]);
    return q;
}
template <typename S>
void unse3(const mpt::SE3State<S>& q, S* p) {
    for (int i = 0; i < 4; ++i) p[i] = q.rotation().coeffs()[i];
    for (int i = 0; i < 3; ++i) p[4 + i] = q.translation()[i];
}

template <int N>
int armLink(const double* lengths, double radius, int nCircles, const double* cxcyr, const double* a, const double* b, uint32_t n,
            uint8_t* validA, uint8_t* link) {
    using Scenario = mpt_demo::LinkManipulatorScenario<double, N>;
    std::vector<shape::Circle<double>> circles;
    for (int i = 0; i < nCircles; ++i) circles.emplace_back(cxcyr[3 * i], cxcyr[3 * i + 1], cxcyr[3 * i + 2]);
    std::vector<double> len(lengths, lengths + N);
    typename Scenario::State goal;
    goal.fill(0);
    Scenario sc(goal, circles, len, radius);
    for (uint32_t i = 0; i < n; ++i) {
        auto qa = vec<N>(a + (size_t)i * N), qb = vec<N>(b + (size_t)i * N);
        if (validA) validA[i] = sc.valid(qa);
        if (link) link[i] = sc.link(qa, qb);
    }
    return 0;
}

// validator handed to the reference's DiscreteMotionValidator: a callback into the caller
struct CallbackValidator {
    int (*fn)(const float* state7, void* user);
    void* user;
    uint64_t* count;
    bool operator()(const mpt::SE3State<float>& q) const {
        float p[7];
        unse3(q, p);
        ++*count;
        return fn(p, user) != 0;
    }
};
}  // namespace

extern "C" {

// ---- interpolate(space, a, b, t): src/mpt/lp_space.hpp:44-54 etc.
int ref_interpolate_l2_3(const double* a, const double* b, const double* t, uint32_t n, double* out) {
    mpt::LPSpace<double, 3, 2> space;
    for (uint32_t i = 0; i < n; ++i) {
        auto q = mpt::interpolate(space, vec<3>(a + 3 * (size_t)i), vec<3>(b + 3 * (size_t)i), t[i]);
        for (int c = 0; c < 3; ++c) out[3 * (size_t)i + c] = q[c];
    }
    return 0;
}
int ref_interpolate_so2(const double* a, const double* b, const double* t, uint32_t n, double* out) {
    mpt::SO2Space<double> space;
    for (uint32_t i = 0; i < n; ++i) out[i] = mpt::interpolate(space, a[i], b[i], t[i]);
    return 0;
}
int ref_interpolate_se3_f64(const double* a, const double* b, const double* t, uint32_t n, double* out) {
    mpt::SE3Space<double, 50> space;
    for (uint32_t i = 0; i < n; ++i) unse3(mpt::interpolate(space, se3(a + 7 * (size_t)i), se3(b + 7 * (size_t)i), t[i]), out + 7 * (size_t)i);
    return 0;
}
int ref_interpolate_se3_f32(const float* a, const float* b, const float* t, uint32_t n, float* out) {
    mpt::SE3Space<float, 50> space;
    for (uint32_t i = 0; i < n; ++i) unse3(mpt::interpolate(space, se3(a + 7 * (size_t)i), se3(b + 7 * (size_t)i), t[i]), out + 7 * (size_t)i);
    return 0;
}

// ---- DiscreteMotionValidator<SE3Space<float,50>, Validator>::operator() (discrete_motion_validator.hpp:71-130)
// with the caller's state validator; returns decisions and the number of validator calls per edge.
int ref_dmv_se3_f32(const float* from, const float* to, uint32_t n, float step, int (*valid)(const float*, void*), void* user,
                    uint8_t* ok, uint64_t* statesPerEdge) {
    using Space = mpt::SE3Space<float, 50>;
    Space space;
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t cnt = 0;
        mpt::DiscreteMotionValidator<Space, CallbackValidator> dmv(space, step, CallbackValidator{valid, user, &cnt});
        ok[i] = dmv(se3(from + 7 * (size_t)i), se3(to + 7 * (size_t)i));
        if (statesPerEdge) statesPerEdge[i] = cnt;
    }
    return 0;
}

// ---- PNG2dScenario::valid / link (demo/png_2d_scenario.hpp:104-117,152-165).  In-range indices only:
// the reference indexes std::vector<bool> out of bounds for x+0.5 >= width on the last row / y+0.5 >= height.
int ref_grid(int width, int height, const uint8_t* occ, const double* a, const double* b, uint32_t n, uint8_t* validA, uint8_t* link) {
    std::vector<bool> obst((size_t)width * height);
    for (size_t i = 0; i < obst.size(); ++i) obst[i] = occ[i] != 0;
    mpt_demo::PNG2dScenario<double> sc(width, height, Eigen::Vector2d(0, 0), obst);
    for (uint32_t i = 0; i < n; ++i) {
        auto qa = vec<2>(a + 2 * (size_t)i), qb = vec<2>(b + 2 * (size_t)i);
        if (validA) validA[i] = sc.valid(qa);
        if (link) link[i] = sc.link(qa, qb);
    }
    return 0;
}

// ---- Holonomic2DPointScenario::valid / link (demo/holonomic_2d_point_scenario.hpp:95-113, shape_hierarchy.hpp:168-273)
int ref_holonomic(int nCircles, const double* cxcyr, int nRects, const double* rects, const double* a, const double* b, uint32_t n,
                  uint8_t* validA, uint8_t* link) {
    std::vector<shape::Circle<double>> circles;
    std::vector<shape::Rect<double>> rs;
    for (int i = 0; i < nCircles; ++i) circles.emplace_back(cxcyr[3 * i], cxcyr[3 * i + 1], cxcyr[3 * i + 2]);
    for (int i = 0; i < nRects; ++i) rs.emplace_back(rects[4 * i], rects[4 * i + 1], rects[4 * i + 2], rects[4 * i + 3]);
    mpt_demo::Holonomic2DPointScenario<double> sc(1024, 512, circles, rs, Eigen::Vector2d(0, 0));
    for (uint32_t i = 0; i < n; ++i) {
        auto qa = vec<2>(a + 2 * (size_t)i), qb = vec<2>(b + 2 * (size_t)i);
        if (validA) validA[i] = sc.valid(qa);
        if (link) link[i] = sc.link(qa, qb);
    }
    return 0;
}

// ---- LinkManipulatorScenario<double,N>::valid / link (demo/link_manipulator_scenario.hpp:99-138)
int ref_linkarm(int nLinks, const double* lengths, double radius, int nCircles, const double* cxcyr, const double* a, const double* b,
                uint32_t n, uint8_t* validA, uint8_t* link) {
    switch (nLinks) {
        case 5: return armLink<5>(lengths, radius, nCircles, cxcyr, a, b, n, validA, link);
        case 8: return armLink<8>(lengths, radius, nCircles, cxcyr, a, b, n, validA, link);
        case 16: return armLink<16>(lengths, radius, nCircles, cxcyr, a, b, n, validA, link);
        case 32: return armLink<32>(lengths, radius, nCircles, cxcyr, a, b, n, validA, link);
    }
    return -1;
}

// ---- GoalState (src/mpt/goal_state.hpp:64-69)
int ref_goal_l2_3(const double* goal, double radius, const double* q, uint32_t n, uint8_t* isGoal, double* dist) {
    using Space = mpt::LPSpace<double, 3, 2>;
    Space space;
    mpt::GoalState<Space> g(radius, vec<3>(goal));
    for (uint32_t i = 0; i < n; ++i) {
        auto r = g(space, vec<3>(q + 3 * (size_t)i));
        isGoal[i] = r.first;
        dist[i] = r.second;
    }
    return 0;
}
}
