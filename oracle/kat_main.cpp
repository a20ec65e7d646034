// kat_main.cpp -- the reference's own known-answer tests for the hot-path arithmetic, re-expressed
// without Eigen/Nigh and run against the oracle.  TEST INFRASTRUCTURE ONLY.
// Each check cites the reference test it restates (paths relative to the reference root).
// Exit code 0 = all pass; prints one line per check.  Driven by tests/test_oracle.py.
#include <cstdio>
#include <cstring>

#include "oracle.hpp"

using namespace oracle;

static int failures = 0;
#define CHECK(name, cond)                                                     \
    do {                                                                      \
        bool ok_ = (cond);                                                    \
        std::printf("%s %s\n", ok_ ? "PASS" : "FAIL", name);                  \
        if (!ok_) ++failures;                                                 \
    } while (0)

static mptg_space_desc lp(int dim, int p, double w = 1.0) {
    mptg_space_desc s{};
    s.n_parts = 1;
    s.scalar = MPTG_F64;
    s.part[0] = {MPTG_PART_LP, p, dim, 0, w};
    return s;
}
static mptg_space_desc so2(int dim, int p) {
    mptg_space_desc s{};
    s.n_parts = 1;
    s.scalar = MPTG_F64;
    s.part[0] = {MPTG_PART_SO2, p, dim, 0, 1.0};
    return s;
}
static mptg_space_desc so3() {
    mptg_space_desc s{};
    s.n_parts = 1;
    s.scalar = MPTG_F64;
    s.part[0] = {MPTG_PART_SO3, 0, 4, 0, 1.0};
    return s;
}
static mptg_space_desc se3(double so3w, double l2w) {  // src/mpt/se3_space.hpp:91-108: rotation first
    mptg_space_desc s{};
    s.n_parts = 2;
    s.scalar = MPTG_F64;
    s.part[0] = {MPTG_PART_SO3, 0, 4, 0, so3w};
    s.part[1] = {MPTG_PART_LP, 2, 3, 0, l2w};
    return s;
}
static mptg_space_desc se2(double so2w, double l2w) {  // src/mpt/se2_space.hpp:62-80: translation first
    mptg_space_desc s{};
    s.n_parts = 2;
    s.scalar = MPTG_F64;
    s.part[0] = {MPTG_PART_LP, 2, 2, 0, l2w};
    s.part[1] = {MPTG_PART_SO2, 1, 1, 0, so2w};
    return s;
}

// Eigen::AngleAxisd(angle, axis) -> quaternion coeffs (x,y,z,w)
static void angleAxis(double angle, const double axis[3], double q[4]) {
    double s = std::sin(angle / 2), c = std::cos(angle / 2);
    q[0] = axis[0] * s, q[1] = axis[1] * s, q[2] = axis[2] * s, q[3] = c;
}
// quaternion -> Eigen::AngleAxisd (angle, axis)
static void toAngleAxis(const double q[4], double* angle, double axis[3]) {
    double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    *angle = 2 * std::atan2(n, std::fabs(q[3]));
    if (q[3] < 0) n = -n;
    axis[0] = q[0] / n, axis[1] = q[1] / n, axis[2] = q[2] / n;
}

int main() {
    const double M_PI_ = 3.14159265358979323846;
    {  // test/lp_space_test.cpp:39-50
        auto s = lp(3, 2);
        double a[3] = {1, 2, 3}, b[3] = {1, 0, -1};
        CHECK("lp_space_test:distance == sqrt(20)", distance<double>(s, a, b) == std::sqrt(0.0 + 2.0 * 2.0 + 4.0 * 4.0));
        double c[3];
        interpolate<double>(s, a, b, 0.1, c);  // test/lp_space_test.cpp:53-66
        CHECK("lp_space_test:interpolate == (1,1.8,2.6)", c[0] == 1.0 && c[1] == 1.8 && c[2] == 2.6);
    }
    {  // test/scaled_space_test.cpp:40-52
        auto s = lp(3, 2, 5.0 / 2.0);
        double a[3] = {1, 2, 3}, b[3] = {1, 0, -1};
        CHECK("scaled_space_test:distance == sqrt(20)*5/2",
              distance<double>(s, a, b) == std::sqrt(0.0 + 2.0 * 2.0 + 4.0 * 4.0) * 5 / 2);
    }
    {  // test/so2_space_test.cpp:39-51
        auto s = so2(1, 1);
        auto d = [&](double x, double y) { return distance<double>(s, &x, &y); };
        CHECK("so2_space_test:d(1,1)==0", d(1.0, 1.0) == 0);
        CHECK("so2_space_test:d(0,2)==2", d(0.0, 2.0) == 2.0);
        CHECK("so2_space_test:d(2,0)==2", d(2.0, 0.0) == 2.0);
        CHECK("so2_space_test:d(-1,3)==2pi-4", d(-1.0, 3.0) == 2 * M_PI_ - 4.0);
        CHECK("so2_space_test:d(3,-1)==2pi-4", d(3.0, -1.0) == 2 * M_PI_ - 4.0);
        // test/so2_space_test.cpp:54-65
        auto s3 = so2(3, 1);
        double a[3] = {1, 2, 3}, b[3] = {1, 0, -1};
        CHECK("so2_space_test:distance_lp1", distance<double>(s3, a, b) == 0.0 + 2.0 + 2 * M_PI_ - 4);
        // test/so2_space_test.cpp:67-84
        auto ip = [&](double x, double y, double t) {
            double o;
            interpolate<double>(s, &x, &y, t, &o);
            return o;
        };
        CHECK("so2_space_test:interp(1,1,0)==1", ip(1.0, 1.0, 0.0) == 1.0);
        CHECK("so2_space_test:interp(1,1,3)==1", ip(1.0, 1.0, 3.0) == 1.0);
        CHECK("so2_space_test:interp(1,2,.5)==1.5", ip(1.0, 2.0, 0.5) == 1.5);
        CHECK("so2_space_test:interp(1,2,-3)==-2", ip(1.0, 2.0, -3.0) == -2.0);
        CHECK("so2_space_test:interp(-1,2,4)==11-4pi", ip(-1.0, 2.0, 4.0) == 11.0 - 4 * M_PI_);
        CHECK("so2_space_test:interp(5pi/6,-5pi/6,1)", ip(5 * M_PI_ / 6, -5 * M_PI_ / 6, 1.0) == -5 * M_PI_ / 6);
        CHECK("so2_space_test:interp(5pi/6,-5pi/6,2)", ip(5 * M_PI_ / 6, -5 * M_PI_ / 6, 2.0) == -3 * M_PI_ / 6);
        CHECK("so2_space_test:interp(-5pi/6,5pi/6,2)", ip(-5 * M_PI_ / 6, 5 * M_PI_ / 6, 2.0) == 3 * M_PI_ / 6);
    }
    double axis[3] = {-1, 2, 3};
    {
        double n = std::sqrt(14.0);
        for (double& v : axis) v /= n;
    }
    {  // test/so3_space_test.cpp:39-56
        auto s = so3();
        double a[4], b[4];
        angleAxis(-1.0, axis, a);
        angleAxis(2.0, axis, b);
        CHECK("so3_space_test:distance == 3/2 (1e-10)", std::fabs(distance<double>(s, a, b) - 3.0 / 2) < 1e-10);
        // test/so3_space_test.cpp:58-79
        angleAxis(0.5, axis, a);
        angleAxis(2.5, axis, b);
        double c[4], ang, ax[3];
        interpolate<double>(s, a, b, 0.1, c);
        toAngleAxis(c, &ang, ax);
        double e = 0;
        for (int i = 0; i < 3; ++i) e += (ax[i] - axis[i]) * (ax[i] - axis[i]);
        CHECK("so3_space_test:interpolate axis (1e-15)", e < 1e-15);
        CHECK("so3_space_test:interpolate angle (1e-10)", std::fabs(ang - (0.5 + 2 * 0.1)) < 1e-10);
    }
    {  // test/se3_space_test.cpp:45-90, weights (1,1) (5,2) (11,1) (1,13)
        const double W[4][2] = {{1, 1}, {5, 2}, {11, 1}, {1, 13}};
        for (auto& w : W) {
            auto s = se3(w[0], w[1]);
            double a[7], b[7];
            angleAxis(-1.0, axis, a);
            angleAxis(2.0, axis, b);
            a[4] = 1, a[5] = 2, a[6] = 3;
            b[4] = 1, b[5] = 0, b[6] = -1;
            double expected = 3.0 / 2 * w[0] + std::sqrt(0.0 + 2.0 * 2.0 + 4.0 * 4.0) * w[1];
            char name[96];
            std::snprintf(name, sizeof name, "se3_space_test:distance_%g_%g (1e-9)", w[0], w[1]);
            CHECK(name, std::fabs(distance<double>(s, a, b) - expected) < 1e-9);
        }
        // test/se3_space_test.cpp:92-118
        auto s = se3(1, 1);
        double a[7], b[7], c[7];
        angleAxis(0.5, axis, a);
        angleAxis(2.5, axis, b);
        a[4] = 1, a[5] = 2, a[6] = 3;
        b[4] = 1, b[5] = 0, b[6] = -1;
        interpolate<double>(s, a, b, 0.1, c);
        CHECK("se3_space_test:interpolate translation == (1,1.8,2.6)", c[4] == 1.0 && c[5] == 1.8 && c[6] == 2.6);
        double ang, ax[3];
        toAngleAxis(c, &ang, ax);
        double e = 0;
        for (int i = 0; i < 3; ++i) e += (ax[i] - axis[i]) * (ax[i] - axis[i]);
        CHECK("se3_space_test:interpolate axis (1e-15)", e < 1e-15);
        CHECK("se3_space_test:interpolate angle (1e-10)", std::fabs(ang - (0.5 + 2 * 0.1)) < 1e-10);
    }
    {  // test/se2_space_test.cpp:45-84, weights (1,1) (5,2) (11,1) (1,13); state = (x, y, angle)
        const double W[4][2] = {{1, 1}, {5, 2}, {11, 1}, {1, 13}};
        for (auto& w : W) {
            auto s = se2(w[0], w[1]);
            double a[3] = {2, 3, -1.0}, b[3] = {0, -1, 2.0};
            char name[96];
            std::snprintf(name, sizeof name, "se2_space_test:distance_%g_%g (exact)", w[0], w[1]);
            CHECK(name, distance<double>(s, a, b) == 3.0 * w[0] + std::sqrt(0.0 + 2.0 * 2.0 + 4.0 * 4.0) * w[1]);
        }
    }
    {  // space.dimensions(): SE(3) == 6 (rrg_rewire_neighbors.hpp:60)
        CHECK("dimensions(SE3)==6", spaceDimensions(se3(50, 1)) == 6 && spaceScalars(se3(50, 1)) == 7);
    }
    {  // planner_integration_test.hpp:143-149 closed-form sphere scenario: start/goal valid, straight line not
        Shapes<double> sh;
        sh.dim = 3;
        double r = std::sqrt(2.0) * 0.95;
        sh.centres = {0, 0, 0};
        sh.radii = {r};
        double x = (std::sqrt(3.0) - r) / 2;
        double goal[3] = {1 - x, 1 - x, 1 - x}, start[3] = {-(1 - x), -(1 - x), -(1 - x)};
        CHECK("integration scenario: start & goal valid", sh.valid(start) && sh.valid(goal));
        CHECK("integration scenario: straight line blocked", !sh.link(start, goal));
    }
    std::printf("%d failures\n", failures);
    return failures ? 1 : 0;
}
