// scenarios.hpp -- the reference's demo / test scenarios expressed against the batched geometry
// back-ends.  Same class names, constructor arguments and accessors as the reference; valid() and
// link() are answered on the device for whole batches through the Geometry each scenario registers
// with makeGeometry() (planner.hpp describes the concept).
//
//   Holonomic2DPointScenario   demo/holonomic_2d_point_scenario.hpp:52-125
//   PNG2dScenario              demo/png_2d_scenario.hpp:71-166
//   LinkManipulatorScenario    demo/link_manipulator_scenario.hpp:55-152
//   SE3RigidBodyScenario       demo/se3_rigid_body_scenario.hpp:228-362  (meshes given as triangle soups)
//   BasicScenario              test/planner_integration_test.hpp:128-150 (N-D sphere obstacle)
//   NaoCupScenario             demo/nao_cup_planning.cpp:50-153 (10 joints, sphere / capsule model of demo/nao_cup/src)
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <memory>
#include <vector>

#include "host.hpp"
#include "spaces.hpp"

namespace mptg {
namespace shape {

template <typename Scalar>
struct Circle {  // demo/shape_hierarchy.hpp:210-273
    Scalar cx, cy, r;
    Circle(Scalar x, Scalar y, Scalar radius) : cx(x), cy(y), r(radius) {}
};
template <typename Scalar>
struct Rect {  // demo/shape_hierarchy.hpp:168-208
    Scalar x0, y0, x1, y1;
    Rect(Scalar a, Scalar b, Scalar c, Scalar d) : x0(a), y0(b), x1(c), y1(d) {}
};

}  // namespace shape

namespace demo {

template <typename Scalar = double>
class Holonomic2DPointScenario {
public:
    using Space = L2Space<Scalar, 2>;
    using Bounds = BoxBounds<Scalar, 2>;
    using State = typename Space::Type;
    using Distance = typename Space::Distance;
    using Goal = GoalState<Space>;

private:
    int width_, height_;
    std::vector<shape::Circle<Scalar>> circles_;
    std::vector<shape::Rect<Scalar>> rects_;
    Space space_;
    Bounds bounds_;
    Goal goal_;

public:
    Holonomic2DPointScenario(int width, int height, const std::vector<shape::Circle<Scalar>>& circles,
                             const std::vector<shape::Rect<Scalar>>& rects, State goalState)
        : width_(width), height_(height), circles_(circles), rects_(rects),
          bounds_(State::Zero(), makeState<Scalar, 2>({Scalar(width), Scalar(height)})), goal_(1e-6, goalState) {}
    const Space& space() const { return space_; }
    const Bounds& bounds() const { return bounds_; }
    const Goal& goal() const { return goal_; }
    Geometry makeGeometry(Context& ctx) const {
        std::vector<double> c, r, rc;
        for (auto& k : circles_) c.push_back(k.cx), c.push_back(k.cy), r.push_back(k.r);
        for (auto& k : rects_) rc.push_back(k.x0), rc.push_back(k.y0), rc.push_back(k.x1), rc.push_back(k.y1);
        return Geometry::shapes(ctx, detail::scalarTag<Scalar>(), 2, c, r, rc);
    }
};

template <typename Scalar = double>
class PNG2dScenario {
public:
    using Space = L2Space<Scalar, 2>;
    using Bounds = BoxBounds<Scalar, 2>;
    using State = typename Space::Type;
    using Distance = typename Space::Distance;
    using Goal = GoalState<Space>;

private:
    int width_, height_;
    Space space_;
    Bounds bounds_;
    Goal goal_;
    std::shared_ptr<const std::vector<std::uint8_t>> isObstacle_;  // shared: scenarios are copied

public:
    PNG2dScenario(int width, int height, State goalState, const std::vector<std::uint8_t>& isObstacle)
        : width_(width), height_(height), bounds_(State::Zero(), makeState<Scalar, 2>({Scalar(width), Scalar(height)})),
          goal_(1e-6, goalState), isObstacle_(std::make_shared<const std::vector<std::uint8_t>>(isObstacle)) {}
    const Space& space() const { return space_; }
    const Bounds& bounds() const { return bounds_; }
    const Goal& goal() const { return goal_; }
    int width() const { return width_; }
    int height() const { return height_; }
    Geometry makeGeometry(Context& ctx) const {
        return Geometry::grid(ctx, detail::scalarTag<Scalar>(), width_, height_, isObstacle_->data());
    }
};

template <typename Scalar, int dimensions>
class LinkManipulatorScenario {
    static_assert(dimensions > 0, "There must be at least one arm");

public:
    using Space = L1Space<Scalar, dimensions>;
    using Bounds = BoxBounds<Scalar, dimensions>;
    using State = typename Space::Type;
    using Distance = typename Space::Distance;
    using Goal = GoalState<Space>;

private:
    Space space_;
    Bounds bounds_;
    Goal goal_;
    std::vector<shape::Circle<Scalar>> circles_;
    std::vector<Scalar> armLengths_;
    Scalar radius_;

public:
    LinkManipulatorScenario(State goalState, const std::vector<shape::Circle<Scalar>>& circles, const std::vector<Scalar>& armLengths,
                            Scalar radius)
        : bounds_(State::Constant(-fp::consts<Scalar>::pi()), State::Constant(fp::consts<Scalar>::pi())), goal_(1e-6, goalState),
          circles_(circles), armLengths_(armLengths), radius_(radius) {}
    const Space& space() const { return space_; }
    const Bounds& bounds() const { return bounds_; }
    const Goal& goal() const { return goal_; }
    Geometry makeGeometry(Context& ctx) const {
        std::vector<double> len(armLengths_.begin(), armLengths_.end()), c;
        for (auto& k : circles_) c.push_back(k.cx), c.push_back(k.cy), c.push_back(k.r);
        return Geometry::linkArm(ctx, detail::scalarTag<Scalar>(), len, radius_, c);
    }
};

// The Nao humanoid bringing a ball over a cup (demo/nao_cup_planning.cpp:50-153): L2 over the ten arm joints, the
// reference's joint limits, start and goal configuration (demo/nao_cup/src/naocup.hpp:254-301), goal radius 1e-5 (:82).
// valid / link are nao_clear / nao_link of naocup.hpp on the device (mptg_naocup_create).
template <typename Scalar>
class NaoCupScenario {
public:
    static constexpr int kDimensions = 10;
    using Space = L2Space<Scalar, kDimensions>;
    using Bounds = BoxBounds<Scalar, kDimensions>;
    using State = typename Space::Type;
    using Config = State;
    using Distance = typename Space::Distance;
    using Goal = GoalState<Space>;

private:
    Space space_;
    Bounds bounds_;
    Goal goal_;
    State start_;
    static State config(int which) {
        double c[4][kDimensions];
        check(mptg_naocup_configs(detail::scalarTag<Scalar>(), c[0], c[1], c[2], c[3]), nullptr, "mptg_naocup_configs");
        State q;
        for (int i = 0; i < kDimensions; ++i) q[i] = (Scalar)c[which][i];
        return q;
    }

public:
    NaoCupScenario() : bounds_(config(2), config(3)), goal_(Scalar(1e-5), config(1)), start_(config(0)) {}
    const Space& space() const { return space_; }
    const Bounds& bounds() const { return bounds_; }
    const Goal& goal() const { return goal_; }
    const State& start() const { return start_; }  // nao_init_config (nao_cup_planning.cpp:184)
    Geometry makeGeometry(Context& ctx) const { return Geometry::naoCup(ctx, detail::scalarTag<Scalar>()); }
};

// Triangle soups: 9 floats per triangle.  The robot soup is recentred on its vertex mean, as the
// reference does when it loads the robot mesh (se3_rigid_body_scenario.hpp:181-193).
template <typename Scalar = float>
class SE3RigidBodyScenario {
public:
    static constexpr std::intmax_t SO3_WEIGHT = 50;  // se3_rigid_body_scenario.hpp:54
    using Space = SE3Space<Scalar, SO3_WEIGHT>;
    using Bounds = SE3Bounds<Scalar>;
    using State = typename Space::Type;
    using Distance = typename Space::Distance;
    using Goal = GoalState<Space>;

private:
    std::shared_ptr<const std::vector<float>> environment_, robot_;
    Space space_;
    Bounds bounds_;
    Goal goal_;
    Distance stepSize_;

    // mean over distinct vertices (the reference loads with aiProcess_JoinIdenticalVertices), :181-193
    static std::vector<float> recentre(const std::vector<float>& tris) {
        std::vector<std::array<float, 3>> verts;
        for (std::size_t i = 0; i + 2 < tris.size(); i += 3) verts.push_back({tris[i], tris[i + 1], tris[i + 2]});
        std::sort(verts.begin(), verts.end());
        verts.erase(std::unique(verts.begin(), verts.end()), verts.end());
        double c[3] = {0, 0, 0};
        for (auto& v : verts)
            for (int k = 0; k < 3; ++k) c[k] += v[k];
        for (int k = 0; k < 3; ++k) c[k] /= verts.empty() ? 1.0 : (double)verts.size();
        std::vector<float> out(tris);
        for (std::size_t i = 0; i < out.size(); ++i) out[i] = (float)(out[i] - c[i % 3]);
        return out;
    }

public:
    using State3 = mptg::State<Scalar, 3>;
    SE3RigidBodyScenario(const std::vector<float>& envTris, const std::vector<float>& robotTris, const State& goal, const State3& min,
                         const State3& max, Scalar checkResolution, bool shiftRobotToCentre = true)
        : environment_(std::make_shared<const std::vector<float>>(envTris)),
          robot_(std::make_shared<const std::vector<float>>(shiftRobotToCentre ? recentre(robotTris) : robotTris)),
          bounds_(BoxBounds<Scalar, 3>(min, max)), goal_(1e-6, goal) {
        // :333  ((max - min).norm() + SO3_WEIGHT*pi/2) * checkResolution
        Scalar n2 = 0;
        for (int i = 0; i < 3; ++i) n2 += (max[i] - min[i]) * (max[i] - min[i]);
        stepSize_ = (std::sqrt(n2) + Scalar(SO3_WEIGHT * 3.14159265358979323846 / 2)) * checkResolution;
    }
    const Space& space() const { return space_; }
    const Bounds& bounds() const { return bounds_; }
    const Goal& goal() const { return goal_; }
    double linkStep() const { return (double)stepSize_; }
    Geometry makeGeometry(Context& ctx) const { return Geometry::meshPair(ctx, detail::scalarTag<Scalar>(), *robot_, *environment_); }
};

}  // namespace demo

namespace test {

// test/planner_integration_test.hpp:36-150: [-1,1]^dim box with a central sphere of radius 0.95*sqrt(dim-1)
template <typename Scalar = double, int dimensions = 3>
class BasicScenario {
public:
    using Space = L2Space<Scalar, dimensions>;
    using Bounds = BoxBounds<Scalar, dimensions>;
    using State = typename Space::Type;
    using Distance = typename Space::Distance;
    using Goal = GoalState<Space>;

    static Scalar defaultObstacleRadius() { return std::sqrt((Scalar)(dimensions - 1)) * Scalar(0.95); }
    static State goalState() {
        const Scalar x = (std::sqrt((Scalar)dimensions) - defaultObstacleRadius()) / 2;
        return State::Constant((Scalar)1 - x);
    }
    static State startState() { return -goalState(); }

private:
    Space space_;
    Bounds bounds_{State::Constant(-1), State::Constant(1)};
    Goal goal_{1e-6, goalState()};
    Scalar radius_;

public:
    explicit BasicScenario(Scalar r = defaultObstacleRadius()) : radius_(r) {}
    const Space& space() const { return space_; }
    const Bounds& bounds() const { return bounds_; }
    const Goal& goal() const { return goal_; }
    Geometry makeGeometry(Context& ctx) const {
        return Geometry::shapes(ctx, detail::scalarTag<Scalar>(), dimensions, std::vector<double>(dimensions, 0.0), {(double)radius_});
    }
};

}  // namespace test
}  // namespace mptg
