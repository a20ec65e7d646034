// reference_planner_parity.cpp -- row a11: the wave planners of include/mptg/planner.hpp against THE
// REFERENCE'S OWN planner classes, in one process, on the same scenario and the same random stream.
//
// Reference side: src/mpt/impl/{prrt,prrt_star,pprm,pprm_irs} compiled from /root/reference (never copied) against the
// stand-in Eigen / Nigh headers under oracle/shim (exhaustive-scan Nigh with the (distance, insertion order) tie
// rule), single_threaded, scenario = the reference's PNG2dScenario::valid / link on a synthetic occupancy grid,
// RNG = std::mt19937_64 (the Scenario's `using RNG`, impl/scenario_rng.hpp:46-54) seeded with a plain integer.
// Our side: Planner<Scenario, Algorithm<wave_size<1>>> over the TEST-ONLY CPU mock of the C ABI
// (tests/cpp/mock_mptg.cpp), same seed.  With one sample per wave the wave planners must consume the generator
// exactly like the reference's worker loop and build the SAME graph: same vertices (bit-identical states), same
// edges, same solution path -- for PRRT, PRRT* (k-nearest and r-nearest rewiring), PPRM and PPRM-IRS (sparse edges).
// TEST INFRASTRUCTURE: only buildable where /root/reference exists (tests/test_host_cpp.py skips it elsewhere).
#include <algorithm>
#include <array>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <memory>
#include <random>
#include <utility>
#include <vector>

// ---- the reference
#include "mpt_stubs.hpp"
#include "nigh/nigh_linear.hpp"

#include <mpt/goal_state.hpp>
#include <mpt/lp_space.hpp>
#include <mpt/planner.hpp>
#include <mpt/pprm.hpp>
#include <mpt/pprm_irs.hpp>
#include <mpt/prrt.hpp>
#include <mpt/prrt_star.hpp>
#include <mpt/se3_space.hpp>

#include <png_2d_scenario.hpp>

// ---- ours
#include "mptg/nigh_binding.hpp"  // the reference-side binding: pack_nearest + nigh::Nigh for the strategy tag mptg::GpuBatch
#include "mptg/planner.hpp"
#include "mptg/scenarios.hpp"

namespace ref = unc::robotics::mpt;

static int failures = 0;
#define EXPECT(cond)                                                    \
    do {                                                                \
        if (!(cond)) {                                                  \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            ++failures;                                                 \
        }                                                               \
    } while (0)

// synthetic occupancy grid: random rectangles and discs
struct Grid {
    int w, h;
    std::vector<std::uint8_t> occ;
    double start[2], goal[2];
};
static Grid makeGrid(int w, int h, unsigned seed) {
    Grid g{w, h, std::vector<std::uint8_t>((size_t)w * h, 0), {0, 0}, {0, 0}};
    std::mt19937 r(seed);
    auto uni = [&](int n) { return (int)(r() % (unsigned)n); };
    for (int k = 0; k < 40; ++k) {
        const int cx = uni(w), cy = uni(h), rx = 4 + uni(w / 12), ry = 4 + uni(h / 12);
        const bool disc = (k & 1) != 0;
        for (int y = std::max(0, cy - ry); y < std::min(h, cy + ry); ++y)
            for (int x = std::max(0, cx - rx); x < std::min(w, cx + rx); ++x)
                if (!disc || (double)(x - cx) * (x - cx) / ((double)rx * rx) + (double)(y - cy) * (y - cy) / ((double)ry * ry) <= 1.0)
                    g.occ[(size_t)w * y + x] = 1;
    }
    auto freeCell = [&](int x0, int y0, double* out) {
        for (int d = 0;; ++d)
            for (int y = std::max(0, y0 - d); y <= std::min(h - 1, y0 + d); ++y)
                for (int x = std::max(0, x0 - d); x <= std::min(w - 1, x0 + d); ++x)
                    if (!g.occ[(size_t)w * y + x]) {
                        out[0] = x, out[1] = y;
                        return;
                    }
    };
    freeCell(w / 8, h / 8, g.start);
    freeCell(w - w / 8, h - h / 8, g.goal);
    return g;
}

// the reference's PNG scenario with bounds that keep x + 0.5 < width (beyond, the reference indexes its bitmap out of
// range) and a caller-chosen goal radius (it hard-wires 1e-6, demo/png_2d_scenario.hpp:99)
struct RefGrid {
    using Base = mpt_demo::PNG2dScenario<double>;
    using Space = Base::Space;
    using Bounds = Base::Bounds;
    using State = Base::State;
    using Distance = Base::Distance;
    using Goal = ref::GoalState<Space>;
    using RNG = std::mt19937_64;
    std::shared_ptr<Base> base;
    Bounds bounds_;
    Goal goal_;
    RefGrid(const Grid& g, double radius)
        : bounds_(State(0, 0), State(g.w - 1, g.h - 1)), goal_(radius, State(g.goal[0], g.goal[1])) {
        std::vector<bool> obst(g.occ.size());
        for (size_t i = 0; i < obst.size(); ++i) obst[i] = g.occ[i] != 0;
        base = std::make_shared<Base>(g.w, g.h, State(g.goal[0], g.goal[1]), obst);
    }
    bool valid(const State& q) const { return base->valid(q); }
    bool link(const State& a, const State& b) const { return base->link(a, b); }
    const Space& space() const { return base->space(); }
    const Bounds& bounds() const { return bounds_; }
    const Goal& goal() const { return goal_; }
};

// The same scenario for the reference's planners, but with BOTH halves of the hot path behind the C ABI: valid / link are
// answered by mptg_valid_batch / mptg_link_batch on the registered grid (one item per call, the reference's calling
// convention), nearest-neighbour search by the strategy tag.  Copied once per worker like any reference scenario; the
// geometry handle is shared.
struct GpuGrid {
    using Space = RefGrid::Space;
    using Bounds = RefGrid::Bounds;
    using State = RefGrid::State;
    using Distance = RefGrid::Distance;
    using Goal = ref::GoalState<Space>;
    using RNG = std::mt19937_64;
    struct Device {
        mptg::Context ctx;
        mptg::Geometry geom;
        Device(const Grid& g) : ctx(-1), geom(mptg::Geometry::grid(ctx, MPTG_F64, g.w, g.h, g.occ.data())) {}
    };
    std::shared_ptr<Device> dev;
    Space space_;
    Bounds bounds_;
    Goal goal_;
    GpuGrid(const Grid& g, double radius)
        : dev(std::make_shared<Device>(g)), bounds_(State(0, 0), State(g.w - 1, g.h - 1)), goal_(radius, State(g.goal[0], g.goal[1])) {}
    bool valid(const State& q) const {
        const double s[2] = {q[0], q[1]};
        std::uint8_t ok = 0;
        dev->geom.valid(s, 1, &ok);
        return ok != 0;
    }
    bool link(const State& a, const State& b) const {
        const double from[2] = {a[0], a[1]}, to[2] = {b[0], b[1]};
        std::uint8_t ok = 0;
        dev->geom.link(nullptr, from, to, 1, 0.0, &ok);
        return ok != 0;
    }
    const Space& space() const { return space_; }
    const Bounds& bounds() const { return bounds_; }
    const Goal& goal() const { return goal_; }
};

struct OurGrid {
    using Space = mptg::L2Space<double, 2>;
    using Bounds = mptg::BoxBounds<double, 2>;
    using State = Space::Type;
    using Distance = Space::Distance;
    using Goal = mptg::GoalState<Space>;
    int w, h;
    Space space_;
    Bounds bounds_;
    Goal goal_;
    std::shared_ptr<const std::vector<std::uint8_t>> occ;
    OurGrid(const Grid& g, double radius)
        : w(g.w), h(g.h), bounds_(State::Zero(), mptg::makeState<double, 2>({double(g.w - 1), double(g.h - 1)})),
          goal_(radius, mptg::makeState<double, 2>({g.goal[0], g.goal[1]})), occ(std::make_shared<const std::vector<std::uint8_t>>(g.occ)) {}
    const Space& space() const { return space_; }
    const Bounds& bounds() const { return bounds_; }
    const Goal& goal() const { return goal_; }
    mptg::Geometry makeGeometry(mptg::Context& ctx) const {
        return mptg::Geometry::grid(ctx, mptg::detail::scalarTag<double>(), w, h, occ->data());
    }
};

template <typename T, typename = void>
struct has_solution_cost : std::false_type {};
template <typename T>
struct has_solution_cost<T, std::void_t<decltype(std::declval<const T&>().solutionCost())>> : std::true_type {};

using P2 = std::array<double, 2>;
using Edge = std::pair<P2, P2>;
struct GraphDump {
    std::vector<P2> vertices;  // visiting order
    std::vector<Edge> edges;   // (vertex, edge target)
    template <class Q>
    void vertex(const Q& q) { vertices.push_back({q[0], q[1]}); }
    template <class Q>
    void edge(const Q& to) { edges.push_back({vertices.back(), {to[0], to[1]}}); }
};

template <class A>
struct is_roadmap : std::bool_constant<std::is_same_v<A, ref::PPRM<ref::single_threaded>> || std::is_same_v<A, ref::PPRMIRS<ref::single_threaded>> ||
                                       std::is_same_v<A, ref::PPRMIRS<ref::single_threaded, ref::keep_dense_edges<true>>>> {};

template <class RefAlgo, class OurAlgo>
void compare(const char* name, const Grid& g, double goalRadius, double goalBias, double range, std::uint64_t seed, std::size_t nodes,
             bool orderedVertices, bool addGoal) {
    using RefState = RefGrid::State;
    using OurState = OurGrid::State;
    const int before = failures;
    ref::Planner<RefGrid, RefAlgo> rp(RefGrid(g, goalRadius), seed);
    mptg::Planner<OurGrid, OurAlgo> op(OurGrid(g, goalRadius), seed);
    if constexpr (!is_roadmap<RefAlgo>::value) {
        rp.setGoalBias(goalBias), op.setGoalBias(goalBias);
        if (std::isfinite(range)) rp.setRange(range), op.setRange(range);
    }
    rp.addStart(RefState(g.start[0], g.start[1]));
    op.addStart(mptg::makeState<double, 2>({g.start[0], g.start[1]}));
    if constexpr (is_roadmap<RefAlgo>::value) {
        if (addGoal) {
            rp.addGoal(RefState(g.goal[0], g.goal[1]));
            op.addGoal(mptg::makeState<double, 2>({g.goal[0], g.goal[1]}));
        }
    }
    rp.solve([&] { return rp.size() >= nodes; });
    op.solve([&] { return op.size() >= nodes; });

    GraphDump rg, og;
    rp.visitGraph(rg);
    op.visitGraph(og);
    EXPECT(rp.size() == op.size());
    EXPECT(rg.vertices.size() == og.vertices.size());
    if (!orderedVertices) {
        std::sort(rg.vertices.begin(), rg.vertices.end());
        std::sort(og.vertices.begin(), og.vertices.end());
    }
    EXPECT(rg.vertices == og.vertices);
    std::sort(rg.edges.begin(), rg.edges.end());
    std::sort(og.edges.begin(), og.edges.end());
    EXPECT(rg.edges.size() == og.edges.size());
    EXPECT(rg.edges == og.edges);
    EXPECT(rp.solved() == op.solved());
    std::vector<RefState> rs = rp.solution();
    std::vector<OurState> os = op.solution();
    EXPECT(rs.size() == os.size());
    for (std::size_t i = 0; i < rs.size() && i < os.size(); ++i) EXPECT(rs[i][0] == os[i][0] && rs[i][1] == os[i][1]);
    if constexpr (has_solution_cost<ref::Planner<RefGrid, RefAlgo>>::value && has_solution_cost<mptg::Planner<OurGrid, OurAlgo>>::value) {
        if (rp.solved() && op.solved()) EXPECT(rp.solutionCost() == op.solutionCost());  // PRRT*: cost after rewiring, bit for bit
    }
    std::printf("%s %s: %zu / %zu vertices, %zu / %zu edges, solved %d / %d, %zu / %zu waypoints (reference / ours)\n",
                failures == before ? "PASS" : "FAIL", name, rg.vertices.size(), og.vertices.size(), rg.edges.size(), og.edges.size(),
                (int)rp.solved(), (int)op.solved(), rs.size(), os.size());
}

// The reference's OWN planner class, unmodified, with its nearest-neighbour strategy switched to mptg::GpuBatch through
// include/mptg/nigh_binding.hpp (every nn_.nearest / nn_.insert of its loop goes through the C ABI) against the same class
// with the stand-in exhaustive-scan Nigh: same random stream, the graphs must be identical.
template <class RefAlgoLinear, class RefAlgoGpu, class GpuScenario = RefGrid>
void compareStrategies(const char* name, const Grid& g, double goalRadius, double goalBias, double range, std::uint64_t seed, std::size_t nodes) {
    using RefState = RefGrid::State;
    const int before = failures;
    ref::Planner<RefGrid, RefAlgoLinear> a(RefGrid(g, goalRadius), seed);
    ref::Planner<GpuScenario, RefAlgoGpu> b(GpuScenario(g, goalRadius), seed);
    static_assert(!std::is_same_v<decltype(a), decltype(b)>, "the strategy tag must select another planner type");
    if constexpr (!is_roadmap<RefAlgoLinear>::value) {
        a.setGoalBias(goalBias), b.setGoalBias(goalBias);
        if (std::isfinite(range)) a.setRange(range), b.setRange(range);
    } else {
        a.addGoal(RefState(g.goal[0], g.goal[1])), b.addGoal(RefState(g.goal[0], g.goal[1]));
    }
    a.addStart(RefState(g.start[0], g.start[1]));
    b.addStart(RefState(g.start[0], g.start[1]));
    a.solve([&] { return a.size() >= nodes; });
    b.solve([&] { return b.size() >= nodes; });
    GraphDump ga, gb;
    a.visitGraph(ga);
    b.visitGraph(gb);
    std::sort(ga.vertices.begin(), ga.vertices.end()), std::sort(gb.vertices.begin(), gb.vertices.end());
    std::sort(ga.edges.begin(), ga.edges.end()), std::sort(gb.edges.begin(), gb.edges.end());
    EXPECT(ga.vertices == gb.vertices);
    EXPECT(ga.edges == gb.edges);
    EXPECT(a.solved() == b.solved());
    std::vector<RefState> sa = a.solution(), sb = b.solution();
    EXPECT(sa.size() == sb.size());
    for (std::size_t i = 0; i < sa.size() && i < sb.size(); ++i) EXPECT(sa[i][0] == sb[i][0] && sa[i][1] == sb[i][1]);
    std::printf("%s reference %s with mptg::GpuBatch: %zu / %zu vertices, %zu / %zu edges, solved %d / %d (stand-in Nigh / libmptg through the binding)\n",
                failures == before ? "PASS" : "FAIL", name, ga.vertices.size(), gb.vertices.size(), ga.edges.size(), gb.edges.size(), (int)a.solved(), (int)b.solved());
}

// the binding's space descriptor and state codec on a compound space: the reference's SE3Space<double, 50> (tuple state,
// Cartesian<Scaled<SO3, 50>, L2>) -- nearest and k-nearest through libmptg against the stand-in exhaustive scan.  The two
// evaluate acos differently (libm there, mptg_fpmath.h here): distances agree to a few ulps, the order wherever
// neighbours are not closer to each other than that.
void compareSE3Binding() {
    using Space = ref::SE3Space<double, 50, 1>;
    using State = Space::Type;
    struct Key {
        const std::vector<State>* states;
        const State& operator()(std::size_t i) const { return (*states)[i]; }
    };
    const int before = failures;
    std::mt19937_64 rng(77);
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    auto randomState = [&] {
        double q[4], n = 0;
        for (double& c : q) c = u(rng), n += c * c;
        n = std::sqrt(n);
        return State(Eigen::Quaternion<double>(q[3] / n, q[0] / n, q[1] / n, q[2] / n), Eigen::Matrix<double, 3, 1>(100 * u(rng), 100 * u(rng), 100 * u(rng)));
    };
    std::vector<State> states;
    for (int i = 0; i < 3000; ++i) states.push_back(randomState());
    Key key{&states};
    unc::robotics::nigh::Nigh<std::size_t, Space, Key, unc::robotics::nigh::NoThreadSafety, unc::robotics::nigh::Linear> linear(Space(), key);
    unc::robotics::nigh::Nigh<std::size_t, Space, Key, unc::robotics::nigh::NoThreadSafety, mptg::GpuBatch> gpu(Space(), key);
    for (std::size_t i = 0; i < states.size(); ++i) linear.insert(i), gpu.insert(i);
    EXPECT(gpu.size() == linear.size());
    std::size_t sameNearest = 0, sameLists = 0;
    double worst = 0;
    const int Q = 200;
    for (int t = 0; t < Q; ++t) {
        const State q = randomState();
        auto a = linear.nearest(q), b = gpu.nearest(q);
        EXPECT(a && b);
        if (a && b) {
            sameNearest += a->first == b->first;
            worst = std::max(worst, std::abs(a->second - b->second) / a->second);
        }
        std::vector<std::tuple<double, std::size_t>> la, lb;
        linear.nearest(la, q, 12), gpu.nearest(lb, q, 12);
        EXPECT(la.size() == 12 && lb.size() == 12);
        bool same = la.size() == lb.size();
        for (std::size_t i = 0; same && i < la.size(); ++i) same = std::get<1>(la[i]) == std::get<1>(lb[i]);
        sameLists += same;
        std::vector<std::tuple<std::size_t, double>> ra, rb;  // the (T, Distance) tuple order and the radius form
        linear.nearest(ra, q, std::numeric_limits<std::size_t>::max(), 40.0), gpu.nearest(rb, q, std::numeric_limits<std::size_t>::max(), 40.0);
        EXPECT(ra.size() == rb.size() || ra.size() > 128);
    }
    EXPECT(sameNearest == (std::size_t)Q && sameLists >= (std::size_t)Q - 1 && worst < 1e-12);
    std::printf("%s binding on SE3Space<double, 50>: %zu / %d nearest and %zu / %d 12-nearest lists identical, largest relative distance difference %.2g\n",
                failures == before ? "PASS" : "FAIL", sameNearest, Q, sameLists, Q, worst);
}

int main() {
    const Grid g = makeGrid(400, 300, 5);
    const double inf = std::numeric_limits<double>::infinity();
    compare<ref::PRRT<ref::single_threaded>, mptg::PRRT<mptg::wave_size<1>>>("PRRT range 20", g, 8.0, 0.05, 20.0, 11, 1500, true, false);
    compare<ref::PRRT<ref::single_threaded>, mptg::PRRT<mptg::wave_size<1>>>("PRRT unbounded", g, 1e-6, 0.1, inf, 12, 400, true, false);
    compare<ref::PRRTStar<ref::single_threaded>, mptg::PRRTStar<mptg::wave_size<1>>>("PRRT* k-nearest", g, 8.0, 0.05, 25.0, 13, 1200, true, false);
    compare<ref::PRRTStar<ref::single_threaded, ref::rewire_r_nearest>, mptg::PRRTStar<mptg::wave_size<1>, mptg::rewire_r_nearest>>(
        "PRRT* r-nearest", g, 8.0, 0.05, 25.0, 14, 1200, true, false);
    compare<ref::PPRM<ref::single_threaded>, mptg::PPRM<mptg::wave_size<1>>>("PPRM", g, 1e-6, 0.0, inf, 15, 600, false, true);
    // PPRM with the incremental roadmap spanner (impl/pprm_irs): the sparse edges only, and with the dense edges kept
    compare<ref::PPRMIRS<ref::single_threaded>, mptg::PPRMIRS<mptg::wave_size<1>>>("PPRM-IRS", g, 1e-6, 0.0, inf, 16, 900, false, true);
    compare<ref::PPRMIRS<ref::single_threaded, ref::keep_dense_edges<true>>, mptg::PPRMIRS<mptg::wave_size<1>, mptg::keep_dense_edges<true>>>(
        "PPRM-IRS keep_dense_edges", g, 1e-6, 0.0, inf, 17, 700, false, true);
    compareSE3Binding();
    // the reference's own planners over include/mptg/nigh_binding.hpp
    compareStrategies<ref::PRRT<ref::single_threaded>, ref::PRRT<ref::single_threaded, mptg::GpuBatch>>("PRRT", g, 8.0, 0.05, 20.0, 21, 800);
    compareStrategies<ref::PRRTStar<ref::single_threaded>, ref::PRRTStar<mptg::GpuBatch, ref::single_threaded>>("PRRT*", g, 8.0, 0.05, 25.0, 22, 600);
    compareStrategies<ref::PRRTStar<ref::single_threaded, ref::rewire_r_nearest>, ref::PRRTStar<ref::single_threaded, ref::rewire_r_nearest, mptg::GpuBatch>>(
        "PRRT* r-nearest", g, 8.0, 0.05, 25.0, 23, 600);
    compareStrategies<ref::PPRM<ref::single_threaded>, ref::PPRM<ref::single_threaded, mptg::GpuBatch>>("PPRM", g, 1e-6, 0.0, inf, 24, 400);
    // ... and with the scenario's valid / link behind the C ABI as well: the whole hot path of the reference's loop swapped
    compareStrategies<ref::PRRTStar<ref::single_threaded>, ref::PRRTStar<ref::single_threaded, mptg::GpuBatch>, GpuGrid>("PRRT* (kNN + validity)", g, 8.0, 0.05, 25.0,
                                                                                                                  25, 500);
    compareStrategies<ref::PPRM<ref::single_threaded>, ref::PPRM<ref::single_threaded, mptg::GpuBatch>, GpuGrid>("PPRM (kNN + validity)", g, 1e-6, 0.0, inf, 26, 300);
    std::printf("%d failures\n", failures);
    return failures ? 1 : 0;
}
