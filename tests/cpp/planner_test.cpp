// planner_test.cpp -- the reference's planner integration tests (test/planner_integration_test.hpp:219-254,
// test/{prrt,prrt_star,pprm}_integration_test.cpp, test/pack_nearest_test.cpp:39-69,
// test/prrt_star_integration_test.cpp:43-57) re-expressed for the wave planners, plus the
// stronger check of SURVEY.md Appendix B item 4 (every returned edge re-validates).
// Links against libmptg.so (GPU) or the test-only mock of the same C ABI (CPU, host logic only).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "mptg/planner.hpp"
#include "mptg/scenarios.hpp"

using namespace mptg;
using namespace std::literals;

static int failures = 0;
#define EXPECT(cond)                                                        \
    do {                                                                    \
        if (!(cond)) {                                                      \
            std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);     \
            ++failures;                                                     \
        }                                                                   \
    } while (0)

template <typename T, typename Q, typename = void>
struct has_add_goal : std::false_type {};
template <typename T, typename Q>
struct has_add_goal<T, Q, std::void_t<decltype(std::declval<T>().addGoal(std::declval<Q>()))>> : std::true_type {};

template <typename Algorithm>
void testSolvingBasicScenario(const char* name) {
    using Scenario = test::BasicScenario<double, 3>;
    using State = Scenario::State;
    Planner<Scenario, Algorithm> planner(Scenario(), 12345);
    planner.addStart(Scenario::startState());
    if constexpr (has_add_goal<Planner<Scenario, Algorithm>, State>::value) planner.addGoal(Scenario::goalState());
    auto t0 = std::chrono::steady_clock::now();
    planner.solveFor([&] { return planner.solved(); }, 10s);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    planner.printStats();
    EXPECT(planner.solved());
    std::vector<State> solution = planner.solution();
    EXPECT(solution.size() > 2);  // start + end + at least one waypoint around the obstacle
    EXPECT(!solution.empty() && solution[0] == Scenario::startState());
    EXPECT(!solution.empty() && solution.back() == Scenario::goalState());
    auto sit = solution.begin();
    planner.solution([&](const State& a) {
        EXPECT(sit != solution.end() && *sit == a);
        ++sit;
    });
    EXPECT(sit == solution.end());
    // the trajectory callback form (impl/link_trajectory.hpp:76-84): link() answers with a bool here, so the trajectory type
    // is std::monostate; one call per edge of the path, consecutive, from the start to the goal
    std::size_t edgesSeen = 0;
    bool chained = true;
    planner.solution([&](const State& a, const std::monostate&, const State& b, bool /*forward*/) {
        chained = chained && edgesSeen + 1 < solution.size() && a == solution[edgesSeen] && b == solution[edgesSeen + 1];
        ++edgesSeen;
    });
    EXPECT(chained && edgesSeen + 1 == solution.size());
    // every edge of the returned path re-validates on the device (stronger than the reference)
    if (solution.size() >= 2) {
        Context ctx;
        Scenario sc;
        Geometry g = sc.makeGeometry(ctx);
        std::vector<State> from(solution.begin(), solution.end() - 1), to(solution.begin() + 1, solution.end());
        std::vector<std::uint8_t> ok(from.size());
        auto desc = sc.space().desc();
        g.link(&desc, from.data(), to.data(), (std::uint32_t)from.size(), 0.0, ok.data());
        for (auto v : ok) EXPECT(v == 1);
    }
    std::size_t verts = 0, edges = 0;
    struct Visitor {
        std::size_t &v, &e;
        void vertex(const State&) { ++v; }
        void edge(const State&) { ++e; }
    };
    planner.visitGraph(Visitor{verts, edges});
    EXPECT(verts == planner.size());
    std::printf("%s %s: solved=%d in %.3f s, %zu nodes, %zu graph edges, %zu waypoints\n", failures ? "FAIL" : "PASS", name,
                (int)planner.solved(), secs, planner.size(), edges, solution.size());
}

// The spanner property PPRM-IRS maintains (impl/pprm_irs/pprm_irs.hpp:350-368): with keep_dense_edges every validated
// edge is in the roadmap, sparse or dense, and for every DENSE edge (u, v) the sparse roadmap joins u and v by a path
// shorter than stretchWeight x d(u, v) -- checked here with an independent all-pairs computation on a small roadmap.
void testSpannerStretch() {
    using Scenario = test::BasicScenario<double, 3>;
    using State = Scenario::State;
    Planner<Scenario, PPRMIRS<keep_dense_edges<true>, wave_size<32>>> planner(Scenario(), 777);
    planner.setStretchWeight(3.0);
    planner.addStart(Scenario::startState());
    planner.addGoal(Scenario::goalState());
    planner.solve([&] { return planner.size() >= 400; });
    std::vector<State> nodes;
    std::vector<std::pair<std::size_t, std::size_t>> sparse;
    struct Visitor {
        std::vector<State>& nodes;
        std::vector<std::pair<State, State>> edges;
        void vertex(const State& q) { nodes.push_back(q); }
        void edge(const State& q) { edges.push_back({nodes.back(), q}); }
    } visitor{nodes, {}};
    planner.visitGraph(visitor);
    auto indexOf = [&](const State& q) { return (std::size_t)(std::find(nodes.begin(), nodes.end(), q) - nodes.begin()); };
    const std::size_t n = nodes.size();
    Scenario sc;
    std::vector<double> dist(n * n, std::numeric_limits<double>::infinity());
    for (std::size_t i = 0; i < n; ++i) dist[i * n + i] = 0;
    for (auto& e : visitor.edges) {
        const std::size_t a = indexOf(e.first), b = indexOf(e.second);
        dist[a * n + b] = dist[b * n + a] = sc.space().distance(e.first, e.second);
    }
    for (std::size_t k = 0; k < n; ++k)
        for (std::size_t i = 0; i < n; ++i)
            for (std::size_t j = 0; j < n; ++j) dist[i * n + j] = std::min(dist[i * n + j], dist[i * n + k] + dist[k * n + j]);
    std::size_t denseEdges = 0, violations = 0;
    planner.visitDenseEdges([&](const State& a, const State& b) {
        ++denseEdges;
        if (!(dist[indexOf(a) * n + indexOf(b)] < 3.0 * sc.space().distance(a, b) * (1 + 1e-12))) ++violations;
    });
    EXPECT(denseEdges > 0 && denseEdges == 2 * planner.denseEdgeCount());
    EXPECT(violations == 0);
    std::printf("%s PPRM-IRS spanner: %zu nodes, %zu sparse + %zu dense edges, %zu dense edges without a sparse path within the stretch\n",
                failures ? "FAIL" : "PASS", n, visitor.edges.size() / 2, denseEdges / 2, violations);
}

// The Scenario concept's optional members (impl/scenario_sampler.hpp:47-217, impl/scenario_rng.hpp:46-54): a scenario may
// bring its own sample(rng), its own sampler() object, its own RNG type, and isGoal / sampleGoal instead of goal().
struct SampleMethodScenario : test::BasicScenario<double, 3> {
    using RNG = std::mt19937;  // not the default twister of a double space
    mutable std::size_t* calls;
    explicit SampleMethodScenario(std::size_t* c) : calls(c) {}
    State sample(RNG& rng) const {
        ++*calls;
        std::uniform_real_distribution<double> u(-1.0, 1.0);
        State q;
        for (int i = 0; i < 3; ++i) q[i] = u(rng);
        return q;
    }
};
struct SamplerMethodScenario : test::BasicScenario<double, 3> {
    struct Sampler {
        std::size_t* calls;
        template <typename RNG>
        State operator()(RNG& rng) {
            ++*calls;
            std::uniform_real_distribution<double> u(-1.0, 1.0);
            State q;
            for (int i = 0; i < 3; ++i) q[i] = u(rng);
            return q;
        }
    };
    std::size_t* calls;
    explicit SamplerMethodScenario(std::size_t* c) : calls(c) {}
    Sampler sampler() const { return Sampler{calls}; }
};
template <typename Scenario>
void testScenarioSamplerOptions(const char* name) {
    std::size_t calls = 0;
    Planner<Scenario, PRRT<wave_size<32>>> planner(Scenario(&calls), 5);
    planner.addStart(Scenario::startState());
    planner.solveFor([&] { return planner.solved(); }, 10s);
    EXPECT(planner.solved());
    EXPECT(calls >= planner.size() - 1);  // every node but the start came out of the scenario's own sampler
    std::printf("%s scenario sampler option (%s): solved=%d, %zu nodes, %zu samples drawn by the scenario\n", failures ? "FAIL" : "PASS", name,
                (int)planner.solved(), planner.size(), calls);
}

void testPRRTStarInvariants() {
    // cost(node) == cost(parent) + distance(parent, node) up to rounding, after rewiring
    using Scenario = test::BasicScenario<double, 3>;
    Planner<Scenario, PRRTStar<wave_size<256>>> planner(Scenario(), 99);
    planner.addStart(Scenario::startState());
    planner.setRange(0.5);
    planner.solve([&] { return planner.size() > 3000; });
    Scenario sc;
    double worst = 0;
    for (std::uint32_t n = 1; n < planner.size(); ++n) {
        const std::uint32_t p = planner.nodeParent(n);
        EXPECT(p < planner.size());
        const double want = planner.nodeCost(p) + sc.space().distance(planner.nodeState(p), planner.nodeState(n));
        worst = std::max(worst, std::abs(want - planner.nodeCost(n)));
    }
    EXPECT(worst < 1e-9);
    if (planner.solved()) EXPECT(planner.solutionCost() > 0);
    std::printf("%s PRRT* invariants: %zu nodes, worst cost residual %.3g, solution cost %.6f\n", failures ? "FAIL" : "PASS", planner.size(),
                worst, planner.solved() ? planner.solutionCost() : -1.0);
}

// measure of the sampled region of the SE(3) scenario: the reference multiplies the measure of every Scaled part by the
// part's weight and the parts with each other (impl/uniform_sampler_scaled.hpp:59-66, impl/uniform_sampler_cartesian.hpp:
// 93-97; SO(3) is pi^2, impl/uniform_sampler_so3.hpp:79-83; a box its volume).  PRRT*'s r-nearest rewire radius is built
// from it (impl/rrg_rewire_neighbors.hpp:102-122), so a wrong measure silently shrinks every neighbourhood.
void testSE3SamplerMeasure() {
    mptg::BoxBounds<double, 3> box;
    const double lo[3] = {-60.0, -60.0, -40.0}, hi[3] = {60.0, 60.0, 40.0};
    for (int i = 0; i < 3; ++i) box.min_[i] = lo[i], box.max_[i] = hi[i];
    const mptg::SE3Space<double, 50, 1> space;
    const mptg::UniformSampler<mptg::SE3Space<double, 50, 1>, mptg::SE3Bounds<double>> sampler(space, mptg::SE3Bounds<double>(box));
    const double pi = 3.14159265358979323846;
    const double want = (pi * pi * 50.0) * ((120.0 * 120.0 * 80.0) * 1.0);
    EXPECT(std::fabs(sampler.measure() - want) <= 1e-12 * want);
    const mptg::UniformSampler<mptg::SE3Space<double, 1, 1>, mptg::SE3Bounds<double>> unweighted{mptg::SE3Space<double, 1, 1>(), mptg::SE3Bounds<double>(box)};
    EXPECT(std::fabs(unweighted.measure() - want / 50.0) <= 1e-12 * want);
}

// The Nao-cup scenario (demo/nao_cup_planning.cpp:50-153) through the planner classes: the start and goal configurations
// are clear, the straight edge between them is not (about a seventh of it is), trees grow from the start with every node
// clear and every tree edge passing the scenario's own link again.
template <typename Algorithm>
void testNaoCupScenario(const char* name, std::size_t nodes) {
    using Scenario = mptg::demo::NaoCupScenario<double>;
    Scenario scenario;
    Planner<Scenario, Algorithm> planner(scenario, 4321);
    planner.addStart(scenario.start());
    planner.setRange(0.5);
    planner.solve([&] { return planner.size() >= nodes || planner.solved(); });
    EXPECT(planner.size() >= std::min<std::size_t>(nodes, 2));
    mptg::Context ctx(0);
    mptg::Geometry geom = scenario.makeGeometry(ctx);
    std::vector<Scenario::State> from, to;
    struct Visitor {  // the reference's graph visitor protocol: vertex(q), then edge(to) for each of its edges
        std::vector<Scenario::State>& from;
        std::vector<Scenario::State>& to;
        Scenario::State current;
        Visitor(std::vector<Scenario::State>& f, std::vector<Scenario::State>& t) : from(f), to(t) {}
        void vertex(const Scenario::State& q) { current = q; }
        void edge(const Scenario::State& q) { from.push_back(current), to.push_back(q); }
    } visitor(from, to);
    planner.visitGraph(visitor);
    std::vector<std::uint8_t> ok(from.size() + 2), valid(to.size() + 2);
    if (!from.empty()) {
        geom.link(nullptr, from.data(), to.data(), (std::uint32_t)from.size(), 0.0, ok.data());
        geom.valid(to.data(), (std::uint32_t)to.size(), valid.data());
    }
    std::size_t bad = 0;
    for (std::size_t i = 0; i < from.size(); ++i) bad += !(ok[i] && valid[i]);
    EXPECT(bad == 0);
    const Scenario::State ends[2] = {scenario.start(), scenario.goal().state()};
    std::uint8_t endsOk[2], direct;
    geom.valid(ends, 2, endsOk);
    geom.link(nullptr, &ends[0], &ends[1], 1, 0.0, &direct);
    EXPECT(endsOk[0] && endsOk[1] && !direct);
    std::printf("%s %s on the Nao-cup scenario: %zu nodes, %zu edges re-validated (%zu bad), solved %d\n", failures ? "FAIL" : "PASS", name,
                planner.size(), from.size(), bad, (int)planner.solved());
}

int main() {
    testSE3SamplerMeasure();
    testNaoCupScenario<PRRT<wave_size<64>>>("PRRT", 300);
    testNaoCupScenario<PRRTStar<wave_size<64>>>("PRRT*", 200);
    // test/pack_nearest_test.cpp:39-69 analogue: the strategy tag is recognised, absent -> void
    static_assert(std::is_same_v<impl::pack_nearest_t<>, void>);
    static_assert(std::is_same_v<impl::pack_nearest_t<int, report_stats<true>>, void>);
    static_assert(std::is_same_v<impl::pack_nearest_t<GpuBatch>, GpuBatch>);
    static_assert(std::is_same_v<impl::pack_nearest_t<report_stats<true>, GpuBatch, single_threaded>, GpuBatch>);
    // test/prrt_star_integration_test.cpp:43-57: option order does not change the planner type
    using S = test::BasicScenario<double, 3>;
    static_assert(std::is_same_v<Planner<S, PRRTStar<report_stats<true>, rewire_r_nearest, GpuBatch>>,
                                 Planner<S, PRRTStar<GpuBatch, rewire_r_nearest, report_stats<true>>>>);
    static_assert(!std::is_same_v<Planner<S, PRRTStar<rewire_r_nearest>>, Planner<S, PRRTStar<rewire_k_nearest>>>);
    static_assert(std::is_same_v<Planner<S, PRRTStar<>>, Planner<S, PRRTStar<rewire_k_nearest>>>);

    testSolvingBasicScenario<PRRT<report_stats<true>>>("PRRT");
    testSolvingBasicScenario<PRRT<GpuBatch, wave_size<64>>>("PRRT wave 64");
    testSolvingBasicScenario<PRRTStar<report_stats<true>>>("PRRT* k-nearest");
    testSolvingBasicScenario<PRRTStar<rewire_r_nearest>>("PRRT* r-nearest");
    testSolvingBasicScenario<PPRM<report_stats<true>>>("PPRM");
    static_assert(std::is_same_v<impl::scenario_rng_t<test::BasicScenario<double, 3>, double>, std::mt19937_64>);
    static_assert(std::is_same_v<impl::scenario_rng_t<test::BasicScenario<float, 3>, float>, std::mt19937>);  // mersenne_twister.hpp:328-332
    static_assert(std::is_same_v<impl::scenario_rng_t<SampleMethodScenario, double>, std::mt19937>);
    testScenarioSamplerOptions<SampleMethodScenario>("sample(rng) + RNG");
    testScenarioSamplerOptions<SamplerMethodScenario>("sampler()");
    // the reference's PPRM-IRS integration test is the same scenario (test/pprm_irs_integration_test.cpp)
    static_assert(std::is_same_v<Planner<S, PPRMIRS<report_stats<true>, keep_dense_edges<true>>>, Planner<S, PPRMIRS<keep_dense_edges<true>, report_stats<true>>>>);
    static_assert(!std::is_same_v<Planner<S, PPRMIRS<>>, Planner<S, PPRMIRS<keep_dense_edges<true>>>>);
    testSolvingBasicScenario<PPRMIRS<report_stats<true>>>("PPRM-IRS");
    testSolvingBasicScenario<PPRMIRS<keep_dense_edges<true>, wave_size<64>>>("PPRM-IRS keep_dense_edges, wave 64");
    testSpannerStretch();
#ifndef MPTG_TEST_MOCK_BACKEND  // the mock backs the batched calls only; the device-resident planner needs the GPU library
    static_assert(!std::is_same_v<Planner<S, PRRT<device_resident>>, Planner<S, PRRT<>>>);
    static_assert(std::is_same_v<Planner<S, PRRT<device_resident, wave_size<4096>>>, Planner<S, PRRT<wave_size<4096>, device_resident>>>);
    testSolvingBasicScenario<PRRT<device_resident, report_stats<true>, wave_size<4096>, max_nodes<(1 << 18)>>>("PRRT device-resident");
    static_assert(!std::is_same_v<Planner<S, PRRTStar<device_resident>>, Planner<S, PRRTStar<>>>);
    testSolvingBasicScenario<PRRTStar<device_resident, report_stats<true>, wave_size<2048>, max_nodes<(1 << 17)>>>("PRRT* device-resident");
    static_assert(!std::is_same_v<Planner<S, PRRTStar<device_resident, rewire_r_nearest>>, Planner<S, PRRTStar<device_resident>>>);
    testSolvingBasicScenario<PRRTStar<device_resident, rewire_r_nearest, wave_size<2048>, max_nodes<(1 << 17)>>>("PRRT* device-resident r-nearest");
    {   // cost(node) == cost(parent) + distance(parent, node) up to rounding after wave-parallel rewiring; rewiring happened
        using Scenario = test::BasicScenario<double, 3>;
        Planner<Scenario, PRRTStar<device_resident, wave_size<1024>, max_nodes<(1 << 16)>>> planner(Scenario(), 99);
        planner.addStart(Scenario::startState());
        planner.setRange(0.5);
        planner.solve([&] { return planner.size() > 20000; });
        Scenario sc;
        double worst = 0;
        for (std::uint32_t n = 1; n < planner.size(); ++n) {
            const std::uint32_t p = planner.nodeParent(n);
            EXPECT(p < planner.size());
            if (p >= planner.size()) break;
            worst = std::max(worst, std::abs(planner.nodeCost(p) + sc.space().distance(planner.nodeState(p), planner.nodeState(n)) - planner.nodeCost(n)));
        }
        EXPECT(worst < 1e-9);
        EXPECT(planner.rewires() > 0);
        std::printf("%s PRRT* device-resident invariants: %zu nodes, %llu rewires, worst cost residual %.3g, solution cost %.6f\n",
                    failures ? "FAIL" : "PASS", planner.size(), (unsigned long long)planner.rewires(), worst, planner.solved() ? (double)planner.solutionCost() : -1.0);
    }
    static_assert(!std::is_same_v<Planner<S, PPRM<device_resident>>, Planner<S, PPRM<>>>);
    testSolvingBasicScenario<PPRM<device_resident, report_stats<true>, wave_size<512>, max_nodes<(1 << 16)>>>("PPRM device-resident");
    static_assert(!std::is_same_v<Planner<S, PPRMIRS<device_resident>>, Planner<S, PPRM<device_resident>>>);
    testSolvingBasicScenario<PPRMIRS<device_resident, wave_size<512>, max_nodes<(1 << 16)>>>("PPRM-IRS device-resident");
#endif
    testPRRTStarInvariants();
    // error behaviour (impl/prrt/prrt.hpp:197-198, impl/pprm/pprm.hpp:179-180)
    {
        Planner<S, PRRT<>> p(S(), 1);
        bool threw = false;
        try {
            p.solve([] { return true; });
        } catch (const std::runtime_error&) {
            threw = true;
        }
        EXPECT(threw);
    }
    std::printf("%d failures\n", failures);
    return failures ? 1 : 0;
}
