// ref_planner_bench.cpp -- the REFERENCE'S OWN multi-threaded planners (src/mpt PRRT / PRRT*: OpenMP worker pool, one
// worker per thread, shared nearest-neighbour structure, lock-free tree; compiled from /root/reference, never copied)
// timed on the host cores.  BASELINE INFRASTRUCTURE ONLY: bench.py runs this program for the "reference's
// multithreaded CPU planner" figure that BASELINE.json's north_star asks to be reported next to the GPU numbers.
// Stand-ins (ours, oracle/shim): Eigen value types, logging sink, and -- because Nigh is an un-vendored dependency
// that is not on this machine -- a concurrent insert-only kd-tree in Nigh's place (shim/nigh/nigh_kdtree.hpp).  So the
// planner loop, sampling, steering, rewiring and PNG2dScenario::valid / link are the reference's; the nearest-
// neighbour structure is a stand-in of the same kind as its default.  The output says so ("nn": ...).
//
// usage: ref_planner_bench --map file.pgm --start X Y --goal X Y [--goal-radius R] [--range R] [--algo prrt|prrtstar]
//                          [--threads N] [--nodes N] [--time-ms T] [--seed S]
//        ref_planner_bench --arm scene.txt --algo pprm|pprmirs|prrtstar [--threads N] [--nodes N] [--time-ms T] [--seed S]
//            (the reference's LinkManipulatorScenario<double, N>, N = 8 or 16; scene.txt: N radius / N lengths /
//             C / C lines "cx cy r" / start (N angles) / goal (N angles))
// Phase 1 runs until the first solution (or the node / time limit); phase 2 continues to the node / time limit.
#include <omp.h>

#include <algorithm>
#include <array>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <random>
#include <string>
#include <utility>
#include <vector>

#include "mpt_stubs.hpp"
#include "nigh/nigh_kdtree.hpp"

#include <mpt/goal_state.hpp>
#include <mpt/lp_space.hpp>
#include <mpt/planner.hpp>
#include <mpt/prrt.hpp>
#include <mpt/pprm.hpp>
#include <mpt/pprm_irs.hpp>
#include <mpt/prrt_star.hpp>

#include <link_manipulator_scenario.hpp>
#include <png_2d_scenario.hpp>

namespace mpt = unc::robotics::mpt;

struct GridScenario {
    using Base = mpt_demo::PNG2dScenario<double>;
    using Space = Base::Space;
    using Bounds = Base::Bounds;
    using State = Base::State;
    using Distance = Base::Distance;
    using Goal = mpt::GoalState<Space>;
    using RNG = std::mt19937_64;
    std::shared_ptr<Base> base;
    Bounds bounds_;
    Goal goal_;
    GridScenario(std::shared_ptr<Base> b, int w, int h, double radius, const State& goal)
        : base(std::move(b)), bounds_(State(0, 0), State(w - 1, h - 1)), goal_(radius, goal) {}
    bool valid(const State& q) const { return base->valid(q); }
    bool link(const State& a, const State& b) const { return base->link(a, b); }
    const Space& space() const { return base->space(); }
    const Bounds& bounds() const { return bounds_; }
    const Goal& goal() const { return goal_; }
};

struct Options {
    std::string map, arm, algo = "prrtstar";
    double start[2] = {0, 0}, goal[2] = {0, 0}, goalRadius = 1e-6, range = INFINITY;
    int threads = 0;
    std::size_t nodes = 200000;
    double timeMs = 10000;
    std::uint64_t seed = 1;
};

static bool readPgm(const std::string& path, int& w, int& h, std::vector<bool>& obst) {
    std::ifstream f(path, std::ios::binary);
    std::string magic;
    int maxv;
    if (!(f >> magic >> w >> h >> maxv) || magic != "P5") return false;
    f.get();
    std::vector<unsigned char> px((size_t)w * h);
    f.read((char*)px.data(), (std::streamsize)px.size());
    if (!f) return false;
    obst.resize(px.size());
    for (size_t i = 0; i < px.size(); ++i) obst[i] = px[i] != 0;  // non-zero = obstacle
    return true;
}

template <class Algo>
int run(const Options& o, int w, int h, std::vector<bool>& obst) {
    using State = GridScenario::State;
    using Clock = std::chrono::steady_clock;
    const State goal(o.goal[0], o.goal[1]);
    auto base = std::make_shared<mpt_demo::PNG2dScenario<double>>(w, h, goal, obst);
    GridScenario scenario(base, w, h, o.goalRadius, goal);
    mpt::Planner<GridScenario, Algo> planner(scenario, o.seed);
    if (std::isfinite(o.range)) planner.setRange(o.range);
    planner.addStart(State(o.start[0], o.start[1]));
    const auto t0 = Clock::now();
    auto elapsed = [&] { return std::chrono::duration<double>(Clock::now() - t0).count(); };
    planner.solve([&] { return planner.solved() || planner.size() >= o.nodes || elapsed() * 1e3 >= o.timeMs; });
    const double first = elapsed();
    const std::size_t firstNodes = planner.size();
    const bool solvedFirst = planner.solved();
    planner.solve([&] { return planner.size() >= o.nodes || elapsed() * 1e3 >= o.timeMs; });
    const double total = elapsed();
    double cost = -1;
    if (planner.solved()) {
        cost = 0;
        auto path = planner.solution();
        for (std::size_t i = 1; i < path.size(); ++i) cost += scenario.space().distance(path[i - 1], path[i]);
    }
    std::printf("{\"impl\": \"reference planner classes (src/mpt, OpenMP worker pool)\", \"nn\": \"stand-in concurrent kd-tree (Nigh absent)\", "
                "\"algo\": \"%s\", \"threads\": %d, \"solved\": %s, \"first_solution_s\": %.6f, \"first_solution_nodes\": %zu, "
                "\"nodes\": %zu, \"seconds\": %.6f, \"nodes_per_s\": %.1f, \"solution_cost\": %.4f}\n",
                o.algo.c_str(), omp_get_max_threads(), planner.solved() ? "true" : "false", solvedFirst ? first : -1.0, firstNodes,
                planner.size(), total, planner.size() / total, cost);
    return 0;
}

// the reference's N-link arm scenario (demo/link_manipulator_scenario.hpp) under its PPRM / PRRT*
template <int N, template <class...> class AlgoT>
int runArm(const Options& o, const char* algoName) {
    using Scenario = mpt_demo::LinkManipulatorScenario<double, N>;
    using State = typename Scenario::State;
    using Clock = std::chrono::steady_clock;
    std::ifstream f(o.arm);
    int n = 0, nc = 0;
    double radius = 0;
    f >> n >> radius;
    if (!f || n != N) return 2;
    std::vector<double> lengths(N);
    for (double& l : lengths) f >> l;
    f >> nc;
    std::vector<shape::Circle<double>> circles;
    for (int i = 0; i < nc; ++i) {
        double x, y, r;
        f >> x >> y >> r;
        circles.emplace_back(x, y, r);
    }
    State start, goal;
    for (int i = 0; i < N; ++i) f >> start[i];
    for (int i = 0; i < N; ++i) f >> goal[i];
    if (!f) return 2;
    Scenario scenario(goal, circles, lengths, radius);
    mpt::Planner<Scenario, AlgoT<>> planner(scenario, o.seed);
    planner.addStart(start);
    if constexpr (std::is_same_v<AlgoT<>, mpt::PPRM<>> || std::is_same_v<AlgoT<>, mpt::PPRMIRS<>>) planner.addGoal(goal);
    const auto t0 = Clock::now();
    auto elapsed = [&] { return std::chrono::duration<double>(Clock::now() - t0).count(); };
    planner.solve([&] { return planner.solved() || planner.size() >= o.nodes || elapsed() * 1e3 >= o.timeMs; });
    const double first = elapsed();
    const std::size_t firstNodes = planner.size();
    const bool solvedFirst = planner.solved();
    planner.solve([&] { return planner.size() >= o.nodes || elapsed() * 1e3 >= o.timeMs; });
    const double total = elapsed();
    std::printf("{\"impl\": \"reference planner classes (src/mpt, OpenMP worker pool) on the reference's LinkManipulatorScenario<double, %d>\", "
                "\"nn\": \"stand-in concurrent kd-tree (Nigh absent)\", \"algo\": \"%s\", \"threads\": %d, \"solved\": %s, "
                "\"first_solution_s\": %.6f, \"first_solution_nodes\": %zu, \"nodes\": %zu, \"seconds\": %.6f, \"nodes_per_s\": %.1f}\n",
                N, algoName, omp_get_max_threads(), planner.solved() ? "true" : "false", solvedFirst ? first : -1.0, firstNodes, planner.size(), total,
                planner.size() / total);
    return 0;
}

int main(int argc, char** argv) {
    Options o;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&] { return i + 1 < argc ? argv[++i] : (char*)"0"; };
        if (a == "--map") o.map = next();
        else if (a == "--arm") o.arm = next();
        else if (a == "--start") o.start[0] = atof(next()), o.start[1] = atof(next());
        else if (a == "--goal") o.goal[0] = atof(next()), o.goal[1] = atof(next());
        else if (a == "--goal-radius") o.goalRadius = atof(next());
        else if (a == "--range") o.range = atof(next());
        else if (a == "--algo") o.algo = next();
        else if (a == "--threads") o.threads = atoi(next());
        else if (a == "--nodes") o.nodes = (std::size_t)atoll(next());
        else if (a == "--time-ms") o.timeMs = atof(next());
        else if (a == "--seed") o.seed = (std::uint64_t)atoll(next());
        else {
            std::fprintf(stderr, "unknown option %s\n", a.c_str());
            return 2;
        }
    }
    if (o.threads > 0) omp_set_num_threads(o.threads);
    if (!o.arm.empty()) {
        std::ifstream f(o.arm);
        int n = 0;
        f >> n;
        int rc = 2;
        if (o.algo == "pprm") rc = n == 8 ? runArm<8, mpt::PPRM>(o, "pprm") : n == 16 ? runArm<16, mpt::PPRM>(o, "pprm") : 2;
        else if (o.algo == "pprmirs") rc = n == 8 ? runArm<8, mpt::PPRMIRS>(o, "pprmirs") : n == 16 ? runArm<16, mpt::PPRMIRS>(o, "pprmirs") : 2;
        else if (o.algo == "prrtstar") rc = n == 8 ? runArm<8, mpt::PRRTStar>(o, "prrtstar") : n == 16 ? runArm<16, mpt::PRRTStar>(o, "prrtstar") : 2;
        if (rc == 2) std::fprintf(stderr, "cannot run the arm scene %s (N = 8 or 16; algo pprm, pprmirs or prrtstar)\n", o.arm.c_str());
        return rc;
    }
    int w = 0, h = 0;
    std::vector<bool> obst;
    if (!readPgm(o.map, w, h, obst)) {
        std::fprintf(stderr, "cannot read binary PGM %s\n", o.map.c_str());
        return 2;
    }
    if (o.threads > 0) omp_set_num_threads(o.threads);
    if (o.algo == "prrt") return run<mpt::PRRT<>>(o, w, h, obst);
    if (o.algo == "prrtstar") return run<mpt::PRRTStar<>>(o, w, h, obst);
    std::fprintf(stderr, "unknown algorithm %s\n", o.algo.c_str());
    return 2;
}
