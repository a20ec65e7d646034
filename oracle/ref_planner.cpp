// ref_planner.cpp -- the REFERENCE'S OWN PLANNERS (src/mpt/impl/prrt/prrt.hpp, prrt_star/prrt_star.hpp, pprm/pprm.hpp
// and everything they pull in: worker pool, samplers, goal bias, steer, node / edge types, rewiring), compiled from
// where they lie under /root/reference, behind a small C interface.  TEST INFRASTRUCTURE ONLY (built into
// oracle/_ref/libref_planner.so; nothing under mpt_b200/ or include/ may use it).
//
// Their external dependencies are absent on this machine, so they are satisfied by stand-ins of OURS under oracle/shim/:
//   Eigen/Dense             value types + element-wise operators
//   nigh/*.hpp              metric spaces (distance arithmetic pinned by the reference's KATs) and an exhaustive-scan
//                           Nigh<> with the (distance, insertion order) tie rule -- Nigh is un-vendored and unpinned
//   mpt/log.hpp, mpt/box_bounds.hpp   replaced through their include guards (logging sink, two corner vectors)
//   png.h                   declarations only
// What runs unmodified is therefore the planner LOOP: Worker::solve / addSample (prrt.hpp:358-452,
// prrt_star.hpp:447-657, pprm.hpp:298-378), UniformBoxSampler (uniform_box_sampler.hpp:60-68), goal-biased sampling
// (prrt.hpp:365-387), GoalState, interpolate, PNG2dScenario::valid / link (demo/png_2d_scenario.hpp).
//
// Randomness.  The reference seeds a std::mt19937_64 per worker from std::random_device; here the Scenario names its
// own generator type (`using RNG`, impl/scenario_rng.hpp:46-54): ReplayRNG serves, word by word, the uniforms of the
// product's counter-based sample stream (include/mptg/mptg.h: uniform 0 of sample g is the goal-bias draw, uniforms
// 1..D are the coordinates).  std::uniform_real_distribution<double> turns one 64-bit word v into v * 2^-64
// (libstdc++ generate_canonical with a 2^64-range generator), so a word m << 11 reproduces the 53-bit uniform m * 2^-53
// exactly, and the reference's samplers then compute u * (max - min) + min like the product's sampler does.
#include <algorithm>
#include <array>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <utility>
#include <vector>

#include "mpt_stubs.hpp"
#include "nigh/nigh_linear.hpp"

#include <mpt/goal_state.hpp>
#include <mpt/lp_space.hpp>
#include <mpt/planner.hpp>
#include <mpt/pprm.hpp>
#include <mpt/prrt.hpp>
#include <mpt/prrt_star.hpp>

#include <png_2d_scenario.hpp>

using namespace unc::robotics;

namespace {

// the uniforms of the sample stream, [nSamples][1 + D] doubles in [0,1) with 53 significant bits
struct Replay {
    const double* u = nullptr;
    uint32_t n = 0, D = 0;
    uint32_t next = 0;       // next sample number to hand out
    uint32_t cur = 0, pos = 0;
    bool biasedLoop = false;  // the worker is still in the goal-biased loop (prrt.hpp:374-387)
    bool overrun = false;
};
Replay* g_replay = nullptr;

struct ReplayRNG {
    using result_type = uint64_t;
    static constexpr result_type min() { return 0; }
    static constexpr result_type max() { return ~uint64_t(0); }
    ReplayRNG() {}
    template <class Seed>
    explicit ReplayRNG(const Seed&) {}
    result_type operator()() {
        Replay& r = *g_replay;
        if (r.pos > r.D) {
            r.overrun = true;
            return 0;
        }
        const double u = r.u[(size_t)r.cur * (1 + r.D) + r.pos++];
        return (uint64_t)(u * 9007199254740992.0) << 11;
    }
};

// The reference's PNG scenario (valid / link / space / bounds are its code) with a caller-chosen goal radius:
// PNG2dScenario hard-wires 1e-6 (demo/png_2d_scenario.hpp:99).
struct GridScenario {
    using Base = mpt_demo::PNG2dScenario<double>;
    using Space = Base::Space;
    using Bounds = Base::Bounds;
    using State = Base::State;
    using Distance = Base::Distance;
    using Goal = mpt::GoalState<Space>;
    using RNG = ReplayRNG;

    std::shared_ptr<Base> base;  // scenarios are copied once per worker (prrt.hpp:341-346)
    Bounds bounds_;
    Goal goal_;

    GridScenario(std::shared_ptr<Base> b, const State& lo, const State& hi, double radius, const State& goal)
        : base(std::move(b)), bounds_(lo, hi), goal_(radius, goal) {}
    bool valid(const State& q) const { return base->valid(q); }
    bool link(const State& a, const State& b) const { return base->link(a, b); }
    const Space& space() const { return base->space(); }
    const Bounds& bounds() const { return bounds_; }
    const Goal& goal() const { return goal_; }
};

using Key = std::pair<uint64_t, uint64_t>;
Key keyOf(const GridScenario::State& q) {
    Key k;
    std::memcpy(&k.first, &q[0], 8);
    std::memcpy(&k.second, &q[1], 8);
    return k;
}

struct TreeVisitor {
    std::vector<GridScenario::State> states, parentStates;
    std::vector<uint8_t> hasParent;
    void vertex(const GridScenario::State& q) {
        states.push_back(q);
        parentStates.push_back(q);
        hasParent.push_back(0);
    }
    void edge(const GridScenario::State& to) {
        parentStates.back() = to;
        hasParent.back() = 1;
    }
};

}  // namespace

extern "C" {

// Planner<GridScenario, PRRT<single_threaded>> on an occupancy grid, fed with the given uniforms, one sample per
// iteration until they are used up.  Returns the tree in insertion order (ObjectPool = std::deque, object_pool.hpp:
// 70-90; visitGraph walks start nodes, then the worker's pool, prrt.hpp:300-316): states, parent indices
// (0xFFFFFFFF for the start), and the first node that satisfies the goal (0xFFFFFFFF if none).
int ref_prrt_grid(int width, int height, const uint8_t* occ, const double* lo, const double* hi, const double* start,
                  const double* goal, double goalRadius, double goalBias, double range, const double* uniforms,
                  uint32_t nSamples, double* statesOut, uint32_t* parentsOut, uint32_t capacity, uint32_t* nNodesOut,
                  uint32_t* goalNodeOut, uint32_t* samplesUsedOut) {
    using State = GridScenario::State;
    std::vector<bool> obst((size_t)width * height);
    for (size_t i = 0; i < obst.size(); ++i) obst[i] = occ[i] != 0;
    auto base = std::make_shared<mpt_demo::PNG2dScenario<double>>(width, height, State(goal[0], goal[1]), obst);
    GridScenario scenario(base, State(lo[0], lo[1]), State(hi[0], hi[1]), goalRadius, State(goal[0], goal[1]));

    Replay replay;
    replay.u = uniforms, replay.n = nSamples, replay.D = 2;
    g_replay = &replay;

    mpt::Planner<GridScenario, mpt::PRRT<mpt::single_threaded>> planner(scenario);
    planner.setGoalBias(goalBias);
    if (std::isfinite(range)) planner.setRange(range);
    planner.addStart(State(start[0], start[1]));

    replay.biasedLoop = goalBias > 0;
    // done() runs once before every iteration -- except that the goal-biased loop, on first seeing a solution,
    // jumps to the unbiased loop, which asks again before sampling (prrt.hpp:376-377,392)
    planner.solve([&]() -> bool {
        if (replay.biasedLoop && planner.solved()) {
            replay.biasedLoop = false;
            return replay.next >= replay.n;  // the unbiased loop asks once more
        }
        if (replay.next >= replay.n) return true;
        replay.cur = replay.next++;
        replay.pos = replay.biasedLoop ? 0 : 1;  // the unbiased loop does not draw the bias uniform
        return false;
    });
    g_replay = nullptr;
    if (replay.overrun) return -2;

    TreeVisitor tv;
    planner.visitGraph(tv);
    const uint32_t n = (uint32_t)tv.states.size();
    *nNodesOut = n;
    *samplesUsedOut = replay.next;
    if (n > capacity) return -1;
    std::map<Key, uint32_t> index;
    for (uint32_t i = 0; i < n; ++i) index.emplace(keyOf(tv.states[i]), i);
    uint32_t goalNode = 0xFFFFFFFFu;
    for (uint32_t i = 0; i < n; ++i) {
        statesOut[2 * (size_t)i] = tv.states[i][0], statesOut[2 * (size_t)i + 1] = tv.states[i][1];
        parentsOut[i] = tv.hasParent[i] ? index.at(keyOf(tv.parentStates[i])) : 0xFFFFFFFFu;
        if (i > 0 && goalNode == 0xFFFFFFFFu && scenario.goal()(scenario.space(), tv.states[i]).first) goalNode = i;
    }
    *goalNodeOut = goalNode;
    return 0;
}
}
