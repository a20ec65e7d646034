// planner.hpp -- Planner<Scenario, Algorithm> with the reference's public API, driven in SAMPLE WAVES
// so the GPU sees large batches (BASELINE.json north_star item 3).
//
// Reference counterparts (paths relative to the reference root):
//   Planner<Scenario,Algorithm>   src/mpt/planner.hpp:41-47 (PlannerResolver)
//   PRRT<Opts...>                 src/mpt/prrt.hpp:81-82,       loop src/mpt/impl/prrt/prrt.hpp:358-452
//   PRRTStar<Opts...>             src/mpt/prrt_star.hpp:95-96,  loop src/mpt/impl/prrt_star/prrt_star.hpp:447-657
//   PPRM<Opts...>                 src/mpt/pprm.hpp:80-81,       loop src/mpt/impl/pprm/pprm.hpp:298-378
//   option tags                   src/mpt/planner_tags.hpp:45-64; nearest tag recognition impl/pack_nearest.hpp:43-83
//   solveFor / solveUntil         src/mpt/impl/planner_base.hpp:58-80
//   k / r rewiring                src/mpt/impl/rrg_rewire_neighbors.hpp:53-61,102-122
// What changes: one host thread draws `wave_size` samples, then runs ONE batched 1-NN, steer, state
// check, edge check (and for PRRT*/PPRM one batched k-NN and one or two batched edge checks) per wave
// through the C ABI, instead of one sample at a time per worker thread.  Samples of one wave do
// not see each other as neighbours -- the same effect as the reference's concurrent workers racing
// on insert.  max_threads<> is accepted and ignored (the wave replaces the worker pool).
//
// Scenario concept (duck typed, as in the reference; impl/scenario_*.hpp):
//   using Space;  const Space& space() const;   const Bounds& bounds() const;
//   goal() -> callable (space, q) -> pair<bool,Distance>  with .state()     [GoalState]   or
//   isGoal(q) -> pair<bool,Distance>  and  sampleGoal(rng) -> State
//   mptg::Geometry makeGeometry(mptg::Context&) const;   // registers the obstacles on the device
//   optional: double linkStep() const;                   // DiscreteMotionValidator step (mesh scenarios)
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <iostream>
#include <limits>
#include <memory>
#include <queue>
#include <random>
#include <type_traits>
#include <utility>
#include <variant>
#include <vector>

#include "host.hpp"
#include "spaces.hpp"
#include "tags.hpp"

namespace mptg {

// ------------------------------------------------------------------ option tags (planner_tags.hpp:45-64)
template <bool report>
struct report_stats : std::bool_constant<report> {};
struct rewire_k_nearest {};
struct rewire_r_nearest {};
template <int threadCount>
struct max_threads {};
using single_threaded = max_threads<1>;
using hardware_concurrency = max_threads<0>;
template <int n>
struct wave_size {};
// the nearest-neighbour strategy tag of this library, GpuBatch (the analogue of nigh::KDTreeBatch<> etc.): tags.hpp
// PRRT<device_resident, ...> / PPRM<device_resident, ...>: keep the tree / roadmap on the GPU (mptg_prrt_*, mptg_pprm_*), sample on the device
struct device_resident {};
template <int n>
struct max_nodes {};

template <typename... Options>
struct PRRT {};
template <typename... Options>
struct PRRTStar {};
template <typename... Options>
struct PPRM {};
// PPRM with the incremental roadmap spanner (src/mpt/pprm_irs.hpp:49-99): options as PPRM plus keep_dense_edges<bool>
template <typename... Options>
struct PPRMIRS {};
template <bool keep>
struct keep_dense_edges : std::bool_constant<keep> {};  // planner_tags.hpp (tag of PPRMIRS, src/mpt/pprm_irs.hpp:58)

namespace impl {

template <typename T, typename... Pack>
constexpr bool pack_contains_v = (std::is_same_v<T, Pack> || ...);

template <template <bool> class Tag, bool def, typename... Pack>
struct pack_bool_tag : std::bool_constant<def> {};
template <template <bool> class Tag, bool def, bool v, typename... Rest>
struct pack_bool_tag<Tag, def, Tag<v>, Rest...> : std::bool_constant<v> {};
template <template <bool> class Tag, bool def, typename U, typename... Rest>
struct pack_bool_tag<Tag, def, U, Rest...> : pack_bool_tag<Tag, def, Rest...> {};
template <template <bool> class Tag, bool def, typename... Pack>
constexpr bool pack_bool_tag_v = pack_bool_tag<Tag, def, Pack...>::value;

template <template <int> class Tag, int def, typename... Pack>
struct pack_int_tag : std::integral_constant<int, def> {};
template <template <int> class Tag, int def, int v, typename... Rest>
struct pack_int_tag<Tag, def, Tag<v>, Rest...> : std::integral_constant<int, v> {};
template <template <int> class Tag, int def, typename U, typename... Rest>
struct pack_int_tag<Tag, def, U, Rest...> : pack_int_tag<Tag, def, Rest...> {};
template <template <int> class Tag, int def, typename... Pack>
constexpr int pack_int_tag_v = pack_int_tag<Tag, def, Pack...>::value;

// impl/pack_nearest.hpp:43-83: the nearest strategy named in an option pack, or void
template <typename... Pack>
struct pack_nearest {
    using type = void;
};
template <typename First, typename... Rest>
struct pack_nearest<First, Rest...> : pack_nearest<Rest...> {};
template <typename... Rest>
struct pack_nearest<GpuBatch, Rest...> {
    using type = GpuBatch;
    static_assert(std::is_void_v<typename pack_nearest<Rest...>::type>, "multiple nearest neighbor strategies");
};
template <typename... Pack>
using pack_nearest_t = typename pack_nearest<Pack...>::type;

// ---- scenario trait resolvers (impl/scenario_goal.hpp:46-88, scenario_goal_sampler.hpp)
template <typename Scenario, typename = void>
struct has_goal_fn : std::false_type {};
template <typename Scenario>
struct has_goal_fn<Scenario, std::void_t<decltype(std::declval<const Scenario&>().goal())>> : std::true_type {};

template <typename Scenario, typename = void>
struct has_link_step : std::false_type {};
template <typename Scenario>
struct has_link_step<Scenario, std::void_t<decltype(std::declval<const Scenario&>().linkStep())>> : std::true_type {};

// ---- random generator and sampler of a scenario (impl/scenario_rng.hpp:46-54, impl/scenario_sampler.hpp:47-217)
// RNG: Scenario::RNG if the scenario names one, else the Mersenne twister of the scalar's width (mersenne_twister.hpp:328-332).
template <typename Scenario, typename Scalar, typename = void>
struct scenario_rng {
    using type = std::conditional_t<(sizeof(Scalar) <= 4), std::mt19937, std::mt19937_64>;
};
template <typename Scenario, typename Scalar>
struct scenario_rng<Scenario, Scalar, std::void_t<typename Scenario::RNG>> {
    using type = typename Scenario::RNG;
};
template <typename Scenario, typename Scalar>
using scenario_rng_t = typename scenario_rng<Scenario, Scalar>::type;

// Sampler, in the reference's order of preference: 1. scenario.sample(rng); 2. scenario.sampler() (an object called with the
// generator); 4. UniformSampler<Space, Bounds> over scenario.bounds().  (3., a Scenario::Sampler type, is declared but left
// unimplemented by the reference, scenario_sampler.hpp:206-216.)
template <typename Scenario, typename RNG, typename = void>
struct has_sample_method : std::false_type {};
template <typename Scenario, typename RNG>
struct has_sample_method<Scenario, RNG, std::void_t<decltype(std::declval<Scenario&>().sample(std::declval<RNG&>()))>> : std::true_type {};
template <typename Scenario, typename = void>
struct has_sampler_method : std::false_type {};
template <typename Scenario>
struct has_sampler_method<Scenario, std::void_t<decltype(std::declval<const Scenario&>().sampler())>> : std::true_type {};

template <typename Scenario>
class SampleMethodSampler {  // scenario_sampler.hpp:143-156
    Scenario& scenario_;

public:
    explicit SampleMethodSampler(Scenario& scenario) : scenario_(scenario) {}
    template <typename RNG>
    decltype(auto) operator()(RNG& rng) {
        return scenario_.sample(rng);
    }
};
template <typename Inner>
struct SamplerMethodSampler : Inner {  // scenario_sampler.hpp:168-175
    template <typename Scenario>
    explicit SamplerMethodSampler(Scenario& scenario) : Inner(scenario.sampler()) {}
};
template <typename Scenario>
struct BoundsSampler : UniformSampler<typename Scenario::Space, std::decay_t<decltype(std::declval<const Scenario&>().bounds())>> {
    using Base = UniformSampler<typename Scenario::Space, std::decay_t<decltype(std::declval<const Scenario&>().bounds())>>;
    explicit BoundsSampler(Scenario& scenario) : Base(scenario.space(), scenario.bounds()) {}
};
template <typename Scenario, typename RNG, typename = void>
struct scenario_sampler {
    using type = BoundsSampler<Scenario>;
};
template <typename Scenario, typename RNG>
struct scenario_sampler<Scenario, RNG, std::enable_if_t<has_sample_method<Scenario, RNG>::value>> {
    using type = SampleMethodSampler<Scenario>;
};
template <typename Scenario, typename RNG>
struct scenario_sampler<Scenario, RNG, std::enable_if_t<!has_sample_method<Scenario, RNG>::value && has_sampler_method<Scenario>::value>> {
    using type = SamplerMethodSampler<std::decay_t<decltype(std::declval<const Scenario&>().sampler())>>;
};
template <typename Scenario, typename RNG>
using scenario_sampler_t = typename scenario_sampler<Scenario, RNG>::type;

template <typename Scenario, typename State>
auto checkGoal(const Scenario& s, const State& q) {
    if constexpr (has_goal_fn<Scenario>::value) return s.goal()(s.space(), q);
    else return s.isGoal(q);
}
template <typename Scenario, typename RNG>
auto sampleGoalState(const Scenario& s, RNG& rng) {
    if constexpr (has_goal_fn<Scenario>::value) {
        (void)rng;
        return s.goal().state();
    } else {
        return s.sampleGoal(rng);
    }
}
template <typename Scenario>
double linkStepOf(const Scenario& s) {
    if constexpr (has_link_step<Scenario>::value) return (double)s.linkStep();
    else return 0.0;
}

template <typename S>
constexpr S E = S(2.71828182845904523536028747135266249775724709369995L);
template <typename S>
constexpr S PI = S(3.14159265358979323846264338327950288419716939937510L);

// solution(fn): the reference accepts a waypoint callback fn(q) or a trajectory callback fn(from, trajectory, to, forward)
// (impl/link_trajectory.hpp:76-110).  The batched back-ends answer link() with a bool, whose stored trajectory type is
// std::monostate (:53-54), so the trajectory form is fn(const State&, const std::monostate&, const State&, bool).
template <typename Fn, typename State>
constexpr bool is_trajectory_callback_v = std::is_invocable_v<Fn&, const State&, const std::monostate&, const State&, bool>;
// tree planners: edges from the start towards the goal, forward = true (impl/prrt/prrt.hpp:236-243)
template <typename State, typename Fn>
void emitSolution(const std::vector<State>& path, Fn& fn) {
    if constexpr (is_trajectory_callback_v<Fn, State>) {
        for (std::size_t i = 1; i < path.size(); ++i) fn(path[i - 1], std::monostate{}, path[i], true);
    } else {
        for (const State& q : path) fn(q);
    }
}
// roadmap planners: an edge was created FROM the node that was being added TO its (older) neighbour, and `forward` says
// whether the path runs along it in that direction (impl/pprm/pprm.hpp:196-215, impl/pprm/edge.hpp)
template <typename State, typename Fn>
void emitRoadmapSolution(const std::vector<State>& states, const std::vector<std::uint32_t>& nodes, Fn& fn) {
    if constexpr (is_trajectory_callback_v<Fn, State>) {
        for (std::size_t i = 1; i < nodes.size(); ++i) fn(states[nodes[i - 1]], std::monostate{}, states[nodes[i]], nodes[i - 1] > nodes[i]);
    } else {
        for (std::uint32_t n : nodes) fn(states[n]);
    }
}

// Wave sizes of the device-resident planners on the way to the configured size.  A tree grows outwards by at most one
// `range` per wave, so the first solution needs a minimum NUMBER of waves whatever their size, and until it is found large
// waves only add nodes (and time per wave) the young tree cannot use: 64, 128, 256 samples, then `hold` (512) samples per
// wave; after 16 waves without a solution the size doubles every fourth wave (a hard problem needs the throughput of
// full waves); once a solution exists it doubles every wave up to the configured size.  Measured on the 3976 x 2603
// map, 12 waves to the first solution in every case: PRRT* 10.5 ms with waves ramped to 8,192 samples, 3.2 ms held at
// 512 (13.4 ms with 8,192-sample waves throughout); PRRT 4.2 -> 0.93 ms (profiles/r2_ramp_caps.txt).
struct WaveRamp {
    std::uint32_t first = 64, hold = 512;
    std::uint32_t now = 0, waves = 0;
    std::uint32_t next(std::uint32_t full, bool solved) {
        if (first == 0) return full;
        ++waves;
        if (now == 0) now = first;
        else if (solved || now < hold || (waves > 16 && (waves - 16) % 4 == 0)) now *= 2;
        const std::uint32_t limit = (solved || waves > 16) ? full : (hold < full ? hold : full);
        if (now > limit) now = limit;
        return now < full ? now : full;
    }
};

struct StageTimer {
    double seconds = 0;
    std::uint64_t calls = 0, items = 0;
    struct Scope {
        StageTimer& t;
        std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
        ~Scope() { t.seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
    };
    Scope time(std::uint64_t n) {
        ++calls;
        items += n;
        return Scope{*this};
    }
};

// ------------------------------------------------------------------ shared wave machinery
template <typename Derived, typename Scenario>
class WavePlannerBase {
protected:
    using Space = typename Scenario::Space;
    using State = typename Space::Type;
    using Distance = typename Space::Distance;
    using RNG = scenario_rng_t<Scenario, Distance>;         // Scenario::RNG, else the twister of the scalar's width
    using Sampler = scenario_sampler_t<Scenario, RNG>;      // scenario.sample(rng) / scenario.sampler() / uniform over bounds()

    Scenario scenario_;
    Context ctx_;
    Geometry geom_;
    mptg_space_desc desc_;
    std::uint32_t capacity_;
    std::unique_ptr<Nearest<std::uint32_t, Space>> nn_;
    std::vector<State> states_;  // host copy of every node's state, index = node id
    RNG rng_;
    Sampler sampler_;
    std::uint32_t wave_;
    double linkStep_;
    StageTimer tNearest1_, tNearestK_, tValid_, tLink_, tSteer_;
    std::uint64_t iterations_ = 0, biasedSamples_ = 0, waves_ = 0;

    WavePlannerBase(const Scenario& scenario, std::uint64_t seed, std::uint32_t wave, int device)
        : scenario_(scenario), ctx_(device), geom_(scenario_.makeGeometry(ctx_)), desc_(scenario_.space().desc()), capacity_(1u << 16),
          nn_(new Nearest<std::uint32_t, Space>(ctx_, scenario_.space(), capacity_)), rng_(seed),
          sampler_(scenario_), wave_(wave), linkStep_(linkStepOf(scenario_)) {}

    // grow the device structure (capacity doubles; states are re-inserted from the host copy)
    void reserve(std::size_t need) {
        if (need <= capacity_) return;
        while (capacity_ < need) capacity_ *= 2;
        nn_.reset(new Nearest<std::uint32_t, Space>(ctx_, scenario_.space(), capacity_));
        std::vector<std::uint32_t> ids(states_.size());
        for (std::uint32_t i = 0; i < ids.size(); ++i) ids[i] = i;
        if (!states_.empty()) nn_->insert(states_.data(), ids.data(), (std::uint32_t)states_.size());
    }
    std::uint32_t addNode(const State& q) {
        reserve(states_.size() + 1);
        const std::uint32_t id = (std::uint32_t)states_.size();
        states_.push_back(q);
        nn_->insert(q, id);
        return id;
    }
    void addNodes(const std::vector<State>& qs) {
        if (qs.empty()) return;
        reserve(states_.size() + qs.size());
        std::vector<std::uint32_t> ids(qs.size());
        for (std::size_t i = 0; i < qs.size(); ++i) ids[i] = (std::uint32_t)(states_.size() + i);
        states_.insert(states_.end(), qs.begin(), qs.end());
        nn_->insert(qs.data(), ids.data(), (std::uint32_t)qs.size());
    }
    void validBatch(const std::vector<State>& qs, std::vector<std::uint8_t>& ok) {
        ok.resize(qs.size());
        if (qs.empty()) return;
        auto scope = tValid_.time(qs.size());
        geom_.valid(qs.data(), (std::uint32_t)qs.size(), ok.data());
    }
    void linkBatch(const std::vector<State>& from, const std::vector<State>& to, std::vector<std::uint8_t>& ok) {
        ok.resize(from.size());
        if (from.empty()) return;
        auto scope = tLink_.time(from.size());
        geom_.link(&desc_, from.data(), to.data(), (std::uint32_t)from.size(), linkStep_, ok.data());
    }
    void steerBatch(const std::vector<State>& near, const std::vector<State>& sample, const std::vector<Distance>& d, Distance range,
                    std::vector<State>& out, std::vector<Distance>* distOut) {
        out.resize(near.size());
        if (distOut) distOut->resize(near.size());
        if (near.empty()) return;
        auto scope = tSteer_.time(near.size());
        check(mptg_steer_batch(ctx_.get(), &desc_, near.data(), sample.data(), d.data(), (std::uint32_t)near.size(), (double)range, out.data(),
                               distOut ? distOut->data() : nullptr),
              ctx_.get(), "mptg_steer_batch");
    }
    void printStageStats() const {
        auto line = [](const char* name, const StageTimer& t) {
            if (t.calls)
                std::clog << "  " << name << ": " << t.calls << " batched calls, " << t.items << " items, " << t.seconds * 1e3 << " ms total, "
                          << (t.items ? t.seconds * 1e6 / t.items : 0.0) << " us/item\n";
        };
        std::clog << "  waves: " << waves_ << " of " << wave_ << " samples, iterations: " << iterations_ << ", biased samples: " << biasedSamples_
                  << ", kernel launches: " << ctx_.launches() << "\n";
        line("nearest1", tNearest1_);
        line("nearestK", tNearestK_);
        line("steer", tSteer_);
        line("valid", tValid_);
        line("validMotion", tLink_);
    }

public:
    // impl/planner_base.hpp:58-80 (the deadline is checked between waves instead of by a timer thread)
    template <typename Rep, typename Period>
    void solveFor(const std::chrono::duration<Rep, Period>& duration) {
        solveUntil(std::chrono::steady_clock::now() + duration);
    }
    template <class Clock, class Duration>
    void solveUntil(const std::chrono::time_point<Clock, Duration>& endTime) {
        static_cast<Derived*>(this)->solve([&] { return Clock::now() >= endTime; });
    }
    template <typename DoneFn, typename Rep, typename Period>
    void solveFor(DoneFn doneFn, const std::chrono::duration<Rep, Period>& duration) {
        solveUntil(doneFn, std::chrono::steady_clock::now() + duration);
    }
    template <typename DoneFn, class Clock, class Duration>
    void solveUntil(DoneFn doneFn, const std::chrono::time_point<Clock, Duration>& endTime) {
        static_cast<Derived*>(this)->solve([&] { return doneFn() || Clock::now() >= endTime; });
    }
    std::size_t size() const { return states_.size(); }
    void setWaveSize(std::uint32_t w) { wave_ = w ? w : 1; }
    std::uint32_t getWaveSize() const { return wave_; }
    const Scenario& scenario() const { return scenario_; }
    Context& context() { return ctx_; }
};

// ------------------------------------------------------------------ PRRT (impl/prrt/prrt.hpp:103-459)
template <typename Scenario, int waveSize, bool reportStats>
class WavePRRT : public WavePlannerBase<WavePRRT<Scenario, waveSize, reportStats>, Scenario> {
    using Base = WavePlannerBase<WavePRRT, Scenario>;
    using typename Base::Distance;
    using typename Base::State;
    static constexpr std::uint32_t NONE = 0xFFFFFFFFu;
    Distance maxDistance_{std::numeric_limits<Distance>::infinity()};
    Distance goalBias_{0.01};
    std::vector<std::uint32_t> parent_;
    std::vector<std::uint32_t> goals_;

public:
    explicit WavePRRT(const Scenario& scenario = Scenario(), std::uint64_t seed = std::random_device{}(), int device = -1)
        : Base(scenario, seed, waveSize, device) {}

    void setGoalBias(Distance bias) { goalBias_ = bias; }
    Distance getGoalBias() const { return goalBias_; }
    void setRange(Distance range) { maxDistance_ = range; }
    Distance getRange() const { return maxDistance_; }

    template <typename... Args>
    void addStart(Args&&... args) {  // :179-187
        this->addNode(State(std::forward<Args>(args)...));
        parent_.push_back(NONE);
    }

    template <typename DoneFn>
    std::enable_if_t<std::is_same_v<bool, std::invoke_result_t<DoneFn>>> solve(DoneFn doneFn) {  // :193-201
        if (this->size() == 0) throw std::runtime_error("there are no valid initial states");
        while (!doneFn()) wave();
    }

    bool solved() const { return !goals_.empty(); }

    std::vector<State> solution() const {  // :252-264
        std::vector<State> path;
        const std::uint32_t g = bestGoal();
        for (std::uint32_t n = g; n != NONE; n = parent_[n]) path.push_back(this->states_[n]);
        std::reverse(path.begin(), path.end());
        return path;
    }
    template <typename Fn>
    void solution(Fn fn) const {  // waypoint or trajectory callback form (:236-260, 266-276)
        impl::emitSolution(solution(), fn);
    }
    void printStats() const {  // :278-289
        std::clog << "nodes in graph: " << this->size() << "\nsolutions: " << goals_.size() << "\n";
        if constexpr (reportStats) this->printStageStats();
    }
    template <typename Visitor>
    void visitGraph(Visitor&& visitor) const {  // :291-308
        for (std::uint32_t n = 0; n < this->states_.size(); ++n) {
            visitor.vertex(this->states_[n]);
            if (parent_[n] != NONE) visitor.edge(this->states_[parent_[n]]);
        }
    }

private:
    std::uint32_t bestGoal() const {  // :203-231 (path cost by summed distances)
        std::uint32_t best = NONE;
        Distance bestCost = std::numeric_limits<Distance>::infinity();
        for (std::uint32_t g : goals_) {
            Distance c = 0;
            for (std::uint32_t n = g; parent_[n] != NONE; n = parent_[n]) c += this->scenario_.space().distance(this->states_[n], this->states_[parent_[n]]);
            if (c < bestCost || best == NONE) bestCost = c, best = g;
        }
        return best;
    }

    // One wave == wave_ iterations of Worker::addSample (:411-452)
    void wave() {
        const std::uint32_t W = this->wave_;
        ++this->waves_;
        this->iterations_ += W;
        std::vector<State> samples(W);
        std::uniform_real_distribution<Distance> uniform01;
        for (std::uint32_t i = 0; i < W; ++i) {
            if (goals_.empty() && goalBias_ > 0 && uniform01(this->rng_) < goalBias_) {  // :365-387
                ++this->biasedSamples_;
                samples[i] = sampleGoalState(this->scenario_, this->rng_);
            } else {
                samples[i] = this->sampler_(this->rng_);
            }
        }
        {
            auto scope = this->tNearest1_.time(W);
            this->nn_->nearest(samples.data(), W, 1);  // :416
        }
        std::vector<State> near, cand;
        std::vector<Distance> d;
        std::vector<std::uint32_t> nearIdx;
        for (std::uint32_t i = 0; i < W; ++i) {
            if (this->nn_->counts()[i] == 0) continue;
            const Distance di = this->nn_->distances()[i];
            if (di == 0) continue;  // :427-428
            nearIdx.push_back(this->nn_->indices()[i]);
            near.push_back(this->states_[nearIdx.back()]);
            cand.push_back(samples[i]);
            d.push_back(di);
        }
        std::vector<State> steered;
        this->steerBatch(near, cand, d, maxDistance_, steered, nullptr);  // :430-434 (d is NOT recomputed)
        std::vector<std::uint8_t> ok;
        this->validBatch(steered, ok);  // :439
        std::vector<State> from, to;
        std::vector<std::uint32_t> fromIdx;
        for (std::size_t i = 0; i < steered.size(); ++i)
            if (ok[i]) from.push_back(near[i]), to.push_back(steered[i]), fromIdx.push_back(nearIdx[i]);
        this->linkBatch(from, to, ok);  // :442
        std::vector<State> fresh;
        for (std::size_t i = 0; i < to.size(); ++i)
            if (ok[i]) {
                const std::uint32_t id = (std::uint32_t)(this->states_.size() + fresh.size());
                fresh.push_back(to[i]);
                parent_.push_back(fromIdx[i]);
                if (checkGoal(this->scenario_, to[i]).first) goals_.push_back(id);  // :443,449-450
            }
        this->addNodes(fresh);  // :446-447
    }
};

// ------------------------------------------------------------------ PRRT* (impl/prrt_star/prrt_star.hpp:158-772)
template <typename Scenario, int waveSize, class Rewire, bool reportStats>
class WavePRRTStar : public WavePlannerBase<WavePRRTStar<Scenario, waveSize, Rewire, reportStats>, Scenario> {
    using Base = WavePlannerBase<WavePRRTStar, Scenario>;
    using typename Base::Distance;
    using typename Base::State;
    static constexpr std::uint32_t NONE = 0xFFFFFFFFu;
    Distance maxDistance_{std::numeric_limits<Distance>::infinity()};
    Distance goalBias_{0.01};
    Distance rewireFactor_{1.1};
    std::vector<std::uint32_t> parent_;
    std::vector<Distance> cost_;
    std::vector<std::vector<std::uint32_t>> children_;
    std::vector<std::uint8_t> isGoal_;
    std::uint32_t solution_ = NONE;
    std::size_t goalCount_ = 0;
    std::uint64_t rewireTests_ = 0, rewireCount_ = 0;

    // rrg_rewire_neighbors.hpp:53-61
    Distance kRRG() const { return rewireFactor_ * E<Distance> * (1 + 1 / static_cast<Distance>(this->scenario_.space().dimensions())); }
    unsigned rewireK(std::size_t n) const { return (unsigned)std::ceil(kRRG() * std::log(Distance(n + 1))); }
    // rrg_rewire_neighbors.hpp:102-122
    Distance rewireRadius(std::size_t n) const {
        const unsigned dim = this->scenario_.space().dimensions();
        const Distance invDim = 1 / Distance(dim);
        const Distance unitBall = std::pow(std::sqrt(PI<Distance>), Distance(dim)) / std::tgamma(Distance(dim) / 2 + 1);
        const Distance rRRG = rewireFactor_ * std::pow(2 * (1 + invDim) * this->sampler_.measure() / unitBall, invDim);
        ++n;
        return rRRG * std::pow(std::log(Distance(n)) / Distance(n), invDim);
    }

public:
    explicit WavePRRTStar(const Scenario& scenario = Scenario(), std::uint64_t seed = std::random_device{}(), int device = -1)
        : Base(scenario, seed, waveSize, device) {}

    void setRewireFactor(Distance f) { rewireFactor_ = f; }
    void setGoalBias(Distance bias) { goalBias_ = bias; }
    Distance getGoalBias() const { return goalBias_; }
    void setRange(Distance range) { maxDistance_ = range; }
    Distance getRange() const { return maxDistance_; }

    template <typename... Args>
    void addStart(Args&&... args) {  // :263-279
        this->addNode(State(std::forward<Args>(args)...));
        parent_.push_back(NONE);
        cost_.push_back(0);
        children_.emplace_back();
        isGoal_.push_back(0);
    }

    template <typename DoneFn>
    std::enable_if_t<std::is_same_v<bool, std::invoke_result_t<DoneFn>>> solve(DoneFn doneFn) {  // :286-309
        if (this->size() == 0) throw std::runtime_error("there are no valid initial states");
        while (!doneFn()) wave();
    }

    bool solved() const { return solution_ != NONE; }
    Distance solutionCost() const { return solved() ? cost_[solution_] : std::numeric_limits<Distance>::quiet_NaN(); }  // :317-322

    std::vector<State> solution() const {  // :325-337
        std::vector<State> path;
        for (std::uint32_t n = solution_; n != NONE; n = parent_[n]) path.push_back(this->states_[n]);
        std::reverse(path.begin(), path.end());
        return path;
    }
    template <typename Fn>
    void solution(Fn fn) const {
        impl::emitSolution(solution(), fn);
    }
    void printStats() const {  // :372-380
        std::clog << "nodes in graph: " << this->size() << "\n";
        if constexpr (reportStats) {
            this->printStageStats();
            std::clog << "  rewire tests: " << rewireTests_ << ", rewires: " << rewireCount_ << ", goals: " << goalCount_ << "\n";
        }
    }
    template <typename Visitor>
    void visitGraph(Visitor&& visitor) const {  // :382-401
        for (std::uint32_t n = 0; n < this->states_.size(); ++n) {
            visitor.vertex(this->states_[n]);
            if (parent_[n] != NONE) visitor.edge(this->states_[parent_[n]]);
        }
    }
    // for tests: tree invariants
    Distance nodeCost(std::uint32_t n) const { return cost_[n]; }
    std::uint32_t nodeParent(std::uint32_t n) const { return parent_[n]; }
    const State& nodeState(std::uint32_t n) const { return this->states_[n]; }

private:
    void noteGoal(std::uint32_t n) {  // foundGoal (:206-219) / setEdge goal branch
        if (solution_ == NONE || cost_[n] < cost_[solution_]) solution_ = n;
    }
    // nonConcurrentPushUpdate (:664-688): subtract delta from the whole subtree
    void pushUpdate(std::uint32_t n, Distance delta) {
        ++rewireCount_;
        std::vector<std::uint32_t> stack{n};
        while (!stack.empty()) {
            const std::uint32_t x = stack.back();
            stack.pop_back();
            if (isGoal_[x]) noteGoal(x);
            for (std::uint32_t c : children_[x]) {
                cost_[c] -= delta;
                stack.push_back(c);
            }
        }
    }
    void reparent(std::uint32_t n, std::uint32_t newParent, Distance newCost) {
        auto& sib = children_[parent_[n]];
        sib.erase(std::find(sib.begin(), sib.end(), n));
        parent_[n] = newParent;
        children_[newParent].push_back(n);
        const Distance delta = cost_[n] - newCost;
        cost_[n] = newCost;
        pushUpdate(n, delta);
    }

    std::size_t truncatedBalls_ = 0;  // r-nearest neighbourhoods cut at MPTG_MAX_K nodes (see the wave below)

public:
    std::size_t truncatedNeighbourhoods() const { return truncatedBalls_; }

    // One wave == wave_ iterations of Worker::addSample (:510-657)
    void wave() {
        const std::uint32_t W = this->wave_;
        ++this->waves_;
        this->iterations_ += W;
        std::vector<State> samples(W);
        std::uniform_real_distribution<Distance> uniform01;
        for (std::uint32_t i = 0; i < W; ++i) {
            if (goalCount_ == 0 && goalBias_ > 0 && uniform01(this->rng_) < goalBias_) {  // :468-481
                ++this->biasedSamples_;
                samples[i] = sampleGoalState(this->scenario_, this->rng_);
            } else {
                samples[i] = this->sampler_(this->rng_);
            }
        }
        {
            auto scope = this->tNearest1_.time(W);
            this->nn_->nearest(samples.data(), W, 1);  // :517
        }
        std::vector<State> near, cand;
        std::vector<Distance> d;
        std::vector<std::uint32_t> nearIdx;
        for (std::uint32_t i = 0; i < W; ++i) {
            if (this->nn_->counts()[i] == 0) continue;
            const Distance di = this->nn_->distances()[i];
            if (di == 0) continue;  // :526-527
            nearIdx.push_back(this->nn_->indices()[i]);
            near.push_back(this->states_[nearIdx.back()]);
            cand.push_back(samples[i]);
            d.push_back(di);
        }
        std::vector<State> steered;
        std::vector<Distance> dNear;
        this->steerBatch(near, cand, d, maxDistance_, steered, &dNear);  // :529-536 (dNear recomputed after steering)
        std::vector<std::uint8_t> ok;
        this->validBatch(steered, ok);  // :539
        std::vector<State> from, to;
        std::vector<std::uint32_t> fromIdx;
        std::vector<Distance> dKeep;
        for (std::size_t i = 0; i < steered.size(); ++i)
            if (ok[i]) from.push_back(near[i]), to.push_back(steered[i]), fromIdx.push_back(nearIdx[i]), dKeep.push_back(dNear[i]);
        this->linkBatch(from, to, ok);  // :545
        // survivors
        std::vector<State> fresh;
        std::vector<std::uint32_t> nearOf;
        std::vector<Distance> dOf;
        for (std::size_t i = 0; i < to.size(); ++i)
            if (ok[i]) fresh.push_back(to[i]), nearOf.push_back(fromIdx[i]), dOf.push_back(dKeep[i]);
        const std::uint32_t S = (std::uint32_t)fresh.size();
        if (S == 0) return;

        // neighbourhoods (:559-562)
        const std::size_t n0 = this->size();
        std::uint32_t k;
        Distance radius = std::numeric_limits<Distance>::infinity();
        if constexpr (std::is_same_v<Rewire, rewire_r_nearest>) {
            k = MPTG_MAX_K;
            radius = rewireRadius(n0);
        } else {
            k = std::min<unsigned>(std::max(1u, rewireK(n0)), MPTG_MAX_K);
        }
        {
            auto scope = this->tNearestK_.time(S);
            this->nn_->nearest(fresh.data(), S, k, radius);
        }
        // copies: addNodes() below may grow (re-create) the device structure that owns these vectors
        const std::vector<std::uint32_t> nIdx = this->nn_->indices();
        const std::vector<Distance> nDist = this->nn_->distances();
        const std::vector<std::uint32_t> nCnt = this->nn_->counts();
        if constexpr (std::is_same_v<Rewire, rewire_r_nearest>) {
            // The reference asks for EVERY node within r (rrg_rewire_neighbors.hpp:125-128); the batched call returns at
            // most MPTG_MAX_K (128) of them, the nearest ones.  A ball that held more is counted and reported with the
            // statistics (dense trees, small free space): its farthest candidates were not offered for rewiring.
            for (std::size_t s = 0; s < S; ++s)
                if (nCnt[s] == k && nDist[s * k + (k - 1)] <= radius) ++truncatedBalls_;
        }

        // candidate parents in (cost + distance) order up to the near node (:565-605), checked in ONE batch
        struct Cand {
            std::uint32_t s, j;  // survivor, neighbour slot
        };
        std::vector<std::vector<std::uint32_t>> order(S);
        std::vector<Cand> pc;
        std::vector<State> pFrom, pTo;
        for (std::uint32_t s = 0; s < S; ++s) {
            rewireTests_ += nCnt[s];
            auto& o = order[s];
            o.resize(nCnt[s]);
            for (std::uint32_t j = 0; j < nCnt[s]; ++j) o[j] = j;
            std::stable_sort(o.begin(), o.end(), [&](std::uint32_t a, std::uint32_t b) {
                return cost_[nIdx[(std::size_t)s * k + a]] + nDist[(std::size_t)s * k + a] < cost_[nIdx[(std::size_t)s * k + b]] + nDist[(std::size_t)s * k + b];
            });
            const Distance parentCost = cost_[nearOf[s]] + dOf[s];
            for (std::uint32_t j : o) {
                const std::uint32_t nb = nIdx[(std::size_t)s * k + j];
                const Distance newCost = cost_[nb] + nDist[(std::size_t)s * k + j];
                if (newCost > parentCost) break;
                if (nb == nearOf[s]) break;
                pc.push_back({s, j});
                pFrom.push_back(this->states_[nb]);
                pTo.push_back(fresh[s]);
            }
        }
        std::vector<std::uint8_t> pOk;
        this->linkBatch(pFrom, pTo, pOk);
        // replay the reference's loop per survivor with the batched answers
        std::vector<std::uint32_t> parentOf(S);
        std::vector<Distance> costOf(S);
        std::vector<std::vector<std::uint8_t>> checked(S);
        {
            std::size_t c = 0;
            for (std::uint32_t s = 0; s < S; ++s) {
                checked[s].assign(nCnt[s], 0);
                std::uint32_t parent = nearOf[s];
                Distance parentCost = cost_[nearOf[s]] + dOf[s];
                for (std::uint32_t j : order[s]) {
                    const std::uint32_t nb = nIdx[(std::size_t)s * k + j];
                    const Distance newCost = cost_[nb] + nDist[(std::size_t)s * k + j];
                    if (newCost > parentCost) break;
                    checked[s][j] = 1;
                    if (nb == nearOf[s]) {
                        parent = nb;
                        parentCost = newCost;
                        break;
                    }
                    const bool valid = pOk[c++] != 0;
                    if (valid) {
                        parent = nb;
                        parentCost = newCost;
                        // the remaining candidates of this survivor were submitted but are not consumed
                        while (c < pc.size() && pc[c].s == s) ++c;
                        break;
                    }
                }
                while (c < pc.size() && pc[c].s == s) ++c;
                parentOf[s] = parent;
                costOf[s] = parentCost;
            }
        }
        // insert (:607-622)
        const std::uint32_t first = (std::uint32_t)this->states_.size();
        for (std::uint32_t s = 0; s < S; ++s) {
            const std::uint32_t id = first + s;
            parent_.push_back(parentOf[s]);
            cost_.push_back(costOf[s]);
            children_.emplace_back();
            children_[parentOf[s]].push_back(id);
            const bool goal = checkGoal(this->scenario_, fresh[s]).first;
            isGoal_.push_back(goal ? 1 : 0);
            if (goal) {
                ++goalCount_;
                noteGoal(id);
            }
        }
        this->addNodes(fresh);
        // rewire (:626-656): edges new -> neighbour that would shorten the neighbour's path, one batch
        std::vector<Cand> rc;
        std::vector<State> rFrom, rTo;
        for (std::uint32_t s = 0; s < S; ++s)
            for (std::uint32_t j = 0; j < nCnt[s]; ++j) {
                if (checked[s][j]) continue;
                const std::uint32_t nb = nIdx[(std::size_t)s * k + j];
                if (cost_[first + s] + nDist[(std::size_t)s * k + j] >= cost_[nb]) continue;
                rc.push_back({s, j});
                rFrom.push_back(fresh[s]);
                rTo.push_back(this->states_[nb]);
            }
        std::vector<std::uint8_t> rOk;
        this->linkBatch(rFrom, rTo, rOk);
        for (std::size_t c = 0; c < rc.size(); ++c) {
            if (!rOk[c]) continue;
            const std::uint32_t id = first + rc[c].s;
            const std::uint32_t nb = nIdx[(std::size_t)rc[c].s * k + rc[c].j];
            const Distance newCost = cost_[id] + nDist[(std::size_t)rc[c].s * k + rc[c].j];
            if (newCost >= cost_[nb]) continue;  // an earlier rewire of this wave already did better
            reparent(nb, id, newCost);
        }
    }
};

// ------------------------------------------------------------------ PPRM (impl/pprm/pprm.hpp:65-389)
// irs = true: PPRM-IRS (impl/pprm_irs/pprm_irs.hpp:350-407) -- a validated edge (sample, neighbour) enters the roadmap as a
// SPARSE edge only if the sparse roadmap does not already join its ends by a path shorter than stretchWeight x its length
// (impl/pprm_irs/shortest_path_check.hpp:111-225); otherwise it is dropped, or kept as a dense edge with keepDense.
template <typename Scenario, int waveSize, bool reportStats, bool irs = false, bool keepDense = false>
class WavePPRM : public WavePlannerBase<WavePPRM<Scenario, waveSize, reportStats, irs, keepDense>, Scenario> {
    using Base = WavePlannerBase<WavePPRM, Scenario>;
    using typename Base::Distance;
    using typename Base::State;
    enum Flags : std::uint8_t { kNone = 0, kStart = 1, kGoal = 2 };  // impl/pprm/component.hpp
    struct Edge {
        std::uint32_t to;
        Distance d;
    };
    std::vector<std::vector<Edge>> adj_;    // PPRM: every edge; PPRM-IRS: the sparse edges (what solution() and visitGraph() walk)
    std::vector<std::vector<Edge>> dense_;  // PPRM-IRS with keep_dense_edges<true>: the edges the spanner left out
    Distance stretchWeight_{5};             // impl/pprm_irs/pprm_irs.hpp:85
    // bounded shortest-path search from the node being added, kept between the checks of its edges (their targets grow
    // with the neighbour distance): cost_[v] is valid where stamp_[v] == search_
    std::vector<Distance> cost_;
    std::vector<std::uint32_t> stamp_;
    std::uint32_t search_ = 0;
    using QItem = std::pair<Distance, std::uint32_t>;
    std::priority_queue<QItem, std::vector<QItem>, std::greater<QItem>> queue_;
    std::size_t sparseChecks_ = 0, sparseKept_ = 0, settled_ = 0;
    std::vector<std::uint32_t> comp_;  // union-find parent
    std::vector<std::uint32_t> compSize_;
    std::vector<std::uint8_t> compFlags_;
    std::vector<std::uint32_t> startNodes_, goalNodes_;
    bool solved_ = false;
    Distance kRRG_;

    std::uint32_t find(std::uint32_t x) {
        while (comp_[x] != x) x = comp_[x] = comp_[comp_[x]];
        return x;
    }
    void merge(std::uint32_t a, std::uint32_t b) {  // :341-362 (union by size; flags OR-ed)
        a = find(a), b = find(b);
        if (a == b) return;
        if (compSize_[a] > compSize_[b]) std::swap(a, b);
        comp_[a] = b;
        compSize_[b] += compSize_[a];
        compFlags_[b] |= compFlags_[a];
        if ((compFlags_[b] & (kStart | kGoal)) == (kStart | kGoal)) solved_ = true;  // component.hpp:97-99
    }

public:
    explicit WavePPRM(const Scenario& scenario = Scenario(), std::uint64_t seed = std::random_device{}(), int device = -1)
        : Base(scenario, seed, waveSize, device), kRRG_(E<Distance> + E<Distance> / this->scenario_.space().dimensions()) {}  // :146

    template <typename... Args>
    void addStart(Args&&... args) {  // :156-164
        std::vector<State> q{State(std::forward<Args>(args)...)};
        addSamples(q, kStart);
    }
    template <typename... Args>
    void addGoal(Args&&... args) {  // :166-169
        std::vector<State> q{State(std::forward<Args>(args)...)};
        addSamples(q, kGoal);
    }

    template <typename DoneFn>
    std::enable_if_t<std::is_same_v<bool, std::invoke_result_t<DoneFn>>> solve(DoneFn doneFn) {  // :171-183
        if (goalNodes_.empty()) {
            std::vector<State> q{sampleGoalState(this->scenario_, this->rng_)};
            addSamples(q, kGoal);
        }
        if (goalNodes_.empty() || startNodes_.empty()) throw std::runtime_error("PPRM requires both start and goal configurations");
        while (!doneFn()) {
            ++this->waves_;
            this->iterations_ += this->wave_;
            std::vector<State> samples(this->wave_);
            for (auto& q : samples) q = this->sampler_(this->rng_);
            addSamples(samples, kNone);
        }
    }
    bool solved() const { return solved_; }
    void setStretchWeight(Distance w) { stretchWeight_ = w; }  // impl/pprm_irs/pprm_irs.hpp:164-166
    std::size_t sparseEdgeChecks() const { return sparseChecks_; }
    std::size_t denseEdgeCount() const {
        std::size_t c = 0;
        for (auto& a : dense_) c += a.size();
        return c / 2;
    }

    // shortest path over the roadmap from any start to any goal (impl/djikstras.hpp, pprm.hpp:218-246), as node indices
    std::vector<std::uint32_t> solutionNodes() const {
        const std::uint32_t NONE = 0xFFFFFFFFu;
        const std::size_t n = this->states_.size();
        std::vector<Distance> dist(n, std::numeric_limits<Distance>::infinity());
        std::vector<std::uint32_t> prev(n, NONE);
        std::vector<std::uint8_t> isGoal(n, 0);
        for (std::uint32_t g : goalNodes_) isGoal[g] = 1;
        using QE = std::pair<Distance, std::uint32_t>;
        std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
        for (std::uint32_t s : startNodes_) dist[s] = 0, pq.push({0, s});
        std::uint32_t hit = NONE;
        while (!pq.empty()) {
            auto [d, u] = pq.top();
            pq.pop();
            if (d > dist[u]) continue;
            if (isGoal[u]) {
                hit = u;
                break;
            }
            for (const Edge& e : adj_[u])
                if (d + e.d < dist[e.to]) dist[e.to] = d + e.d, prev[e.to] = u, pq.push({dist[e.to], e.to});
        }
        std::vector<std::uint32_t> path;
        for (std::uint32_t x = hit; x != NONE; x = prev[x]) path.push_back(x);
        std::reverse(path.begin(), path.end());
        return path;
    }
    std::vector<State> solution() const {
        std::vector<State> path;
        for (std::uint32_t n : solutionNodes()) path.push_back(this->states_[n]);
        return path;
    }
    template <typename Fn>
    void solution(Fn fn) const {  // waypoint or trajectory callback form (:196-246)
        impl::emitRoadmapSolution(this->states_, solutionNodes(), fn);
    }
    void printStats() const {
        std::clog << "nodes in graph: " << this->size() << "\n";
        if constexpr (reportStats) this->printStageStats();
    }
    template <typename Visitor>
    void visitGraph(Visitor&& visitor) const {  // :380-387
        for (std::uint32_t n = 0; n < this->states_.size(); ++n) {
            visitor.vertex(this->states_[n]);
            for (const Edge& e : adj_[n]) visitor.edge(this->states_[e.to]);
        }
    }
    std::size_t edgeCount() const {
        std::size_t c = 0;
        for (auto& a : adj_) c += a.size();
        return c / 2;
    }
    // PPRM-IRS with keep_dense_edges<true>: the edges the spanner left out, each from both of its ends
    template <typename Fn>
    void visitDenseEdges(Fn fn) const {
        for (std::uint32_t n = 0; n < dense_.size(); ++n)
            for (const Edge& e : dense_[n]) fn(this->states_[n], this->states_[e.to]);
    }

private:
    // Worker::addSample for a batch (:298-339)
    void addSamples(const std::vector<State>& samples, std::uint8_t flags) {
        std::vector<std::uint8_t> ok;
        this->validBatch(samples, ok);  // :299
        std::vector<State> keep;
        for (std::size_t i = 0; i < samples.size(); ++i)
            if (ok[i]) keep.push_back(samples[i]);
        if (keep.empty()) return;
        const std::size_t n0 = this->size();
        std::vector<State> fresh;
        std::vector<std::uint32_t> slot;  // index into `keep` (for neighbour rows)
        std::uint32_t k = 0;
        if (n0 > 0) {
            k = std::min<unsigned>(std::max(1, (int)std::ceil(kRRG_ * std::log(Distance(n0 + 1)))), MPTG_MAX_K);  // :302-303
            auto scope = this->tNearestK_.time(keep.size());
            this->nn_->nearest(keep.data(), (std::uint32_t)keep.size(), k);  // :304
        }
        const Distance minDist = std::numeric_limits<Distance>::epsilon();  // :306-308
        for (std::uint32_t i = 0; i < keep.size(); ++i) {
            if (n0 > 0 && this->nn_->counts()[i] > 0 && this->nn_->distances()[(std::size_t)i * k] < minDist) continue;
            fresh.push_back(keep[i]);
            slot.push_back(i);
        }
        // all (sample, neighbour) edges in one batch (:325-326)
        std::vector<State> from, to;
        std::vector<std::pair<std::uint32_t, std::uint32_t>> pairs;  // (fresh index, neighbour node)
        std::vector<Distance> pd;
        if (n0 > 0)
            for (std::uint32_t f = 0; f < fresh.size(); ++f) {
                const std::uint32_t i = slot[f];
                for (std::uint32_t j = 0; j < this->nn_->counts()[i]; ++j) {
                    const std::uint32_t nb = this->nn_->indices()[(std::size_t)i * k + j];
                    from.push_back(fresh[f]);
                    to.push_back(this->states_[nb]);
                    pairs.push_back({f, nb});
                    pd.push_back(this->nn_->distances()[(std::size_t)i * k + j]);
                }
            }
        std::vector<std::uint8_t> eok;
        this->linkBatch(from, to, eok);
        const std::uint32_t first = (std::uint32_t)this->states_.size();
        for (std::uint32_t f = 0; f < fresh.size(); ++f) {
            const std::uint32_t id = first + f;
            std::uint8_t fl = flags;
            if (!(fl & kGoal) && checkGoal(this->scenario_, fresh[f]).first) fl |= kGoal;  // :312-316
            adj_.emplace_back();
            comp_.push_back(id);
            compSize_.push_back(1);
            compFlags_.push_back(fl);
            if (fl & kGoal) goalNodes_.push_back(id);
            if (fl & kStart) startNodes_.push_back(id);
            if ((fl & (kStart | kGoal)) == (kStart | kGoal)) solved_ = true;
        }
        if constexpr (irs) dense_.resize(adj_.size());
        std::uint32_t searchFrom = 0xFFFFFFFFu;
        for (std::size_t e = 0; e < pairs.size(); ++e)
            if (eok[e]) {
                const std::uint32_t id = first + pairs[e].first, nb = pairs[e].second;
                if constexpr (irs) {  // addEdge, impl/pprm_irs/pprm_irs.hpp:350-368; edges of a sample arrive nearest first (:340-342)
                    if (searchFrom != id) beginSearch(searchFrom = id);
                    if (needsSparseEdge(id, nb, stretchWeight_ * pd[e], pd[e])) {
                        adj_[id].push_back({nb, pd[e]});
                        adj_[nb].push_back({id, pd[e]});
                    } else if constexpr (keepDense) {
                        dense_[id].push_back({nb, pd[e]});
                        dense_[nb].push_back({id, pd[e]});
                    } else {
                        continue;
                    }
                } else {
                    adj_[id].push_back({nb, pd[e]});
                    adj_[nb].push_back({id, pd[e]});
                }
                merge(id, nb);  // :327-334
            }
        this->addNodes(fresh);  // :337
    }

    // ShortestPathCheck::reset (impl/pprm_irs/shortest_path_check.hpp:82-96)
    void beginSearch(std::uint32_t from) {
        cost_.resize(adj_.size());
        stamp_.resize(adj_.size(), 0);
        if (++search_ == 0) std::fill(stamp_.begin(), stamp_.end(), 0), search_ = 1;
        queue_ = {};
        cost_[from] = 0, stamp_[from] = search_;
        queue_.push({Distance(0), from});
    }
    // ShortestPathCheck::operator() (:119-225): true when no path of sparse edges from `from` to `v` is shorter than
    // `target`, i.e. the spanner needs the edge; the search is Dijkstra's, resumed where the previous check of the same
    // node stopped (its bound only grows), and a kept edge becomes part of it at once (:219-222).
    bool needsSparseEdge(std::uint32_t from, std::uint32_t v, Distance target, Distance edgeLength) {
        ++sparseChecks_;
        if (stamp_[v] == search_ && cost_[v] < target) return false;  // :133-140
        while (!queue_.empty()) {
            const auto [priority, top] = queue_.top();
            const Distance pathCost = cost_[top];
            if (pathCost >= target) break;  // :159-160
            queue_.pop();
            if (pathCost != priority) continue;  // a stale entry: the node was settled through a shorter path (:166-171)
            ++settled_;
            bool found = top == v;
            for (const Edge& edge : adj_[top]) {
                const Distance d = pathCost + edge.d;
                if (edge.to == v && d < target) found = true;
                if (stamp_[edge.to] == search_ && !(d < cost_[edge.to])) continue;
                cost_[edge.to] = d, stamp_[edge.to] = search_;
                queue_.push({d, edge.to});
            }
            if (found) return false;  // :208-209
        }
        (void)from;
        cost_[v] = edgeLength, stamp_[v] = search_;  // the new sparse edge, for the checks that follow (:219-222)
        queue_.push({edgeLength, v});
        ++sparseKept_;
        return true;
    }
};

// ------------------------------------------------------------------ device-resident PRRT
// Same Planner interface, but nothing of a wave crosses PCIe except two words: samples are drawn on the device
// (counter-based generator, mptg.h "sampling"), the tree lives in device memory (mptg_prrt_*).  Needs a scenario
// whose goal() is a GoalState (state + radius) and whose bounds() are a box (or SE(3) translation box).
namespace detail {
template <typename S, int N>
void fillBounds(const BoxBounds<S, N>& b, int offset, double* lo, double* hi) {
    for (int i = 0; i < N; ++i) lo[offset + i] = (double)b.min()[i], hi[offset + i] = (double)b.max()[i];
}
template <typename S>
void fillBounds(const SE3Bounds<S>& b, int, double* lo, double* hi) {
    fillBounds(b.translation, 4, lo, hi);  // rotation first (se3_space.hpp:53-79): scalars 0..3 are the quaternion
}
}  // namespace detail

template <typename Scenario, int waveSize, int maxNodes, bool reportStats>
class DevicePRRT {
    using Space = typename Scenario::Space;
    using State = typename Space::Type;
    using Distance = typename Space::Distance;
    static constexpr std::uint32_t NONE = 0xFFFFFFFFu;
    Scenario scenario_;
    Context ctx_;
    Geometry geom_;
    mptg_space_desc desc_;
    std::uint64_t seed_;
    mptg_prrt* prrt_ = nullptr;
    Distance maxDistance_{std::numeric_limits<Distance>::infinity()};
    Distance goalBias_{0.01};
    std::vector<State> starts_;
    std::uint32_t wave_ = waveSize, size_ = 0, goalNode_ = NONE;
    impl::WaveRamp ramp_;  // wave sizes until the configured size is reached (time to the first solution)

public:
    // first wave's size (0: full-size waves from the start) and the size held until the first solution is found
    void setWaveRamp(std::uint32_t firstWave, std::uint32_t holdAt = 512) { ramp_ = impl::WaveRamp{firstWave, holdAt}; }

private:

    std::uint64_t waves_ = 0;
    double seconds_ = 0;
    mutable std::vector<State> states_;  // host mirror, refreshed on demand
    mutable std::vector<std::uint32_t> parent_;

    void create() {  // deferred to the first solve() so that setRange / setGoalBias apply
        if (prrt_) return;
        double lo[MPTG_MAX_SCALARS] = {0}, hi[MPTG_MAX_SCALARS] = {0};
        detail::fillBounds(scenario_.bounds(), 0, lo, hi);
        const auto& goal = scenario_.goal();
        const State g = goal.state();
        mptg_prrt_params prm{};
        prm.space = &desc_, prm.lo = lo, prm.hi = hi;
        prm.range = std::isfinite((double)maxDistance_) ? (double)maxDistance_ : 1.7e308;
        prm.goal_bias = (double)goalBias_, prm.goal_state = g.data(), prm.goal_radius = (double)goal.radius();
        prm.link_step = impl::linkStepOf(scenario_), prm.seed = seed_, prm.capacity = (std::uint32_t)maxNodes, prm.max_wave = (std::uint32_t)waveSize;
        check(mptg_prrt_create(ctx_.get(), geom_.get(), &prm, &prrt_), ctx_.get(), "mptg_prrt_create");
        for (const State& q : starts_) check(mptg_prrt_add_start(prrt_, q.data()), ctx_.get(), "mptg_prrt_add_start");
        size_ = (std::uint32_t)starts_.size();
    }
    void mirror() const {
        if (states_.size() == size_ || !prrt_) return;
        states_.resize(size_);
        parent_.resize(size_);
        check(mptg_prrt_get_tree(prrt_, 0, size_, states_.data(), parent_.data()), ctx_.get(), "mptg_prrt_get_tree");
    }

public:
    explicit DevicePRRT(const Scenario& scenario = Scenario(), std::uint64_t seed = std::random_device{}(), int device = -1)
        : scenario_(scenario), ctx_(device), geom_(scenario_.makeGeometry(ctx_)), desc_(scenario_.space().desc()), seed_(seed) {}
    DevicePRRT(const DevicePRRT&) = delete;
    DevicePRRT& operator=(const DevicePRRT&) = delete;
    ~DevicePRRT() {
        if (prrt_) mptg_prrt_destroy(prrt_);
    }

    void setGoalBias(Distance bias) { goalBias_ = bias; }
    Distance getGoalBias() const { return goalBias_; }
    void setRange(Distance range) { maxDistance_ = range; }
    Distance getRange() const { return maxDistance_; }
    void setWaveSize(std::uint32_t w) { wave_ = w ? std::min<std::uint32_t>(w, waveSize) : 1; }

    template <typename... Args>
    void addStart(Args&&... args) {
        starts_.emplace_back(std::forward<Args>(args)...);
        if (prrt_) {
            check(mptg_prrt_add_start(prrt_, starts_.back().data()), ctx_.get(), "mptg_prrt_add_start");
            ++size_;
        }
    }
    template <typename DoneFn>
    std::enable_if_t<std::is_same_v<bool, std::invoke_result_t<DoneFn>>> solve(DoneFn doneFn) {
        if (starts_.empty()) throw std::runtime_error("there are no valid initial states");  // prrt.hpp:197-198
        create();
        const auto t0 = std::chrono::steady_clock::now();
        while (!doneFn() && size_ < (std::uint32_t)maxNodes) {
            check(mptg_prrt_wave(prrt_, ramp_.next(wave_, goalNode_ != NONE), &size_, &goalNode_), ctx_.get(), "mptg_prrt_wave");
            ++waves_;
        }
        seconds_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    template <typename Rep, typename Period>
    void solveFor(const std::chrono::duration<Rep, Period>& duration) {
        solveUntil(std::chrono::steady_clock::now() + duration);
    }
    template <class Clock, class Duration>
    void solveUntil(const std::chrono::time_point<Clock, Duration>& endTime) {
        solve([&] { return Clock::now() >= endTime; });
    }
    template <typename DoneFn, typename Rep, typename Period>
    void solveFor(DoneFn doneFn, const std::chrono::duration<Rep, Period>& duration) {
        const auto endTime = std::chrono::steady_clock::now() + duration;
        solve([&] { return doneFn() || std::chrono::steady_clock::now() >= endTime; });
    }
    bool solved() const { return goalNode_ != NONE; }
    std::size_t size() const { return prrt_ ? size_ : starts_.size(); }
    std::vector<State> solution() const {
        std::vector<State> path;
        if (!solved()) return path;
        mirror();
        for (std::uint32_t n = goalNode_; n != NONE; n = parent_[n]) path.push_back(states_[n]);
        std::reverse(path.begin(), path.end());
        return path;
    }
    template <typename Fn>
    void solution(Fn fn) const {
        impl::emitSolution(solution(), fn);
    }
    template <typename Visitor>
    void visitGraph(Visitor&& visitor) const {
        mirror();
        for (std::uint32_t n = 0; n < states_.size(); ++n) {
            visitor.vertex(states_[n]);
            if (parent_[n] != NONE) visitor.edge(states_[parent_[n]]);
        }
    }
    void printStats() const {
        std::clog << "nodes in graph: " << size() << "\nsolutions: " << (solved() ? 1 : 0) << "\n";
        if constexpr (reportStats)
            std::clog << "  device-resident waves: " << waves_ << " of " << wave_ << " samples, " << (prrt_ ? mptg_prrt_samples_drawn(prrt_) : 0)
                      << " samples drawn, " << seconds_ * 1e3 << " ms in solve(), kernel launches: " << ctx_.launches() << "\n";
    }
    const Scenario& scenario() const { return scenario_; }
    Context& context() { return ctx_; }
};

// ------------------------------------------------------------------ device-resident PRRT* 
// Planner<Scenario, PRRTStar<device_resident, ...>>: tree, costs, parent choice and rewiring on the GPU (mptg_prrtstar_*,
// wave-parallel semantics stated in mptg.h); k-nearest rewiring, or radius rewiring with rewire_r_nearest.
template <typename Scenario, int waveSize, int maxNodes, bool reportStats, bool rNearest = false>
class DevicePRRTStar {
    using Space = typename Scenario::Space;
    using State = typename Space::Type;
    using Distance = typename Space::Distance;
    static constexpr std::uint32_t NONE = 0xFFFFFFFFu;
    Scenario scenario_;
    Context ctx_;
    Geometry geom_;
    mptg_space_desc desc_;
    std::uint64_t seed_;
    mptg_prrtstar* prrt_ = nullptr;
    Distance rewireFactor_{1.1};
    Distance maxDistance_{std::numeric_limits<Distance>::infinity()};
    Distance goalBias_{0.01};
    std::vector<State> starts_;
    std::uint32_t wave_ = waveSize, size_ = 0, goalNode_ = NONE;
    impl::WaveRamp ramp_;  // wave sizes until the configured size is reached (time to the first solution)

public:
    // first wave's size (0: full-size waves from the start) and the size held until the first solution is found
    void setWaveRamp(std::uint32_t firstWave, std::uint32_t holdAt = 512) { ramp_ = impl::WaveRamp{firstWave, holdAt}; }

private:

    std::uint64_t waves_ = 0;
    mutable std::uint64_t mirroredWaves_ = ~0ull;
    double seconds_ = 0;
    mutable std::vector<State> states_;  // host mirror, refreshed on demand
    mutable std::vector<std::uint32_t> parent_;
    mutable std::vector<Distance> cost_;

    void create() {  // deferred to the first solve() so that setRange / setGoalBias apply
        if (prrt_) return;
        double lo[MPTG_MAX_SCALARS] = {0}, hi[MPTG_MAX_SCALARS] = {0};
        detail::fillBounds(scenario_.bounds(), 0, lo, hi);
        const auto& goal = scenario_.goal();
        const State g = goal.state();
        mptg_prrt_params prm{};
        prm.space = &desc_, prm.lo = lo, prm.hi = hi;
        prm.range = std::isfinite((double)maxDistance_) ? (double)maxDistance_ : 1.7e308;
        prm.goal_bias = (double)goalBias_, prm.goal_state = g.data(), prm.goal_radius = (double)goal.radius();
        prm.link_step = impl::linkStepOf(scenario_), prm.seed = seed_, prm.capacity = (std::uint32_t)maxNodes, prm.max_wave = (std::uint32_t)waveSize;
        check(mptg_prrtstar_create(ctx_.get(), geom_.get(), &prm, (double)rewireFactor_, &prrt_), ctx_.get(), "mptg_prrtstar_create");
        if constexpr (rNearest) {  // rrg_rewire_neighbors.hpp:102-122
            const unsigned dim = scenario_.space().dimensions();
            const Distance invDim = 1 / Distance(dim);
            const Distance unitBall = std::pow(std::sqrt(impl::PI<Distance>), Distance(dim)) / std::tgamma(Distance(dim) / 2 + 1);
            const UniformSampler<Space, std::decay_t<decltype(scenario_.bounds())>> sampler(scenario_.space(), scenario_.bounds());
            const Distance rRRG = rewireFactor_ * std::pow(2 * (1 + invDim) * sampler.measure() / unitBall, invDim);
            check(mptg_prrtstar_set_rewire_radius(prrt_, (double)rRRG), ctx_.get(), "mptg_prrtstar_set_rewire_radius");
        }
        for (const State& q : starts_) check(mptg_prrtstar_add_start(prrt_, q.data()), ctx_.get(), "mptg_prrtstar_add_start");
        size_ = (std::uint32_t)starts_.size();
    }
    void mirror() const {
        if (!prrt_ || (states_.size() == size_ && mirroredWaves_ == waves_)) return;  // rewiring changes old nodes too
        mirroredWaves_ = waves_;
        states_.resize(size_);
        parent_.resize(size_);
        cost_.resize(size_);
        check(mptg_prrtstar_get_tree(prrt_, 0, size_, states_.data(), parent_.data(), cost_.data()), ctx_.get(), "mptg_prrtstar_get_tree");
    }

public:
    explicit DevicePRRTStar(const Scenario& scenario = Scenario(), std::uint64_t seed = std::random_device{}(), int device = -1)
        : scenario_(scenario), ctx_(device), geom_(scenario_.makeGeometry(ctx_)), desc_(scenario_.space().desc()), seed_(seed) {}
    DevicePRRTStar(const DevicePRRTStar&) = delete;
    DevicePRRTStar& operator=(const DevicePRRTStar&) = delete;
    ~DevicePRRTStar() {
        if (prrt_) mptg_prrtstar_destroy(prrt_);
    }

    void setRewireFactor(Distance f) { rewireFactor_ = f; }
    void setGoalBias(Distance bias) { goalBias_ = bias; }
    Distance getGoalBias() const { return goalBias_; }
    void setRange(Distance range) { maxDistance_ = range; }
    Distance getRange() const { return maxDistance_; }
    void setWaveSize(std::uint32_t w) { wave_ = w ? std::min<std::uint32_t>(w, waveSize) : 1; }

    template <typename... Args>
    void addStart(Args&&... args) {
        starts_.emplace_back(std::forward<Args>(args)...);
        if (prrt_) {
            check(mptg_prrtstar_add_start(prrt_, starts_.back().data()), ctx_.get(), "mptg_prrtstar_add_start");
            ++size_;
        }
    }
    template <typename DoneFn>
    std::enable_if_t<std::is_same_v<bool, std::invoke_result_t<DoneFn>>> solve(DoneFn doneFn) {
        if (starts_.empty()) throw std::runtime_error("there are no valid initial states");  // prrt.hpp:197-198
        create();
        const auto t0 = std::chrono::steady_clock::now();
        while (!doneFn() && size_ < (std::uint32_t)maxNodes) {
            check(mptg_prrtstar_wave(prrt_, ramp_.next(wave_, goalNode_ != NONE), &size_, &goalNode_), ctx_.get(), "mptg_prrtstar_wave");
            ++waves_;
        }
        seconds_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    template <typename Rep, typename Period>
    void solveFor(const std::chrono::duration<Rep, Period>& duration) {
        solveUntil(std::chrono::steady_clock::now() + duration);
    }
    template <class Clock, class Duration>
    void solveUntil(const std::chrono::time_point<Clock, Duration>& endTime) {
        solve([&] { return Clock::now() >= endTime; });
    }
    template <typename DoneFn, typename Rep, typename Period>
    void solveFor(DoneFn doneFn, const std::chrono::duration<Rep, Period>& duration) {
        const auto endTime = std::chrono::steady_clock::now() + duration;
        solve([&] { return doneFn() || std::chrono::steady_clock::now() >= endTime; });
    }
    bool solved() const { return goalNode_ != NONE; }
    Distance solutionCost() const {  // prrt_star.hpp:317-322
        if (!solved()) return std::numeric_limits<Distance>::quiet_NaN();
        mirror();
        return cost_[goalNode_];
    }
    // for tests: tree invariants
    Distance nodeCost(std::uint32_t n) const { return mirror(), cost_[n]; }
    std::uint32_t nodeParent(std::uint32_t n) const { return mirror(), parent_[n]; }
    const State& nodeState(std::uint32_t n) const { return mirror(), states_[n]; }
    std::uint64_t rewires() const { return prrt_ ? mptg_prrtstar_rewires(prrt_) : 0; }
    std::size_t size() const { return prrt_ ? size_ : starts_.size(); }
    std::vector<State> solution() const {
        std::vector<State> path;
        if (!solved()) return path;
        mirror();
        for (std::uint32_t n = goalNode_; n != NONE; n = parent_[n]) path.push_back(states_[n]);
        std::reverse(path.begin(), path.end());
        return path;
    }
    template <typename Fn>
    void solution(Fn fn) const {
        impl::emitSolution(solution(), fn);
    }
    template <typename Visitor>
    void visitGraph(Visitor&& visitor) const {
        mirror();
        for (std::uint32_t n = 0; n < states_.size(); ++n) {
            visitor.vertex(states_[n]);
            if (parent_[n] != NONE) visitor.edge(states_[parent_[n]]);
        }
    }
    void printStats() const {
        std::clog << "nodes in graph: " << size() << "\nsolutions: " << (solved() ? 1 : 0) << "\n";
        if constexpr (reportStats)
            std::clog << "  device-resident waves: " << waves_ << " of " << wave_ << " samples, " << (prrt_ ? mptg_prrtstar_samples_drawn(prrt_) : 0)
                      << " samples drawn, " << rewires() << " rewires, " << seconds_ * 1e3 << " ms in solve(), kernel launches: " << ctx_.launches() << "\n";
    }
    const Scenario& scenario() const { return scenario_; }
    Context& context() { return ctx_; }
};

// ------------------------------------------------------------------ device-resident PPRM
// Planner<Scenario, PPRM<device_resident, ...>>: the roadmap, its components and every stage of PPRM's addSample
// (impl/pprm/pprm.hpp:298-339) stay on the GPU (mptg_pprm_*); two words come back per wave.  solution() runs the
// reference's Dijkstra (pprm.hpp:218-246) over a host mirror fetched on demand.
// irs = true: PPRM-IRS on the device (mptg_pprm_set_spanner): the edge rows, components and solution() describe the sparse
// roadmap of the spanner; stretch weight 5 as the reference's default (impl/pprm_irs/pprm_irs.hpp:85), setStretchWeight
// before the first addStart / addGoal.
template <typename Scenario, int waveSize, int maxNodes, bool reportStats, bool irs = false>
class DevicePPRM {
    using Space = typename Scenario::Space;
    using State = typename Space::Type;
    using Distance = typename Space::Distance;
    static constexpr std::uint32_t NONE = 0xFFFFFFFFu;
    Scenario scenario_;
    Context ctx_;
    Geometry geom_;
    mptg_space_desc desc_;
    mptg_pprm* pprm_ = nullptr;
    std::uint32_t wave_ = waveSize, size_ = 0, solved_ = 0, stride_ = 0;
    impl::WaveRamp ramp_;  // wave sizes until the configured size is reached (time to the first solution)

public:
    // first wave's size (0: full-size waves from the start) and the size held until the first solution is found
    void setWaveRamp(std::uint32_t firstWave, std::uint32_t holdAt = 512) { ramp_ = impl::WaveRamp{firstWave, holdAt}; }

private:

    std::size_t starts_ = 0, goals_ = 0;
    std::uint64_t waves_ = 0;
    double seconds_ = 0;
    mutable std::vector<State> states_;  // host mirror, refreshed on demand
    mutable std::vector<std::uint32_t> edgeIdx_;
    mutable std::vector<Distance> edgeDist_;
    mutable std::vector<std::uint8_t> marks_;

    void mirror() const {
        if (states_.size() == size_) return;
        states_.resize(size_), marks_.resize(size_);
        edgeIdx_.resize((std::size_t)size_ * stride_), edgeDist_.resize((std::size_t)size_ * stride_);
        check(mptg_pprm_get_graph(pprm_, 0, size_, states_.data(), edgeIdx_.data(), edgeDist_.data(), marks_.data(), nullptr), ctx_.get(),
              "mptg_pprm_get_graph");
    }
    std::uint32_t add(const State& q, std::uint32_t marks) {
        std::uint32_t node = NONE;
        check(mptg_pprm_add_state(pprm_, q.data(), marks, &node), ctx_.get(), "mptg_pprm_add_state");
        size_ = mptg_pprm_size(pprm_);
        return node;
    }

public:
    explicit DevicePPRM(const Scenario& scenario = Scenario(), std::uint64_t seed = std::random_device{}(), int device = -1)
        : scenario_(scenario), ctx_(device), geom_(scenario_.makeGeometry(ctx_)), desc_(scenario_.space().desc()) {
        double lo[MPTG_MAX_SCALARS] = {0}, hi[MPTG_MAX_SCALARS] = {0};
        detail::fillBounds(scenario_.bounds(), 0, lo, hi);
        mptg_pprm_params prm{};
        prm.space = &desc_, prm.lo = lo, prm.hi = hi;
        State g{};
        if constexpr (impl::has_goal_fn<Scenario>::value) {  // GoalState: goal test on the device
            g = scenario_.goal().state();
            prm.goal_state = g.data(), prm.goal_radius = (double)scenario_.goal().radius();
        }
        prm.link_step = impl::linkStepOf(scenario_), prm.seed = seed, prm.capacity = (std::uint32_t)maxNodes, prm.max_wave = (std::uint32_t)waveSize;
        check(mptg_pprm_create(ctx_.get(), geom_.get(), &prm, &pprm_), ctx_.get(), "mptg_pprm_create");
        stride_ = mptg_pprm_row_stride(pprm_);
        if constexpr (irs) check(mptg_pprm_set_spanner(pprm_, 5.0, 0), ctx_.get(), "mptg_pprm_set_spanner");
    }
    // impl/pprm_irs/pprm_irs.hpp:164-166 (only with PPRMIRS<device_resident, ...>, before the first state is added)
    void setStretchWeight(Distance w) {
        static_assert(irs, "setStretchWeight belongs to PPRMIRS");
        check(mptg_pprm_set_spanner(pprm_, (double)w, 0), ctx_.get(), "mptg_pprm_set_spanner");
    }
    DevicePPRM(const DevicePPRM&) = delete;
    DevicePPRM& operator=(const DevicePPRM&) = delete;
    ~DevicePPRM() {
        if (pprm_) mptg_pprm_destroy(pprm_);
    }
    void setWaveSize(std::uint32_t w) { wave_ = w ? std::min<std::uint32_t>(w, waveSize) : 1; }

    template <typename... Args>
    void addStart(Args&&... args) {  // pprm.hpp:156-164
        if (add(State(std::forward<Args>(args)...), MPTG_PPRM_START) != NONE) ++starts_;
    }
    template <typename... Args>
    void addGoal(Args&&... args) {  // :166-169
        if (add(State(std::forward<Args>(args)...), MPTG_PPRM_GOAL) != NONE) ++goals_;
    }
    template <typename DoneFn>
    std::enable_if_t<std::is_same_v<bool, std::invoke_result_t<DoneFn>>> solve(DoneFn doneFn) {  // :171-183
        if (goals_ == 0) {
            std::mt19937_64 unused;
            addGoal(impl::sampleGoalState(scenario_, unused));
        }
        if (goals_ == 0 || starts_ == 0) throw std::runtime_error("PPRM requires both start and goal configurations");
        const auto t0 = std::chrono::steady_clock::now();
        while (!doneFn() && size_ < (std::uint32_t)maxNodes) {
            check(mptg_pprm_wave(pprm_, ramp_.next(wave_, solved_ != 0), &size_, &solved_), ctx_.get(), "mptg_pprm_wave");
            ++waves_;
        }
        seconds_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    template <typename Rep, typename Period>
    void solveFor(const std::chrono::duration<Rep, Period>& duration) {
        solveUntil(std::chrono::steady_clock::now() + duration);
    }
    template <class Clock, class Duration>
    void solveUntil(const std::chrono::time_point<Clock, Duration>& endTime) {
        solve([&] { return Clock::now() >= endTime; });
    }
    template <typename DoneFn, typename Rep, typename Period>
    void solveFor(DoneFn doneFn, const std::chrono::duration<Rep, Period>& duration) {
        const auto endTime = std::chrono::steady_clock::now() + duration;
        solve([&] { return doneFn() || std::chrono::steady_clock::now() >= endTime; });
    }
    bool solved() const { return solved_ != 0; }
    std::size_t size() const { return size_; }
    std::size_t edgeCount() const {
        mirror();
        std::size_t c = 0;
        for (std::uint32_t e : edgeIdx_) c += e != NONE;
        return c;
    }
    // shortest path over the roadmap from any start to any goal (impl/djikstras.hpp, pprm.hpp:218-246), as node indices
    std::vector<std::uint32_t> solutionNodes() const {
        mirror();
        const std::size_t n = states_.size();
        std::vector<std::vector<std::pair<std::uint32_t, Distance>>> adj(n);
        for (std::uint32_t i = 0; i < n; ++i)
            for (std::uint32_t j = 0; j < stride_; ++j) {
                const std::uint32_t nb = edgeIdx_[(std::size_t)i * stride_ + j];
                if (nb == NONE) continue;
                const Distance d = edgeDist_[(std::size_t)i * stride_ + j];
                adj[i].push_back({nb, d}), adj[nb].push_back({i, d});
            }
        std::vector<Distance> dist(n, std::numeric_limits<Distance>::infinity());
        std::vector<std::uint32_t> prev(n, NONE);
        using QE = std::pair<Distance, std::uint32_t>;
        std::priority_queue<QE, std::vector<QE>, std::greater<QE>> pq;
        for (std::uint32_t s = 0; s < n; ++s)
            if (marks_[s] & MPTG_PPRM_START) dist[s] = 0, pq.push({0, s});
        std::uint32_t hit = NONE;
        while (!pq.empty()) {
            auto [d, u] = pq.top();
            pq.pop();
            if (d > dist[u]) continue;
            if (marks_[u] & MPTG_PPRM_GOAL) {
                hit = u;
                break;
            }
            for (auto [v, w] : adj[u])
                if (d + w < dist[v]) dist[v] = d + w, prev[v] = u, pq.push({dist[v], v});
        }
        std::vector<std::uint32_t> path;
        for (std::uint32_t x = hit; x != NONE; x = prev[x]) path.push_back(x);
        std::reverse(path.begin(), path.end());
        return path;
    }
    std::vector<State> solution() const {
        std::vector<State> path;
        for (std::uint32_t n : solutionNodes()) path.push_back(states_[n]);
        return path;
    }
    template <typename Fn>
    void solution(Fn fn) const {
        impl::emitRoadmapSolution(states_, solutionNodes(), fn);
    }
    template <typename Visitor>
    void visitGraph(Visitor&& visitor) const {  // :380-387 (each edge from both of its ends)
        mirror();
        std::vector<std::vector<std::uint32_t>> back(states_.size());
        for (std::uint32_t i = 0; i < states_.size(); ++i)
            for (std::uint32_t j = 0; j < stride_; ++j)
                if (edgeIdx_[(std::size_t)i * stride_ + j] != NONE) back[edgeIdx_[(std::size_t)i * stride_ + j]].push_back(i);
        for (std::uint32_t i = 0; i < states_.size(); ++i) {
            visitor.vertex(states_[i]);
            for (std::uint32_t j = 0; j < stride_; ++j)
                if (edgeIdx_[(std::size_t)i * stride_ + j] != NONE) visitor.edge(states_[edgeIdx_[(std::size_t)i * stride_ + j]]);
            for (std::uint32_t b : back[i]) visitor.edge(states_[b]);
        }
    }
    void printStats() const {
        std::clog << "nodes in graph: " << size() << "\n";
        if constexpr (reportStats)
            std::clog << "  device-resident waves: " << waves_ << " of " << wave_ << " samples, " << mptg_pprm_samples_drawn(pprm_) << " samples drawn, "
                      << seconds_ * 1e3 << " ms in solve(), kernel launches: " << ctx_.launches() << "\n";
    }
    const Scenario& scenario() const { return scenario_; }
    Context& context() { return ctx_; }
};

// ------------------------------------------------------------------ resolvers (planner.hpp:41-47)
template <typename Scenario, typename Algorithm>
struct PlannerResolver;

template <typename Scenario, typename... Options>
struct PlannerResolver<Scenario, PRRT<Options...>> {
    using type = std::conditional_t<pack_contains_v<device_resident, Options...>,
                                    DevicePRRT<Scenario, pack_int_tag_v<wave_size, 16384, Options...>, pack_int_tag_v<max_nodes, 1 << 22, Options...>,
                                               pack_bool_tag_v<report_stats, false, Options...>>,
                                    WavePRRT<Scenario, pack_int_tag_v<wave_size, 1024, Options...>, pack_bool_tag_v<report_stats, false, Options...>>>;
};
template <typename Scenario, typename... Options>
struct PlannerResolver<Scenario, PRRTStar<Options...>> {
    static constexpr bool kNearest = pack_contains_v<rewire_k_nearest, Options...>;
    static constexpr bool rNearest = pack_contains_v<rewire_r_nearest, Options...>;
    static_assert(!(kNearest && rNearest), "RRT* tags cannot include both k_nearest and r_nearest");
    using Rewire = std::conditional_t<!rNearest, rewire_k_nearest, rewire_r_nearest>;
    using type = std::conditional_t<pack_contains_v<device_resident, Options...>,
                                    DevicePRRTStar<Scenario, pack_int_tag_v<wave_size, 16384, Options...>, pack_int_tag_v<max_nodes, 1 << 21, Options...>,
                                                   pack_bool_tag_v<report_stats, false, Options...>, rNearest>,
                                    WavePRRTStar<Scenario, pack_int_tag_v<wave_size, 1024, Options...>, Rewire, pack_bool_tag_v<report_stats, false, Options...>>>;
};
template <typename Scenario, typename... Options>
struct PlannerResolver<Scenario, PPRM<Options...>> {
    using type = std::conditional_t<pack_contains_v<device_resident, Options...>,
                                    DevicePPRM<Scenario, pack_int_tag_v<wave_size, 4096, Options...>, pack_int_tag_v<max_nodes, 1 << 20, Options...>,
                                               pack_bool_tag_v<report_stats, false, Options...>>,
                                    WavePPRM<Scenario, pack_int_tag_v<wave_size, 1024, Options...>, pack_bool_tag_v<report_stats, false, Options...>>>;
};

template <typename Scenario, typename... Options>
struct PlannerResolver<Scenario, PPRMIRS<Options...>> {  // src/mpt/pprm_irs.hpp:55-72
    static_assert(!(pack_contains_v<device_resident, Options...> && pack_bool_tag_v<keep_dense_edges, false, Options...>),
                  "keep_dense_edges is an option of the host-driven PPRMIRS; the device-resident roadmap holds the sparse edges");
    using type = std::conditional_t<pack_contains_v<device_resident, Options...>,
                                    DevicePPRM<Scenario, pack_int_tag_v<wave_size, 4096, Options...>, pack_int_tag_v<max_nodes, 1 << 20, Options...>,
                                               pack_bool_tag_v<report_stats, false, Options...>, true>,
                                    WavePPRM<Scenario, pack_int_tag_v<wave_size, 1024, Options...>, pack_bool_tag_v<report_stats, false, Options...>, true,
                                             pack_bool_tag_v<keep_dense_edges, false, Options...>>>;
};

}  // namespace impl

template <typename Scenario, typename Algorithm>
using Planner = typename impl::PlannerResolver<Scenario, Algorithm>::type;

}  // namespace mptg
