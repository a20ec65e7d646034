// spaces.hpp -- host-side state / space / bounds / goal / sampler types with the reference's names,
// laid out so that a std::vector<State> IS the array-of-states buffer the C ABI takes.
//
// Reference counterparts (paths relative to the reference root):
//   LPSpace / L1Space / L2Space     src/mpt/lp_space.hpp, src/mpt/impl/metrics.hpp
//   SO2Space, SO3Space              src/mpt/so2_space.hpp, src/mpt/so3_space.hpp
//   SE3State / SE3Space<S,so3,l2>   src/mpt/se3_space.hpp:53-113 (rotation first)
//   SE2State / SE2Space<S,so2,l2>   src/mpt/se2_space.hpp:46-85 (translation first)
//   BoxBounds, Unbounded            src/mpt/box_bounds.hpp, src/mpt/unbounded.hpp
//   GoalState                       src/mpt/goal_state.hpp:44-89
//   UniformSampler                  src/mpt/uniform_box_sampler.hpp:60-68, impl/uniform_sampler_so3.hpp:55-68,
//                                   impl/uniform_sampler_cartesian.hpp:75-78
// The reference stores states in Eigen types; here a state is a flat array of scalars in C-ABI
// order (see INTEGRATION.md for the codec a maintainer would write for Eigen-typed states).
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <initializer_list>
#include <limits>
#include <ostream>
#include <random>
#include <utility>

#include "mptg.h"
#include "mptg_space.h"

namespace mptg {

template <typename S, int N>
struct State {
    S v[N];
    static constexpr int scalars = N;
    S& operator[](int i) { return v[i]; }
    const S& operator[](int i) const { return v[i]; }
    const S* data() const { return v; }
    S* data() { return v; }
    void fill(S x) {
        for (int i = 0; i < N; ++i) v[i] = x;
    }
    static State Zero() {
        State s;
        s.fill(S(0));
        return s;
    }
    static State Constant(S x) {
        State s;
        s.fill(x);
        return s;
    }
    State operator-() const {
        State s;
        for (int i = 0; i < N; ++i) s.v[i] = -v[i];
        return s;
    }
    bool operator==(const State& o) const {
        for (int i = 0; i < N; ++i)
            if (!(v[i] == o.v[i])) return false;
        return true;
    }
    bool operator!=(const State& o) const { return !(*this == o); }
    template <typename Char, typename Traits>
    friend std::basic_ostream<Char, Traits>& operator<<(std::basic_ostream<Char, Traits>& out, const State& s) {
        for (int i = 0; i < N; ++i) out << (i ? " " : "") << s.v[i];
        return out;
    }
};

template <typename S, int N>
State<S, N> makeState(std::initializer_list<S> il) {
    State<S, N> s = State<S, N>::Zero();
    int i = 0;
    for (S x : il)
        if (i < N) s.v[i++] = x;
    return s;
}

// SE3State: rotation() = quaternion coeffs (x,y,z,w) in v[0..3], translation() in v[4..6]
template <typename S>
struct SE3State : State<S, 7> {
    SE3State() = default;
    SE3State(const State<S, 7>& s) : State<S, 7>(s) {}
    S* rotation() { return this->v; }
    const S* rotation() const { return this->v; }
    S* translation() { return this->v + 4; }
    const S* translation() const { return this->v + 4; }
    static SE3State identityAt(S x, S y, S z) {
        SE3State q;
        q.v[0] = q.v[1] = q.v[2] = S(0);
        q.v[3] = S(1);
        q.v[4] = x, q.v[5] = y, q.v[6] = z;
        return q;
    }
};

namespace detail {
template <typename S>
constexpr int scalarTag() {
    return sizeof(S) == 4 ? MPTG_F32 : MPTG_F64;
}
template <typename S, typename State_>
S hostDistance(const mptg_space_desc& d, const State_& a, const State_& b) {
    const DevSpace<S> sp = makeDevSpace<S>(d);
    return dev::distance<S>(sp, [&](int c) { return a.v[c]; }, [&](int c) { return b.v[c]; });
}
}  // namespace detail

// Base of every space: one mptg_space_desc built from compile-time parts.
template <typename S, typename StateT, typename Derived>
struct SpaceBase {
    using Scalar = S;
    using Distance = S;
    using Type = StateT;
    static constexpr int scalars = StateT::scalars;
    mptg_space_desc desc() const { return static_cast<const Derived*>(this)->makeDesc(); }
    unsigned dimensions() const { return (unsigned)mptg_space_dimensions_of(desc()); }
    Distance distance(const Type& a, const Type& b) const { return detail::hostDistance<S>(desc(), a, b); }

private:
    static int mptg_space_dimensions_of(const mptg_space_desc& d) {
        int n = 0;
        for (int i = 0; i < d.n_parts; ++i) n += d.part[i].kind == MPTG_PART_SO3 ? 3 : d.part[i].dim;
        return n;
    }
};

template <typename S, int N, int P>
struct LPSpace : SpaceBase<S, State<S, N>, LPSpace<S, N, P>> {
    mptg_space_desc makeDesc() const {
        mptg_space_desc d{};
        d.n_parts = 1;
        d.scalar = detail::scalarTag<S>();
        d.part[0] = {MPTG_PART_LP, P, N, 0, 1.0};
        return d;
    }
};
template <typename S, int N>
using L2Space = LPSpace<S, N, 2>;
template <typename S, int N>
using L1Space = LPSpace<S, N, 1>;
template <typename S, int N>
using LInfSpace = LPSpace<S, N, 0>;

template <typename S, int N = 1, int P = 1>
struct SO2Space : SpaceBase<S, State<S, N>, SO2Space<S, N, P>> {
    mptg_space_desc makeDesc() const {
        mptg_space_desc d{};
        d.n_parts = 1;
        d.scalar = detail::scalarTag<S>();
        d.part[0] = {MPTG_PART_SO2, P, N, 0, 1.0};
        return d;
    }
};

template <typename S>
struct SO3Space : SpaceBase<S, State<S, 4>, SO3Space<S>> {
    mptg_space_desc makeDesc() const {
        mptg_space_desc d{};
        d.n_parts = 1;
        d.scalar = detail::scalarTag<S>();
        d.part[0] = {MPTG_PART_SO3, 0, 4, 0, 1.0};
        return d;
    }
};

template <typename S, std::intmax_t so3wt = 1, std::intmax_t l2wt = 1>
struct SE3Space : SpaceBase<S, SE3State<S>, SE3Space<S, so3wt, l2wt>> {
    mptg_space_desc makeDesc() const {
        mptg_space_desc d{};
        d.n_parts = 2;
        d.scalar = detail::scalarTag<S>();
        d.part[0] = {MPTG_PART_SO3, 0, 4, 0, (double)so3wt};
        d.part[1] = {MPTG_PART_LP, 2, 3, 0, (double)l2wt};
        return d;
    }
};

template <typename S, std::intmax_t so2wt = 1, std::intmax_t l2wt = 1>
struct SE2Space : SpaceBase<S, State<S, 3>, SE2Space<S, so2wt, l2wt>> {
    mptg_space_desc makeDesc() const {
        mptg_space_desc d{};
        d.n_parts = 2;
        d.scalar = detail::scalarTag<S>();
        d.part[0] = {MPTG_PART_LP, 2, 2, 0, (double)l2wt};
        d.part[1] = {MPTG_PART_SO2, 1, 1, 0, (double)so2wt};
        return d;
    }
};

// interpolate(space, a, b, t): same overload set as the reference's free functions
template <typename Space>
typename Space::Type interpolate(const Space& space, const typename Space::Type& a, const typename Space::Type& b,
                                 typename Space::Distance t) {
    using S = typename Space::Scalar;
    typename Space::Type q;
    const DevSpace<S> sp = makeDevSpace<S>(space.desc());
    dev::interpolate<S>(sp, a.v, b.v, t, q.v);
    return q;
}

// ------------------------------------------------------------------ bounds
struct Unbounded {};

template <typename S, int N>
struct BoxBounds {
    State<S, N> min_, max_;
    BoxBounds() = default;
    BoxBounds(const State<S, N>& mn, const State<S, N>& mx) : min_(mn), max_(mx) {}
    const State<S, N>& min() const { return min_; }
    const State<S, N>& max() const { return max_; }
    S measure() const {
        S m = S(1);
        for (int i = 0; i < N; ++i) m *= max_[i] - min_[i];
        return m;
    }
};

// SE(3) bounds of the reference: std::tuple<Unbounded, BoxBounds<S,3>> (se3_rigid_body_scenario.hpp:245)
template <typename S>
struct SE3Bounds {
    BoxBounds<S, 3> translation;
    SE3Bounds() = default;
    explicit SE3Bounds(const BoxBounds<S, 3>& t) : translation(t) {}
    S measure() const { return translation.measure() * S(9.869604401089358); }  // volume of SO(3) ~ pi^2 (S^3/2, half angle metric)
};

// ------------------------------------------------------------------ goal
template <typename Space>
class GoalState {
    using State_ = typename Space::Type;
    using Distance = typename Space::Distance;
    Distance radius_;
    State_ goal_;

public:
    GoalState(Distance radius, const State_& goal) : radius_(radius), goal_(goal) {}
    const State_& state() const { return goal_; }
    Distance radius() const { return radius_; }
    // goal_state.hpp:64-69: (true, 0) inside the radius, otherwise (false, distance - radius)
    std::pair<bool, Distance> operator()(const Space& space, const State_& q) const {
        const Distance d = space.distance(q, goal_);
        return d <= radius_ ? std::make_pair(true, Distance(0)) : std::make_pair(false, d - radius_);
    }
};

// ------------------------------------------------------------------ samplers
template <typename Space, typename Bounds>
struct UniformSampler;

// uniform_box_sampler.hpp:60-68: one uniform_real_distribution(min,max) per coordinate, in index order
template <typename S, int N, int P>
struct UniformSampler<LPSpace<S, N, P>, BoxBounds<S, N>> {
    BoxBounds<S, N> bounds;
    UniformSampler(const LPSpace<S, N, P>&, const BoxBounds<S, N>& b) : bounds(b) {}
    template <typename RNG>
    State<S, N> operator()(RNG& rng) const {
        State<S, N> q;
        for (int i = 0; i < N; ++i) {
            std::uniform_real_distribution<S> dist(bounds.min()[i], bounds.max()[i]);
            q[i] = dist(rng);
        }
        return q;
    }
    S measure() const { return bounds.measure(); }
};

// impl/uniform_sampler_so3.hpp:55-68
template <typename S, typename RNG>
inline void sampleSO3(RNG& rng, S* xyzw) {
    std::uniform_real_distribution<S> dist01(0, 1);
    std::uniform_real_distribution<S> dist2pi(0, S(2) * fp::consts<S>::pi());
    const S a = dist01(rng);
    const S b = dist2pi(rng);
    const S c = dist2pi(rng);
    const S w = std::sqrt(1 - a) * std::sin(b);
    const S x = std::sqrt(1 - a) * std::cos(b);
    const S y = std::sqrt(a) * std::sin(c);
    const S z = std::sqrt(a) * std::cos(c);
    xyzw[0] = x, xyzw[1] = y, xyzw[2] = z, xyzw[3] = w;
}

// impl/uniform_sampler_cartesian.hpp:75-78: components in tuple order -> rotation first, then the box
template <typename S, std::intmax_t so3wt, std::intmax_t l2wt>
struct UniformSampler<SE3Space<S, so3wt, l2wt>, SE3Bounds<S>> {
    SE3Bounds<S> bounds;
    UniformSampler(const SE3Space<S, so3wt, l2wt>&, const SE3Bounds<S>& b) : bounds(b) {}
    template <typename RNG>
    SE3State<S> operator()(RNG& rng) const {
        SE3State<S> q;
        sampleSO3<S>(rng, q.rotation());
        for (int i = 0; i < 3; ++i) {
            std::uniform_real_distribution<S> dist(bounds.translation.min()[i], bounds.translation.max()[i]);
            q.translation()[i] = dist(rng);
        }
        return q;
    }
    // Product of the parts' measures, each Scaled part times its weight (impl/uniform_sampler_scaled.hpp:59-66,
    // impl/uniform_sampler_cartesian.hpp:93-97; SO(3): pi^2, impl/uniform_sampler_so3.hpp:79-83): pi^2 so3wt * volume l2wt.
    // (r1 returned pi^2 * volume: with so3wt = 50 that made the r-nearest rewire radius 50^(-1/6) = 0.52x the reference's.)
    S measure() const { return bounds.measure() * S(so3wt) * S(l2wt); }
};

}  // namespace mptg
