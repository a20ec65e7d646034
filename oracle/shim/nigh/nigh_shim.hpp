// nigh_shim.hpp -- MINIMAL STAND-IN (ours) for the parts of UNC-Robotics/nigh that the reference's
// space headers name.  TEST INFRASTRUCTURE ONLY.  Nigh is an un-vendored dependency of the reference
// (test/CMakeLists.txt:29) and is not on this machine, so the metric arithmetic below is OURS, written
// to satisfy the reference's own known-answer tests (test/{lp,so2,so3,scaled,se2,se3}_space_test.cpp).
// What this buys: the reference's interpolate() overloads, DiscreteMotionValidator and demo scenario
// checks -- THEIR code, compiled from where it lies under /root/reference -- run on our inputs.
#pragma once
#include <Eigen/Dense>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <ratio>
#include <tuple>
#include <utility>

namespace unc::robotics::nigh {
struct Linear {};
template <std::size_t B = 8> struct KDTreeBatch {};
struct NoThreadSafety {};
struct Concurrent {};

namespace impl {
template <class T> constexpr T PI = T(3.14159265358979323846264338327950288419716939937510582097494459230781640628620L);
namespace so2 {
// counter-clockwise distance in [0, 2pi) and wrap to [-pi, pi] by repeated +-2pi
template <class S> S ccwDist(S a, S b) { S d = b - a; if (d < S(0)) d = d + S(2) * PI<S>; return d; }
template <class S> S bound(S x) { while (x > PI<S>) x = x - S(2) * PI<S>; while (x < -PI<S>) x = x + S(2) * PI<S>; return x; }
}  // namespace so2
}  // namespace impl

namespace metric {
template <int p> struct LP {};
using L1 = LP<1>; using L2 = LP<2>; using LInf = LP<-1>;
template <int p = 1> struct SO2 {};
struct SO3 {};
template <class M, class W> struct Scaled {};
template <class... M> struct Cartesian {};

template <class T, class M> struct Space;

// ---- LP over Eigen column vectors
template <class S, int N, int p>
struct Space<Eigen::Matrix<S, N, 1>, LP<p>> {
    using Type = Eigen::Matrix<S, N, 1>; using Distance = S; using Metric = LP<p>;
    static constexpr int kDimensions = N;
    constexpr unsigned dimensions() const { return N; }
    static S& coeff(Type& q, std::size_t i) { return q[(int)i]; }
    static const S& coeff(const Type& q, std::size_t i) { return q[(int)i]; }
    Distance distance(const Type& a, const Type& b) const {
        if (p == 2) { S s = 0; for (int i = 0; i < N; ++i) { S d = a[i] - b[i]; s = (i == 0) ? d * d : std::fma(d, d, s); } return std::sqrt(s); }
        if (p == 1) { S s = 0; for (int i = 0; i < N; ++i) s = (i == 0) ? std::abs(a[i] - b[i]) : s + std::abs(a[i] - b[i]); return s; }
        S s = 0; for (int i = 0; i < N; ++i) s = std::abs(a[i] - b[i]) > s ? std::abs(a[i] - b[i]) : s; return s;
    }
};
// ---- SO2 over a scalar
template <class S, int p>
struct Space<S, SO2<p>> {
    using Type = S; using Distance = S;
    constexpr unsigned dimensions() const { return 1; }
    static S& coeff(S& q, std::size_t) { return q; }
    static const S& coeff(const S& q, std::size_t) { return q; }
    Distance distance(const S& a, const S& b) const { S d = std::abs(a - b); if (d > impl::PI<S>) d = S(2) * impl::PI<S> - d; return d; }
};
// ---- SO3 over Eigen quaternions (distance = acos|a.b|, half the rotation angle)
template <class S>
struct Space<Eigen::Quaternion<S>, SO3> {
    using Type = Eigen::Quaternion<S>; using Distance = S;
    constexpr unsigned dimensions() const { return 3; }
    static S& coeff(Type& q, std::size_t i) { return q.coeffs()[(int)i]; }
    static const S& coeff(const Type& q, std::size_t i) { return q.coeffs()[(int)i]; }
    Distance distance(const Type& a, const Type& b) const {
        S d = std::abs(a.coeffs().dot(b.coeffs())); return std::acos(d > S(1) ? S(1) : d);
    }
};
// ---- Scaled
template <class T, class M, std::intmax_t num, std::intmax_t den>
struct Space<T, Scaled<M, std::ratio<num, den>>> {
    using Type = T; using Distance = typename Space<T, M>::Distance;
    Space<T, M> inner_;
    const Space<T, M>& space() const { return inner_; }
    constexpr unsigned dimensions() const { return inner_.dimensions(); }
    Distance distance(const T& a, const T& b) const { return inner_.distance(a, b) * num / den; }
};
// ---- Cartesian over tuple-like states
template <std::size_t I, class T> struct cartesian_state_element {
    using type = std::tuple_element_t<I, T>;
    static type& get(T& q) { return std::get<I>(q); }
    static const type& get(const T& q) { return std::get<I>(q); }
};
template <std::size_t I, class T> using cartesian_state_element_t = typename cartesian_state_element<I, T>::type;

template <class T, class... M>
struct Space<T, Cartesian<M...>> {
    using Type = T;
    template <std::size_t... I> static auto subspaces(std::index_sequence<I...>) -> std::tuple<Space<std::tuple_element_t<I, T>, M>...>;
    using Tuple = decltype(subspaces(std::index_sequence_for<M...>{}));
    Tuple spaces_;
    using Distance = typename std::tuple_element_t<0, Tuple>::Distance;
    template <std::size_t... I> unsigned dims(std::index_sequence<I...>) const { return (std::get<I>(spaces_).dimensions() + ...); }
    unsigned dimensions() const { return dims(std::index_sequence_for<M...>{}); }
    template <std::size_t... I> Distance dist(const T& a, const T& b, std::index_sequence<I...>) const {
        Distance parts[] = {std::get<I>(spaces_).distance(cartesian_state_element<I, T>::get(a), cartesian_state_element<I, T>::get(b))...};
        Distance s = parts[0]; for (std::size_t i = 1; i < sizeof...(M); ++i) s = s + parts[i]; return s;
    }
    Distance distance(const T& a, const T& b) const { return dist(a, b, std::index_sequence_for<M...>{}); }
};

template <class S, int N> using L2Space = Space<Eigen::Matrix<S, N, 1>, L2>;
template <class S, int N> using L1Space = Space<Eigen::Matrix<S, N, 1>, L1>;
template <class S, int N, int p> using LPSpace = Space<Eigen::Matrix<S, N, 1>, LP<p>>;
template <class S> using SO2Space = Space<S, SO2<1>>;
template <class S> using SO3Space = Space<Eigen::Quaternion<S>, SO3>;
template <class Sp, class W> using ScaledSpace = Space<typename Sp::Type, Scaled<typename Sp::Metric, W>>;
}  // namespace metric
}  // namespace unc::robotics::nigh

namespace std {
template <std::size_t I, class T, class... M>
const auto& get(const unc::robotics::nigh::metric::Space<T, unc::robotics::nigh::metric::Cartesian<M...>>& s) { return std::get<I>(s.spaces_); }
}
