// nigh_binding.hpp -- THE REFERENCE-SIDE BINDING: a nearest-neighbour strategy tag for UNC-Robotics/mpt's own planners.
//
// What a maintainer of the reference adds to run its UNMODIFIED planner classes (src/mpt/impl/{prrt,prrt_star,pprm,
// pprm_irs}) against libmptg's nearest-neighbour structure:
//
//     #include <mpt/prrt_star.hpp>          // the reference, with Nigh on the include path
//     #include <mptg/nigh_binding.hpp>      // this file (after the reference's headers)
//     using Algorithm = unc::robotics::mpt::PRRTStar<mptg::GpuBatch>;   // the strategy tag, like nigh::KDTreeBatch<8>
//     unc::robotics::mpt::Planner<Scenario, Algorithm> planner(scenario);
//
// Two pieces, both in the reference's own extension points:
//   * impl::pack_nearest<mptg::GpuBatch, Rest...>   -- the option parser recognises the tag (src/mpt/impl/pack_nearest.hpp:54-83)
//   * nigh::Nigh<T, Space, KeyFn, Concurrency, mptg::GpuBatch> -- a specialisation with the members the planners call on nn_
//     (prrt.hpp:122,178,186,408,447; prrt_star.hpp:183,262,278,507,559-562,619; pprm.hpp:81,153,302-304,337;
//     rrg_rewire_neighbors.hpp:65-67,125-128): size(), insert(node), nearest(q) -> optional<pair<T, Distance>>,
//     nearest(nbh, q, k [, r]) filling (T, Distance) or (Distance, T) tuples in ascending (distance, insertion) order.
// The space descriptor and the state codec are derived from Nigh's metric tags (metric::LP / SO2 / SO3 / Scaled /
// Cartesian over Eigen vectors, scalars, quaternions and tuple-like states; SE(3): rotation then translation).
//
// One query per call is the reference's calling convention, not a way to use a GPU: every call here is a synchronous
// round trip.  It exists so that the reference's planners run and produce the same graphs with the strategy switched
// (tests/cpp/reference_planner_parity.cpp does exactly that); the batched wave planners of planner.hpp are the fast path.
#pragma once

#include <cstddef>
#include <cstdint>
#include <limits>
#include <mutex>
#include <optional>
#include <ratio>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "mptg.h"
#include "tags.hpp"

namespace mptg::nighbind {
namespace nm = unc::robotics::nigh::metric;

inline void fail(mptg_ctx* ctx, const char* what) { throw std::runtime_error(std::string(what) + ": " + mptg_last_error(ctx)); }

// ---- Nigh metric space -> mptg_space_desc parts (weights multiply through Scaled, parts concatenate through Cartesian)
inline void addPart(mptg_space_desc& d, int kind, int p, int dim, double w) {
    if (d.n_parts >= MPTG_MAX_PARTS) throw std::invalid_argument("mptg: more metric parts than MPTG_MAX_PARTS");
    mptg_space_part& part = d.part[d.n_parts++];
    part.kind = kind, part.p = p, part.dim = dim, part.weight = w;
}
template <class S, int N, int p>
void describe(const nm::Space<Eigen::Matrix<S, N, 1>, nm::LP<p>>*, double w, mptg_space_desc& d) {
    static_assert(p == 1 || p == 2 || p == -1, "L1, L2 and L-infinity are carried");
    addPart(d, MPTG_PART_LP, p == -1 ? 0 : p, N, w);
}
template <class S, int p>
void describe(const nm::Space<S, nm::SO2<p>>*, double w, mptg_space_desc& d) {
    addPart(d, MPTG_PART_SO2, p, 1, w);
}
template <class S, int N, int p>
void describe(const nm::Space<Eigen::Matrix<S, N, 1>, nm::SO2<p>>*, double w, mptg_space_desc& d) {
    addPart(d, MPTG_PART_SO2, p, N, w);
}
template <class S>
void describe(const nm::Space<Eigen::Quaternion<S>, nm::SO3>*, double w, mptg_space_desc& d) {
    addPart(d, MPTG_PART_SO3, 0, 4, w);
}
template <class T, class M, std::intmax_t num, std::intmax_t den>
void describe(const nm::Space<T, nm::Scaled<M, std::ratio<num, den>>>*, double w, mptg_space_desc& d) {
    describe(static_cast<const nm::Space<T, M>*>(nullptr), w * double(num) / double(den), d);
}
template <class T, class... M, std::size_t... I>
void describeTuple(double w, mptg_space_desc& d, std::index_sequence<I...>) {
    (describe(static_cast<const nm::Space<nm::cartesian_state_element_t<I, T>, M>*>(nullptr), w, d), ...);
}
template <class T, class... M>
void describe(const nm::Space<T, nm::Cartesian<M...>>*, double w, mptg_space_desc& d) {
    describeTuple<T, M...>(w, d, std::index_sequence_for<M...>{});
}

// ---- state codec, driven by the space type: the scalars of a state in the order of its parts (quaternions as x, y, z, w)
template <class S, int N, class M>
S* pack(const nm::Space<Eigen::Matrix<S, N, 1>, M>*, const Eigen::Matrix<S, N, 1>& q, S* out) {
    for (int i = 0; i < N; ++i) *out++ = q[i];
    return out;
}
template <class S, int p>
S* pack(const nm::Space<S, nm::SO2<p>>*, const S& q, S* out) {
    *out++ = q;
    return out;
}
template <class S>
S* pack(const nm::Space<Eigen::Quaternion<S>, nm::SO3>*, const Eigen::Quaternion<S>& q, typename nm::Space<Eigen::Quaternion<S>, nm::SO3>::Distance* out) {
    for (int i = 0; i < 4; ++i) *out++ = q.coeffs()[i];
    return out;
}
template <class T, class M, std::intmax_t num, std::intmax_t den, class S>
S* pack(const nm::Space<T, nm::Scaled<M, std::ratio<num, den>>>*, const T& q, S* out) {
    return pack(static_cast<const nm::Space<T, M>*>(nullptr), q, out);
}
template <class T, class S, class... M, std::size_t... I>
S* packTuple(const T& q, S* out, std::index_sequence<I...>) {
    ((out = pack(static_cast<const nm::Space<nm::cartesian_state_element_t<I, T>, M>*>(nullptr), nm::cartesian_state_element<I, T>::get(q), out)), ...);
    return out;
}
template <class T, class... M, class S>
S* pack(const nm::Space<T, nm::Cartesian<M...>>*, const T& q, S* out) {
    return packTuple<T, S, M...>(q, out, std::index_sequence_for<M...>{});
}
}  // namespace mptg::nighbind

// 1. the option parser recognises the tag (mirrors src/mpt/impl/pack_nearest.hpp:54-83)
namespace unc::robotics::mpt::impl {
template <typename... Rest>
struct pack_nearest<mptg::GpuBatch, Rest...> {
    using type = mptg::GpuBatch;
    static_assert(std::is_void_v<pack_nearest_t<Rest...>>, "multiple nearest neighbor strategies");
};
}  // namespace unc::robotics::mpt::impl

// 2. the structure the planners hold as nn_
namespace unc::robotics::nigh {
template <class T, class Space, class KeyFn, class Concurrency>
class Nigh<T, Space, KeyFn, Concurrency, mptg::GpuBatch> {
public:
    using Distance = typename Space::Distance;

private:
    static_assert(std::is_same_v<Distance, float> || std::is_same_v<Distance, double>, "float or double spaces");
    Space space_;
    KeyFn key_;
    mptg_space_desc desc_{};
    int scalars_ = 0;
    mptg_ctx* ctx_ = nullptr;
    mptg_knn* knn_ = nullptr;
    std::uint32_t capacity_ = 1u << 14;
    std::vector<T> handles_;        // node handle of every inserted point, index = the device's point index
    std::vector<Distance> packed_;  // host copy of the inserted states, for growing the device structure
    mutable std::mutex mutex_;      // the context is single-owner: concurrent workers take turns (Concurrency = Concurrent)
    mutable std::vector<Distance> q_, dist_;
    mutable std::vector<std::uint32_t> idx_;

    void create() {
        if (mptg_knn_create(ctx_, &desc_, capacity_, &knn_) != MPTG_OK) mptg::nighbind::fail(ctx_, "mptg_knn_create");
        if (!handles_.empty() && mptg_knn_insert(knn_, packed_.data(), (std::uint32_t)handles_.size(), nullptr) != MPTG_OK)
            mptg::nighbind::fail(ctx_, "mptg_knn_insert");
    }
    template <class Key>
    void query(const Key& q, std::uint32_t k, double r, std::uint32_t& count) const {
        q_.resize((std::size_t)scalars_);
        mptg::nighbind::pack(static_cast<const Space*>(nullptr), q, q_.data());
        idx_.resize(k), dist_.resize(k);
        if (mptg_knn_query(knn_, q_.data(), 1, k, r, idx_.data(), dist_.data(), &count) != MPTG_OK) mptg::nighbind::fail(ctx_, "mptg_knn_query");
    }

public:
    explicit Nigh(const Space& space = Space(), const KeyFn& key = KeyFn()) : space_(space), key_(key) {
        desc_.scalar = sizeof(Distance) == 4 ? MPTG_F32 : MPTG_F64;
        mptg::nighbind::describe(static_cast<const Space*>(nullptr), 1.0, desc_);
        scalars_ = mptg_space_scalars(&desc_);
        if (scalars_ <= 0) throw std::invalid_argument("mptg: this metric space is not carried by the C ABI");
        if (mptg_ctx_create(-1, &ctx_) != MPTG_OK) throw std::runtime_error(std::string("mptg_ctx_create: ") + mptg_last_error(nullptr));
        create();
    }
    Nigh(const Nigh&) = delete;
    Nigh& operator=(const Nigh&) = delete;
    ~Nigh() {
        if (knn_) mptg_knn_destroy(knn_);
        if (ctx_) mptg_ctx_destroy(ctx_);
    }
    const Space& metricSpace() const { return space_; }
    std::size_t size() const {
        std::lock_guard<std::mutex> lock(mutex_);
        return handles_.size();
    }
    void insert(const T& t) {
        std::lock_guard<std::mutex> lock(mutex_);
        const std::size_t at = packed_.size();
        packed_.resize(at + (std::size_t)scalars_);
        mptg::nighbind::pack(static_cast<const Space*>(nullptr), key_(t), packed_.data() + at);
        handles_.push_back(t);
        if (handles_.size() > capacity_) {  // grow: a new structure of twice the size, every point again
            mptg_knn_destroy(knn_);
            knn_ = nullptr;
            capacity_ *= 2;
            create();
        } else if (mptg_knn_insert(knn_, packed_.data() + at, 1, nullptr) != MPTG_OK) {
            mptg::nighbind::fail(ctx_, "mptg_knn_insert");
        }
    }
    template <class Key>
    std::optional<std::pair<T, Distance>> nearest(const Key& q) const {
        std::lock_guard<std::mutex> lock(mutex_);
        if (handles_.empty()) return std::nullopt;
        std::uint32_t count = 0;
        query(q, 1, -1.0, count);
        if (count == 0) return std::nullopt;  // a NaN query
        return std::make_pair(handles_[idx_[0]], dist_[0]);
    }
    // k nearest within r, ascending; k beyond MPTG_MAX_K (the radius form passes SIZE_MAX) is answered with the
    // MPTG_MAX_K nearest inside the ball
    template <class Tuple, class Key, class Alloc>
    void nearest(std::vector<Tuple, Alloc>& out, const Key& q, std::size_t k, Distance r = std::numeric_limits<Distance>::infinity()) const {
        std::lock_guard<std::mutex> lock(mutex_);
        out.clear();
        if (handles_.empty() || k == 0) return;
        std::uint32_t count = 0;
        query(q, (std::uint32_t)(k < (std::size_t)MPTG_MAX_K ? k : (std::size_t)MPTG_MAX_K), r < std::numeric_limits<Distance>::infinity() ? (double)r : -1.0, count);
        for (std::uint32_t i = 0; i < count; ++i) {
            if constexpr (std::is_same_v<std::tuple_element_t<0, Tuple>, T>) out.emplace_back(handles_[idx_[i]], dist_[i]);
            else out.emplace_back(dist_[i], handles_[idx_[i]]);
        }
    }
    template <class Fn>
    void visit(Fn&& fn) const {
        for (const T& t : handles_) fn(t);
    }
};
}  // namespace unc::robotics::nigh
