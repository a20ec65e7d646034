// nigh_linear.hpp -- MINIMAL STAND-IN (ours) for nigh::Nigh<T, Space, KeyFn, Concurrency, Strategy>: an exhaustive
// scan with the (distance, insertion order) tie rule.  TEST INFRASTRUCTURE ONLY.  It lets the reference's own
// planner classes (src/mpt/impl/prrt/prrt.hpp ...) compile and run here, single-threaded, so that their loop --
// THEIR code -- can be compared with the device-resident planner on the same samples.
#pragma once
#include <algorithm>
#include <cstddef>
#include <limits>
#include <optional>
#include <tuple>
#include <utility>
#include <vector>

#include "nigh_shim.hpp"

namespace unc::robotics::nigh {
template <unsigned degree = 0, unsigned minDegree = 0, unsigned maxDegree = 0, unsigned maxNumPtsPerLeaf = 0, unsigned removedCacheSize = 0,
          bool rebalancing = false>
struct GNAT {};
template <class Space, class Concurrency>
using auto_strategy_t = Linear;
using metric::cartesian_state_element;
using metric::cartesian_state_element_t;

template <class T, class Space, class KeyFn, class Concurrency = NoThreadSafety, class Strategy = Linear>
class Nigh {
    Space space_;
    KeyFn key_;
    std::vector<T> items_;

public:
    using Distance = typename Space::Distance;
    explicit Nigh(const Space& space = Space(), const KeyFn& key = KeyFn()) : space_(space), key_(key) {}
    const Space& metricSpace() const { return space_; }
    std::size_t size() const { return items_.size(); }
    void insert(const T& t) { items_.push_back(t); }
    template <class Key>
    std::optional<std::pair<T, Distance>> nearest(const Key& q) const {
        if (items_.empty()) return std::nullopt;
        std::size_t best = 0;
        Distance bd = std::numeric_limits<Distance>::infinity();
        for (std::size_t i = 0; i < items_.size(); ++i) {
            const Distance d = space_.distance(key_(items_[i]), q);
            if (d < bd) bd = d, best = i;  // strict: the first inserted wins ties
        }
        return std::make_pair(items_[best], bd);
    }
    // k nearest within r, ascending by (distance, insertion order); result tuples are (T, Distance) or (Distance, T)
    template <class Tuple, class Key, class Alloc>
    void nearest(std::vector<Tuple, Alloc>& out, const Key& q, std::size_t k, Distance r = std::numeric_limits<Distance>::infinity()) const {
        std::vector<std::pair<Distance, std::size_t>> all;
        for (std::size_t i = 0; i < items_.size(); ++i) {
            const Distance d = space_.distance(key_(items_[i]), q);
            if (d <= r) all.emplace_back(d, i);
        }
        std::sort(all.begin(), all.end());
        if (all.size() > k) all.resize(k);
        out.clear();
        for (auto& [d, i] : all) {
            if constexpr (std::is_same_v<std::tuple_element_t<0, Tuple>, T>) out.emplace_back(items_[i], d);
            else out.emplace_back(d, items_[i]);
        }
    }
    template <class Fn>
    void visit(Fn&& fn) const {
        for (const T& t : items_) fn(t);
    }
};
}  // namespace unc::robotics::nigh
