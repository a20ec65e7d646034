// nigh_kdtree.hpp -- STAND-IN (ours) for nigh::Nigh<T, Space, KeyFn, nigh::Concurrent, Strategy>: what the reference's
// planners instantiate when they run multi-threaded (src/mpt/impl/prrt/prrt.hpp:121-122, prrt_star.hpp:182-183).
// TEST / BASELINE INFRASTRUCTURE ONLY.  Nigh itself is an un-vendored, unpinned dependency that is not on this machine;
// this file gives the reference's planner classes a concurrent nearest-neighbour structure of the same kind as Nigh's
// default (an insert-only kd-tree searched without locks while other threads insert), so that their multi-threaded
// planner loop -- THEIR code -- can be timed on the host cores as the CPU baseline of bench.py.
//   L_p vector spaces: insert-only kd-tree, one point per node, axis = depth mod dimensions, children linked with a
//                      compare-and-swap; searches read the links with acquire loads and prune with the split-plane
//                      distance (a lower bound for every L_p norm).
//   other spaces:      exhaustive scan under a shared mutex.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstddef>
#include <limits>
#include <mutex>
#include <optional>
#include <queue>
#include <shared_mutex>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "nigh_linear.hpp"

namespace unc::robotics::nigh {
namespace shim_detail {
template <class Space, class = void>
struct is_lp_vector : std::false_type {};
template <class Space>
struct is_lp_vector<Space, std::void_t<decltype(Space::kDimensions)>> : std::true_type {};
}  // namespace shim_detail

template <class T, class Space, class KeyFn, class Strategy>
class Nigh<T, Space, KeyFn, Concurrent, Strategy> {
public:
    using Distance = typename Space::Distance;

private:
    using Key = typename Space::Type;
    static constexpr bool kd = shim_detail::is_lp_vector<Space>::value;
    struct Node {
        T item;
        Key key;
        std::atomic<Node*> child[2];
        Node(const T& t, const Key& k) : item(t), key(k) { child[0] = nullptr, child[1] = nullptr; }
    };
    Space space_;
    KeyFn key_;
    std::atomic<Node*> root_{nullptr};
    std::atomic<std::size_t> size_{0};
    // fallback
    mutable std::shared_mutex mutex_;
    std::vector<T> items_;

    static int dims() {
        if constexpr (kd) return Space::kDimensions;
        else return 1;
    }

    using Entry = std::pair<Distance, T>;
    struct Search {
        const Nigh& nn;
        const Key& q;
        std::size_t k;
        Distance r;  // current pruning radius
        std::priority_queue<Entry, std::vector<Entry>, std::less<Entry>> heap;  // max-heap on distance
        void visit(const Node* n, int axis) {
            if (!n) return;
            const Distance d = nn.space_.distance(n->key, q);
            if (d <= r) {
                if (heap.size() < k) heap.emplace(d, n->item);
                else if (d < heap.top().first) heap.pop(), heap.emplace(d, n->item);
                if (heap.size() == k && heap.top().first < r) r = heap.top().first;
            }
            const Distance diff = q[axis] - n->key[axis];
            const int near = diff < 0 ? 0 : 1;
            const int next = axis + 1 == dims() ? 0 : axis + 1;
            visit(n->child[near].load(std::memory_order_acquire), next);
            if ((diff < 0 ? -diff : diff) <= r) visit(n->child[near ^ 1].load(std::memory_order_acquire), next);
        }
    };

public:
    explicit Nigh(const Space& space = Space(), const KeyFn& key = KeyFn()) : space_(space), key_(key) {}
    Nigh(const Nigh&) = delete;
    ~Nigh() {
        std::vector<Node*> stack;
        if (Node* r = root_.load()) stack.push_back(r);
        while (!stack.empty()) {
            Node* n = stack.back();
            stack.pop_back();
            for (int c = 0; c < 2; ++c)
                if (Node* ch = n->child[c].load()) stack.push_back(ch);
            delete n;
        }
    }
    const Space& metricSpace() const { return space_; }
    std::size_t size() const { return size_.load(std::memory_order_relaxed); }

    void insert(const T& t) {
        if constexpr (kd) {
            Node* fresh = new Node(t, key_(t));
            std::atomic<Node*>* link = &root_;
            int axis = 0;
            for (;;) {
                Node* n = link->load(std::memory_order_acquire);
                if (!n) {
                    if (link->compare_exchange_strong(n, fresh, std::memory_order_release, std::memory_order_acquire)) break;
                }  // lost the race: n now holds the winner, descend through it
                link = &n->child[fresh->key[axis] < n->key[axis] ? 0 : 1];
                axis = axis + 1 == dims() ? 0 : axis + 1;
            }
        } else {
            std::unique_lock<std::shared_mutex> lock(mutex_);
            items_.push_back(t);
        }
        size_.fetch_add(1, std::memory_order_relaxed);
    }

    template <class Q>
    std::optional<std::pair<T, Distance>> nearest(const Q& q) const {
        std::vector<std::tuple<T, Distance>> out;
        nearest(out, q, 1);
        if (out.empty()) return std::nullopt;
        return std::make_pair(std::get<0>(out[0]), std::get<1>(out[0]));
    }

    template <class Tuple, class Q, class Alloc>
    void nearest(std::vector<Tuple, Alloc>& out, const Q& q, std::size_t k, Distance r = std::numeric_limits<Distance>::infinity()) const {
        out.clear();
        std::vector<Entry> found;
        if constexpr (kd) {
            Search s{*this, q, k, r, {}};
            s.visit(root_.load(std::memory_order_acquire), 0);
            found.reserve(s.heap.size());
            while (!s.heap.empty()) found.push_back(s.heap.top()), s.heap.pop();
            std::reverse(found.begin(), found.end());
        } else {
            std::shared_lock<std::shared_mutex> lock(mutex_);
            for (const T& t : items_) {
                const Distance d = space_.distance(key_(t), q);
                if (d <= r) found.emplace_back(d, t);
            }
            std::sort(found.begin(), found.end(), [](const Entry& a, const Entry& b) { return a.first < b.first; });
            if (found.size() > k) found.resize(k);
        }
        for (auto& [d, t] : found) {
            if constexpr (std::is_same_v<std::tuple_element_t<0, Tuple>, T>) out.emplace_back(t, d);
            else out.emplace_back(d, t);
        }
    }
};
}  // namespace unc::robotics::nigh
