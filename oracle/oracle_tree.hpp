// oracle_tree.hpp -- exact tree-accelerated kNN on the CPU.  TEST INFRASTRUCTURE / CPU BASELINE ONLY.
//
// Role: stands in for the reference's default nearest-neighbour structure (nigh::KDTreeBatch<8>,
// demo/se3_rigid_body_scenario.hpp:241, impl/nearest_strategy.hpp:56-58) when bench.py times the CPU
// path on large trees.  Nigh itself is not in /root/reference, so this is a
// port of the idea -- a median-split bounding-box tree with leaves of <= 8 points searched depth
// first, nearer child first -- not of its code.  Results are identical to oracle::knnBrute
// (same distance(), same (distance,index) order); tests/test_oracle.py checks that.
#pragma once

#include <queue>

#include "oracle.hpp"

namespace oracle {

template <typename S>
struct BoxTree {
    static constexpr int LEAF = 8;
    mptg_space_desc sp{};
    int D = 0;
    std::vector<S> pts;         // canonicalised copies (SO3 part: w >= 0), tree order
    std::vector<uint32_t> ids;  // original index per tree slot
    struct Node {
        int left = -1, right = -1;
        uint32_t begin = 0, end = 0;
    };
    std::vector<Node> nodes;
    std::vector<S> lo, hi;  // per node, D scalars each

    void build(const mptg_space_desc& space, const S* p, uint32_t n) {
        sp = space;
        D = spaceScalars(sp);
        pts.assign(p, p + (size_t)n * D);
        // q and -q are the same rotation and distance() only sees |a.b| (exact under negation)
        int off = 0;
        for (int i = 0; i < sp.n_parts; ++i) {
            if (sp.part[i].kind == MPTG_PART_SO3)
                for (uint32_t j = 0; j < n; ++j) {
                    S* q = &pts[(size_t)j * D + off];
                    if (q[3] < 0) q[0] = -q[0], q[1] = -q[1], q[2] = -q[2], q[3] = -q[3];
                }
            off += partScalars(sp.part[i]);
        }
        ids.resize(n);
        for (uint32_t i = 0; i < n; ++i) ids[i] = i;
        std::vector<uint32_t> order(ids);
        nodes.clear();
        lo.clear();
        hi.clear();
        if (n) buildRec(order, 0, n);
        // permute points into tree order
        std::vector<S> sorted((size_t)n * D);
        for (uint32_t i = 0; i < n; ++i) std::copy(&pts[(size_t)order[i] * D], &pts[(size_t)order[i] * D] + D, &sorted[(size_t)i * D]);
        pts.swap(sorted);
        ids = order;
    }

    std::vector<S> dimWeight() const {
        std::vector<S> w(D, S(1));
        int off = 0;
        for (int i = 0; i < sp.n_parts; ++i) {
            for (int j = 0; j < partScalars(sp.part[i]); ++j) w[off + j] = S(sp.part[i].weight);
            off += partScalars(sp.part[i]);
        }
        return w;
    }

    int buildRec(std::vector<uint32_t>& order, uint32_t b, uint32_t e) {
        int me = (int)nodes.size();
        nodes.push_back(Node{-1, -1, b, e});
        lo.resize(lo.size() + D, std::numeric_limits<S>::infinity());
        hi.resize(hi.size() + D, -std::numeric_limits<S>::infinity());
        for (uint32_t i = b; i < e; ++i)
            for (int d = 0; d < D; ++d) {
                S v = pts[(size_t)order[i] * D + d];
                lo[(size_t)me * D + d] = std::min(lo[(size_t)me * D + d], v);
                hi[(size_t)me * D + d] = std::max(hi[(size_t)me * D + d], v);
            }
        if (e - b <= LEAF) return me;
        static thread_local std::vector<S> w;
        w = dimWeight();
        int axis = 0;
        S best = -1;
        for (int d = 0; d < D; ++d) {
            S ext = (hi[(size_t)me * D + d] - lo[(size_t)me * D + d]) * w[d];
            if (ext > best) best = ext, axis = d;
        }
        uint32_t mid = b + (e - b) / 2;
        std::nth_element(order.begin() + b, order.begin() + mid, order.begin() + e, [&](uint32_t x, uint32_t y) {
            S vx = pts[(size_t)x * D + axis], vy = pts[(size_t)y * D + axis];
            return vx < vy || (vx == vy && x < y);
        });
        int l = buildRec(order, b, mid);
        int r = buildRec(order, mid, e);
        nodes[me].left = l;
        nodes[me].right = r;
        return me;
    }

    // Lower bound of distance(q, p) over every p inside node `n`'s box.  Each step is monotone in
    // floating point w.r.t. the point coordinates, so bound <= computed distance for every member.
    S lowerBound(int n, const S* q) const {
        const S* L = &lo[(size_t)n * D];
        const S* H = &hi[(size_t)n * D];
        S total = 0;
        int off = 0;
        for (int i = 0; i < sp.n_parts; ++i) {
            const auto& part = sp.part[i];
            S d = 0;
            if (part.kind == MPTG_PART_LP) {
                S diff[MPTG_MAX_SCALARS];
                for (int j = 0; j < part.dim; ++j) {
                    S v = q[off + j];
                    diff[j] = v < L[off + j] ? L[off + j] - v : (v > H[off + j] ? v - H[off + j] : S(0));
                }
                d = lpNorm(diff, part.dim, part.p);
            } else if (part.kind == MPTG_PART_SO2) {
                const S pi = fp::consts<S>::pi();
                S diff[MPTG_MAX_SCALARS];
                for (int j = 0; j < part.dim; ++j) {
                    S v = q[off + j];
                    if (v >= L[off + j] && v <= H[off + j]) {
                        diff[j] = 0;
                    } else {
                        S d0 = fp::abs_(v - L[off + j]), d1 = fp::abs_(v - H[off + j]);
                        if (d0 > pi) d0 = S(2) * pi - d0;
                        if (d1 > pi) d1 = S(2) * pi - d1;
                        diff[j] = std::max(S(0), std::min(d0, d1));
                    }
                }
                d = lpNorm(diff, part.dim, part.p);
            } else {
                // |q.p| <= max(S(q), S(-q)), S(q) = sum_i max(q_i lo_i, q_i hi_i)  (support of the box)
                // Mirror distance()'s fma chain with the maximising / minimising box corner per
                // coordinate: fma(q,p,acc) is monotone in p (fixed q) and in acc, so
                // dotLo <= fl-dot(q,p) <= dotHi for every p in the box.
                S dotHi = 0, dotLo = 0;
                for (int j = 0; j < 4; ++j) {
                    S v = q[off + j];
                    S cHi = v >= S(0) ? H[off + j] : L[off + j];
                    S cLo = v >= S(0) ? L[off + j] : H[off + j];
                    if (j == 0) dotHi = v * cHi, dotLo = v * cLo;
                    else dotHi = fp::fma_(v, cHi, dotHi), dotLo = fp::fma_(v, cLo, dotLo);
                }
                S ad = std::max(dotHi, -dotLo);
                if (ad > S(1)) ad = S(1);
                if (ad < S(0)) ad = S(0);
                // acos01 is a polynomial and not proven monotone in the last ulp: shave the bound
                d = fp::acos01(ad) * (S(1) - S(8) * fp::consts<S>::eps());
            }
            if (part.weight != 1.0) d = d * S(part.weight);
            total = (i == 0) ? d : total + d;
            off += partScalars(part);
        }
        return total;
    }

    // returns number of distance evaluations
    uint64_t query(const S* q, uint32_t k, double radius, uint32_t* idxOut, S* distOut, uint32_t* countOut) const {
        using E = std::pair<S, uint32_t>;
        std::priority_queue<E> heap;  // max-heap on (d, idx)
        const bool bounded = radius >= 0 && std::isfinite(radius);
        const S r = S(radius);
        uint64_t evals = 0;
        struct Item {
            int node;
            S lb;
        };
        std::vector<Item> stack;
        if (!nodes.empty()) stack.push_back({0, lowerBound(0, q)});
        while (!stack.empty()) {
            Item it = stack.back();
            stack.pop_back();
            if (bounded && it.lb > r) continue;
            if (heap.size() == k && it.lb > heap.top().first) continue;
            const Node& n = nodes[it.node];
            if (n.left < 0) {
                for (uint32_t i = n.begin; i < n.end; ++i) {
                    S d = distance(sp, &pts[(size_t)i * D], q);
                    ++evals;
                    if (d != d) continue;
                    if (bounded && !(d <= r)) continue;
                    E e{d, ids[i]};
                    if (heap.size() < k) heap.push(e);
                    else if (e < heap.top()) {
                        heap.pop();
                        heap.push(e);
                    }
                }
            } else {
                S l0 = lowerBound(n.left, q), l1 = lowerBound(n.right, q);
                if (l0 <= l1) {
                    stack.push_back({n.right, l1});
                    stack.push_back({n.left, l0});
                } else {
                    stack.push_back({n.left, l0});
                    stack.push_back({n.right, l1});
                }
            }
        }
        uint32_t cnt = (uint32_t)heap.size();
        for (uint32_t j = k; j-- > 0;) {
            if (j < cnt) {
                idxOut[j] = heap.top().second;
                distOut[j] = heap.top().first;
                heap.pop();
            } else {
                idxOut[j] = MPTG_NO_INDEX;
                distOut[j] = std::numeric_limits<S>::infinity();
            }
        }
        if (countOut) *countOut = cnt;
        return evals;
    }
};

}  // namespace oracle
