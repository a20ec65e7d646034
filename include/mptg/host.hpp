// host.hpp -- C++17 RAII layer over the C ABI (mptg.h).  Header only; link with libmptg.so.
//
// Mirrors the reference's error behaviour: misuse and device failures surface as
// std::runtime_error (the reference throws std::runtime_error for misuse, e.g.
// src/mpt/impl/prrt/prrt.hpp:197-198).  There is no CPU fallback anywhere in this layer.
#pragma once

#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "mptg.h"

namespace mptg {

inline void check(int rc, const mptg_ctx* ctx, const char* what) {
    if (rc != MPTG_OK) throw std::runtime_error(std::string(what) + ": " + mptg_last_error(ctx) + " (mptg status " + std::to_string(rc) + ")");
}

class Context {
    mptg_ctx* h_ = nullptr;

public:
    explicit Context(int device = -1) { check(mptg_ctx_create(device, &h_), nullptr, "mptg_ctx_create"); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    Context(Context&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    ~Context() {
        if (h_) mptg_ctx_destroy(h_);
    }
    mptg_ctx* get() const { return h_; }
    void sync() { check(mptg_sync(h_), h_, "mptg_sync"); }
    std::uint64_t launches() const { return mptg_ctx_launch_count(h_); }
    // FFMA rate of the device in TFLOP/s (the FP32 yardstick of the bench's rooflines)
    double fp32Tflops() {
        double t = 0;
        check(mptg_probe_fp32_tflops(h_, &t), h_, "mptg_probe_fp32_tflops");
        return t;
    }
};

// Scenario geometry registered on the device (mptg_geom).  Move only.
class Geometry {
    mptg_ctx* ctx_ = nullptr;
    mptg_geom* h_ = nullptr;

public:
    Geometry() = default;
    Geometry(mptg_ctx* ctx, mptg_geom* h) : ctx_(ctx), h_(h) {}
    Geometry(const Geometry&) = delete;
    Geometry& operator=(const Geometry&) = delete;
    Geometry(Geometry&& o) noexcept : ctx_(o.ctx_), h_(o.h_) { o.h_ = nullptr; }
    Geometry& operator=(Geometry&& o) noexcept {
        if (this != &o) {
            if (h_) mptg_geom_destroy(h_);
            ctx_ = o.ctx_;
            h_ = o.h_;
            o.h_ = nullptr;
        }
        return *this;
    }
    ~Geometry() {
        if (h_) mptg_geom_destroy(h_);
    }
    mptg_geom* get() const { return h_; }
    explicit operator bool() const { return h_ != nullptr; }

    static Geometry grid(Context& c, int scalar, int width, int height, const std::uint8_t* occupancy) {
        mptg_geom* g;
        check(mptg_grid_create(c.get(), scalar, width, height, occupancy, &g), c.get(), "mptg_grid_create");
        return Geometry(c.get(), g);
    }
    static Geometry shapes(Context& c, int scalar, int dim, const std::vector<double>& centres, const std::vector<double>& radii,
                           const std::vector<double>& rects = {}) {
        mptg_geom* g;
        check(mptg_shapes_create(c.get(), scalar, dim, (int)radii.size(), centres.data(), radii.data(), (int)(rects.size() / 4),
                                 rects.data(), &g),
              c.get(), "mptg_shapes_create");
        return Geometry(c.get(), g);
    }
    static Geometry linkArm(Context& c, int scalar, const std::vector<double>& lengths, double linkRadius,
                            const std::vector<double>& cxcyr) {
        mptg_geom* g;
        check(mptg_linkarm_create(c.get(), scalar, (int)lengths.size(), lengths.data(), linkRadius, (int)(cxcyr.size() / 3),
                                  cxcyr.data(), &g),
              c.get(), "mptg_linkarm_create");
        return Geometry(c.get(), g);
    }
    static Geometry naoCup(Context& c, int scalar) {
        mptg_geom* g;
        check(mptg_naocup_create(c.get(), scalar, &g), c.get(), "mptg_naocup_create");
        return Geometry(c.get(), g);
    }
    static Geometry meshPair(Context& c, int scalar, const std::vector<float>& robotTris, const std::vector<float>& envTris) {
        mptg_geom* g;
        check(mptg_mesh_pair_create(c.get(), scalar, (std::uint32_t)(robotTris.size() / 9), robotTris.data(),
                                    (std::uint32_t)(envTris.size() / 9), envTris.data(), &g),
              c.get(), "mptg_mesh_pair_create");
        return Geometry(c.get(), g);
    }

    // scenario.valid(q) for n states (AoS, host); near (optional): the near-contact flags of mptg_valid_batch
    void valid(const void* states, std::uint32_t n, std::uint8_t* ok, std::uint8_t* near = nullptr) const {
        check(mptg_valid_batch(h_, states, n, ok, near), ctx_, "mptg_valid_batch");
    }
    // scenario.link(a, b) for n edges
    void link(const mptg_space_desc* space, const void* from, const void* to, std::uint32_t n, double step, std::uint8_t* ok,
              std::uint8_t* near = nullptr) const {
        check(mptg_link_batch(h_, space, from, to, n, step, ok, near), ctx_, "mptg_link_batch");
    }
};

// Batched nearest-neighbour structure with the shape of nigh::Nigh (insert / size / nearest).
// T is the caller's node handle type (the reference stores Node*); handles are kept on the host in
// insertion order, the device identifies nodes by that dense index.
template <typename T, typename Space>
class Nearest {
    using State = typename Space::Type;
    using Distance = typename Space::Distance;
    mptg_ctx* ctx_;
    mptg_knn* h_ = nullptr;
    mptg_space_desc desc_;
    std::vector<T> handles_;
    std::vector<std::uint32_t> idx_;
    std::vector<Distance> dist_;
    std::vector<std::uint32_t> cnt_;

public:
    Nearest(Context& ctx, const Space& space, std::uint32_t capacity, int strategy = MPTG_KNN_AUTO) : ctx_(ctx.get()), desc_(space.desc()) {
        check(mptg_knn_create(ctx_, &desc_, capacity, &h_), ctx_, "mptg_knn_create");
        if (strategy != MPTG_KNN_AUTO) check(mptg_knn_set_strategy(h_, strategy), ctx_, "mptg_knn_set_strategy");
    }
    Nearest(const Nearest&) = delete;
    Nearest& operator=(const Nearest&) = delete;
    ~Nearest() {
        if (h_) mptg_knn_destroy(h_);
    }

    std::size_t size() const { return mptg_knn_size(h_); }
    const T& handle(std::uint32_t index) const { return handles_[index]; }

    // nn.insert(node): states[i] belongs to handles[i]
    std::uint32_t insert(const State* states, const T* handles, std::uint32_t count) {
        std::uint32_t first = 0;
        check(mptg_knn_insert(h_, states, count, &first), ctx_, "mptg_knn_insert");
        handles_.insert(handles_.end(), handles, handles + count);
        return first;
    }
    std::uint32_t insert(const State& state, const T& handle) { return insert(&state, &handle, 1); }

    // nn.nearest(nbh, q, k, r) for a batch: row q of (indices, distances) has counts()[q] entries,
    // ascending by (distance, index).  The vectors are owned by this object and reused.
    void nearest(const State* queries, std::uint32_t Q, std::uint32_t k, Distance radius = std::numeric_limits<Distance>::infinity()) {
        idx_.resize((std::size_t)Q * k);
        dist_.resize((std::size_t)Q * k);
        cnt_.resize(Q);
        const double r = radius < std::numeric_limits<Distance>::infinity() ? (double)radius : -1.0;
        check(mptg_knn_query(h_, queries, Q, k, r, idx_.data(), dist_.data(), cnt_.data()), ctx_, "mptg_knn_query");
    }
    const std::vector<std::uint32_t>& indices() const { return idx_; }
    const std::vector<Distance>& distances() const { return dist_; }
    const std::vector<std::uint32_t>& counts() const { return cnt_; }

    // nn.nearest(q): single query convenience (one wave of one)
    bool nearest(const State& q, T* handleOut, Distance* distOut) {
        nearest(&q, 1, 1);
        if (cnt_[0] == 0) return false;
        *handleOut = handles_[idx_[0]];
        *distOut = dist_[0];
        return true;
    }
};

}  // namespace mptg
