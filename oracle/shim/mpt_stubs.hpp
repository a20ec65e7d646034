// mpt_stubs.hpp -- stand-ins (ours) selected through the reference's include guards for the two reference headers
// that need the real Eigen: mpt/log.hpp (Eigen stream I/O) and mpt/box_bounds.hpp (Eigen::DenseBase expressions).
// TEST INFRASTRUCTURE ONLY.  Neither is on the hot path: logging is a sink that drops everything, BoxBounds keeps
// the two corner vectors.
#pragma once
#include <cstddef>
#include <iostream>

#include <Eigen/Dense>

#define MPT_LOG_HPP_
namespace unc::robotics::mpt::log {
template <class T>
const char* type_name() { return "?"; }
struct Event {
    template <class T>
    Event& operator<<(const T&) { return *this; }
};
}  // namespace unc::robotics::mpt::log
#define MPT_LOG(...) if (true) {} else ::unc::robotics::mpt::log::Event()

#define MPT_BOX_BOUNDS_HPP
namespace unc::robotics::mpt {
template <typename S, int dim>
class BoxBounds {
    Eigen::Matrix<S, dim, 1> min_, max_;

public:
    BoxBounds() {}
    BoxBounds(const Eigen::Matrix<S, dim, 1>& mn, const Eigen::Matrix<S, dim, 1>& mx) : min_(mn), max_(mx) {}
    unsigned size() const { return dim; }
    S measure() const {  // box_bounds.hpp:88-90: (max - min).prod()
        S m = max_[0] - min_[0];
        for (int i = 1; i < dim; ++i) m = m * (max_[i] - min_[i]);
        return m;
    }
    const Eigen::Matrix<S, dim, 1>& min() const { return min_; }
    const Eigen::Matrix<S, dim, 1>& max() const { return max_; }
};
}  // namespace unc::robotics::mpt
