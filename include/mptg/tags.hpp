// tags.hpp -- option tags shared by the planner layer (planner.hpp) and the reference-side binding (nigh_binding.hpp).
#pragma once

namespace mptg {
// the nearest-neighbour strategy tag of this library (the analogue of nigh::KDTreeBatch<> etc.)
struct GpuBatch {};
}  // namespace mptg
