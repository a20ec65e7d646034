// topk.cuh -- warp-cooperative sorted top-k list held in registers (one warp per query).
//
// Slot j of the ascending list lives in lane (j % 32), register (j / 32); KPL registers per lane give
// room for 32*KPL >= k slots.  Order is the total order (distance, index): this is the tie rule of
// the kNN contract (north_star: "ties broken by node index"), so results do not depend on the order
// candidates are visited in (brute-force scan, BVH traversal, multi-GPU merge all agree).
// Empty slots hold (+inf, MPTG_NO_INDEX), which compares greater than any real candidate.
#pragma once

#include <cuda_runtime.h>

#include "../../include/mptg/mptg.h"
#include "../../include/mptg/mptg_fpmath.h"

namespace mptg {

constexpr unsigned FULL_MASK = 0xffffffffu;

template <typename S, int KPL>
struct WarpTopK {
    S d[KPL];
    uint32_t i[KPL];
    S kthD;         // warp-uniform: distance in slot k-1
    uint32_t kthI;  // warp-uniform: index in slot k-1
    uint32_t k1;    // k - 1: slot k-1 lives in register k1 / 32 of lane k1 % 32

    __device__ __forceinline__ void init(uint32_t k) {
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
            d[r] = fp::consts<S>::inf();
            i[r] = MPTG_NO_INDEX;
        }
        kthD = fp::consts<S>::inf();
        kthI = MPTG_NO_INDEX;
        k1 = k - 1;
    }

    // (nd, ni) < current k-th in the total order
    __device__ __forceinline__ bool beats(S nd, uint32_t ni) const {
        return nd < kthD || (nd == kthD && ni < kthI);
    }

    // Insert a warp-uniform candidate.  Every lane must call this with the same (nd, ni).
    __device__ __forceinline__ void insert(S nd, uint32_t ni, int lane) {
        unsigned carryGt = 0;  // "slot before register r's lane 0 is greater": bit of lane 31, previous register
        S carryD = S(0);
        uint32_t carryI = 0;
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
            const bool gt = d[r] > nd || (d[r] == nd && i[r] > ni);
            const unsigned m = __ballot_sync(FULL_MASK, gt);
            S pd = __shfl_up_sync(FULL_MASK, d[r], 1);
            uint32_t pi = __shfl_up_sync(FULL_MASK, i[r], 1);
            bool pgt = lane > 0 ? ((m >> (lane - 1)) & 1u) : (carryGt != 0);
            if (lane == 0) {
                pd = carryD;
                pi = carryI;
            }
            if (r + 1 < KPL) {  // carry the OLD last slot of this register row to the next row
                carryD = __shfl_sync(FULL_MASK, d[r], 31);
                carryI = __shfl_sync(FULL_MASK, i[r], 31);
                carryGt = (m >> 31) & 1u;
            }
            if (gt) {
                d[r] = pgt ? pd : nd;
                i[r] = pgt ? pi : ni;
            }
        }
        S kd = d[0];
        uint32_t ki = i[0];
        const int kr = (int)(k1 >> 5);
#pragma unroll
        for (int r = 1; r < KPL; ++r)
            if (r == kr) {
                kd = d[r];
                ki = i[r];
            }
        kthD = __shfl_sync(FULL_MASK, kd, (int)(k1 & 31u));
        kthI = __shfl_sync(FULL_MASK, ki, (int)(k1 & 31u));
    }

    // Offer one candidate per lane (cand == false for lanes without one); inserts those that beat the
    // current k-th, lowest lane first.  radius: only d <= radius qualifies.
    __device__ __forceinline__ void offer(bool cand, S nd, uint32_t ni, S radius, int lane) {
        cand = cand && nd <= radius && beats(nd, ni);
        unsigned m = __ballot_sync(FULL_MASK, cand);
        while (m) {
            const int src = __ffs(m) - 1;
            const S bd = __shfl_sync(FULL_MASK, nd, src);
            const uint32_t bi = __shfl_sync(FULL_MASK, ni, src);
            insert(bd, bi, lane);
            cand = cand && lane != src && beats(nd, ni);
            m = __ballot_sync(FULL_MASK, cand);
        }
    }

    // Write the first k slots of the list for one query; returns (via all lanes) the number of
    // real entries.
    __device__ __forceinline__ uint32_t store(uint32_t k, uint32_t* idxOut, S* distOut, int lane) const {
        uint32_t count = 0;
#pragma unroll
        for (int r = 0; r < KPL; ++r) {
            const uint32_t slot = (uint32_t)r * 32u + (uint32_t)lane;
            const bool real = slot < k && i[r] != MPTG_NO_INDEX;
            if (slot < k) {
                idxOut[slot] = i[r];
                distOut[slot] = d[r];
            }
            count += __popc(__ballot_sync(FULL_MASK, real));
        }
        return count;
    }
};

}  // namespace mptg
