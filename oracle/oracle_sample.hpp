// oracle_sample.hpp -- CPU restatement of the reference's uniform samplers over a counter-based
// generator.  TEST INFRASTRUCTURE ONLY (see oracle.hpp header): only tests/, __graft_entry__.smoke()
// and bench.py's CPU-baseline leg may use it.
//
// Distributions (reference): box coordinates uniform_real_distribution(min,max), one draw per
// coordinate in index order (src/mpt/uniform_box_sampler.hpp:60-68); SO(2) coordinates
// uniform_real_distribution(-pi,pi) (src/mpt/impl/uniform_sampler_so2.hpp:58-64); SO(3): a~U[0,1),
// b,c~U[0,2pi), quaternion (w,x,y,z) = (sqrt(1-a) sin b, sqrt(1-a) cos b, sqrt(a) sin c, sqrt(a) cos c)
// (src/mpt/impl/uniform_sampler_so3.hpp:55-68); compound spaces sample their parts in order
// (src/mpt/impl/uniform_sampler_cartesian.hpp:75-78).  uniform_real_distribution(a,b)(g) is
// generate_canonical(g) * (b - a) + a (libstdc++ bits/random.h).
// Generator: the reference seeds a std::mt19937_64 per worker from std::random_device, so it defines no
// sequence.  The product's generator is specified in include/mptg/mptg.h (Philox4x32-10 counter stream);
// it is restated here from the published algorithm (Salmon et al., "Parallel random numbers: as easy as
// 1, 2, 3", SC'11): 10 rounds of
//   (c0,c1,c2,c3) <- (hi(M1*c2)^c1^k0, lo(M1*c2), hi(M0*c0)^c3^k1, lo(M0*c0)),  k0 += W0, k1 += W1
// with M0 = 0xD2511F53, M1 = 0xCD9E8D57, W0 = 0x9E3779B9, W1 = 0xBB67AE85.
#pragma once

#include <cstdint>

#include "oracle.hpp"

namespace oracle {

struct Philox {
    uint32_t c[4];
    static void block(uint64_t seed, uint64_t g, uint32_t blk, uint32_t out[4]) {
        uint32_t c0 = (uint32_t)g, c1 = (uint32_t)(g >> 32), c2 = blk, c3 = 0;
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
        for (int r = 0; r < 10; ++r) {
            const uint64_t a = (uint64_t)0xD2511F53u * c0, b = (uint64_t)0xCD9E8D57u * c2;
            const uint32_t n0 = (uint32_t)(b >> 32) ^ c1 ^ k0, n1 = (uint32_t)b, n2 = (uint32_t)(a >> 32) ^ c3 ^ k1, n3 = (uint32_t)a;
            c0 = n0, c1 = n1, c2 = n2, c3 = n3;
            k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
        }
        out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
    }
};

// the words of sample g's stream, then uniforms in [0,1): float from the top 24 bits of one word, double
// from the top 53 bits of two
template <typename S>
struct SampleStream {
    uint64_t seed, g;
    uint32_t w[4];
    uint32_t pos = 0;
    SampleStream(uint64_t s, uint64_t g_) : seed(s), g(g_) {}
    uint32_t word() {
        if (pos % 4 == 0) Philox::block(seed, g, pos / 4, w);
        return w[pos++ % 4];
    }
    S next();
};
template <>
inline float SampleStream<float>::next() {
    return (float)(word() >> 8) * (1.0f / 16777216.0f);
}
template <>
inline double SampleStream<double>::next() {
    const uint64_t hi = word(), lo = word();
    return (double)(((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

// uniforms -> state (ABI order: SO3 parts as x y z w)
template <typename S, typename Next>
void sampleFromUniforms(const mptg_space_desc& sp, const double* lo, const double* hi, Next&& next, S* q) {
    int off = 0;
    for (int i = 0; i < sp.n_parts; ++i) {
        const auto& part = sp.part[i];
        if (part.kind == MPTG_PART_SO3) {
            const S a = next();
            const S twoPi = S(2) * mptg::fp::consts<S>::pi();
            const S b = next() * twoPi, c = next() * twoPi;
            S sb, cb, sc, cc;
            mptg::fp::sincos_(b, &sb, &cb);
            mptg::fp::sincos_(c, &sc, &cc);
            const S r1 = mptg::fp::sqrt_(S(1) - a), r2 = mptg::fp::sqrt_(a);
            q[off + 0] = r1 * cb, q[off + 1] = r2 * sc, q[off + 2] = r2 * cc, q[off + 3] = r1 * sb;
            off += 4;
        } else if (part.kind == MPTG_PART_SO2) {
            const S pi = mptg::fp::consts<S>::pi();
            for (int c = 0; c < part.dim; ++c) q[off + c] = next() * (pi - (-pi)) + (-pi);
            off += part.dim;
        } else {
            for (int c = 0; c < part.dim; ++c) {
                const S l = (S)lo[off + c], h = (S)hi[off + c];
                q[off + c] = next() * (h - l) + l;
            }
            off += part.dim;
        }
    }
}

// sample number g; uniform 0 is the goal-bias draw (src/mpt/impl/prrt/prrt.hpp:377-379)
template <typename S>
void sampleState(const mptg_space_desc& sp, const double* lo, const double* hi, uint64_t seed, uint64_t g, const S* goal, double goalBias,
                 S* q) {
    SampleStream<S> st(seed, g);
    const S draw = st.next();
    if (goal && draw < (S)goalBias) {
        int D = 0;
        for (int i = 0; i < sp.n_parts; ++i) D += sp.part[i].kind == MPTG_PART_SO3 ? 4 : sp.part[i].dim;
        for (int c = 0; c < D; ++c) q[c] = goal[c];
        return;
    }
    sampleFromUniforms<S>(sp, lo, hi, [&]() { return st.next(); }, q);
}

}  // namespace oracle
