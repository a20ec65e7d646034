// knn_build.cu -- device-side construction of the kNN index (float32 spaces).
//
// Same structure as the host build in knn_bvh.cuh (kd-style median splits on the widest weighted
// coordinate, split positions aligned to powers of 32, 32 points per leaf, 32 children per node),
// built without leaving the GPU: one pass per binary level
//     segment extents (block-reduced min/max, then atomics)  ->  split axis per segment
//     ->  64-bit keys (segment, ordered coordinate)  ->  radix sort of (key, point id)
// followed by kernels that emit the blocked leaf / box arrays, the half-precision copies and their
// error bounds.  The radix sort is cub::DeviceRadixSort (library code, build path only -- the search
// kernels are hand written).  Any permutation gives a correct index (the search is exact for any
// boxes that contain their members); the ordering only decides how well it prunes.
#include <cub/device/device_radix_sort.cuh>
#include <cuda_fp16.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "../../include/mptg/mptg_space.h"
#include "knn_index.cuh"

namespace mptg {
namespace {

__device__ __forceinline__ int orderedInt(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float fromOrderedInt(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }
__device__ __forceinline__ uint32_t orderedUint(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct SegLevel {        // one binary level: segments in position order
    const uint32_t* begin;   // [nSeg]
    const uint32_t* mid;     // [nSeg]  (== end when the segment is not split any further)
    const uint32_t* end;     // [nSeg]
    const uint32_t* childBase;  // [nSeg] index of the first child segment in the next level
    uint32_t nSeg;
};

// canonical AoS copy: SO(3) parts flipped to w >= 0.  `canon` (float) drives the partition (extents, sort keys);
// double-precision sets also keep the exact values in `exact` for the emitted leaves and boxes -- rounding to
// float is monotone, so any partition of the rounded values is a valid partition of the exact ones.
template <typename S>
__global__ void canonKernel(DevSpace<float> sp, const S* __restrict__ pts, uint32_t stride, uint32_t n, float* __restrict__ canon,
                            S* __restrict__ exact) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int p = 0; p < sp.nParts; ++p) {
        const int off = sp.off[p];
        bool flip = false;
        if (sp.kind[p] == MPTG_PART_SO3) flip = pts[(size_t)(off + 3) * stride + i] < S(0);
        for (int j = 0; j < sp.dim[p]; ++j) {
            const S v = pts[(size_t)(off + j) * stride + i];
            const S w = flip ? -v : v;
            canon[(size_t)i * sp.D + off + j] = (float)w;
            if (sizeof(S) == 8) exact[(size_t)i * sp.D + off + j] = w;
        }
    }
}

__global__ void iotaKernel(uint32_t* ids, uint32_t* segOf, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ids[i] = i, segOf[i] = 0;
}

__global__ void resetExtentKernel(int* segMin, int* segMax, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) segMin[i] = 0x7fffffff, segMax[i] = (int)0x80000000;
}

// per-segment min / max of every coordinate; a block whose positions all lie in one segment reduces first
constexpr int EXT_THREADS = 256;
__global__ void __launch_bounds__(EXT_THREADS) segExtentKernel(const float* __restrict__ canon, const uint32_t* __restrict__ ids,
                                                               const uint32_t* __restrict__ segOf, uint32_t n, int D, int* segMin,
                                                               int* segMax) {
    __shared__ int sMin[EXT_THREADS / 32], sMax[EXT_THREADS / 32];
    const uint32_t i = blockIdx.x * EXT_THREADS + threadIdx.x;
    const uint32_t first = blockIdx.x * EXT_THREADS, last = min(n, first + EXT_THREADS) - 1;
    const bool uniform = segOf[first] == segOf[last];  // segments are contiguous position ranges
    const bool have = i < n;
    const uint32_t seg = have ? segOf[i] : 0;
    const float* pt = have ? canon + (size_t)ids[i] * D : canon;
    for (int c = 0; c < D; ++c) {
        const int v = have ? orderedInt(pt[c]) : 0;
        if (uniform) {
            int mn = have ? v : 0x7fffffff, mx = have ? v : (int)0x80000000;
            for (int o = 16; o > 0; o >>= 1) {
                mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
                mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            }
            if ((threadIdx.x & 31) == 0) sMin[threadIdx.x >> 5] = mn, sMax[threadIdx.x >> 5] = mx;
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int w = 1; w < EXT_THREADS / 32; ++w) mn = min(mn, sMin[w]), mx = max(mx, sMax[w]);
                const uint32_t s0 = segOf[first];
                atomicMin(segMin + (size_t)s0 * D + c, mn);
                atomicMax(segMax + (size_t)s0 * D + c, mx);
            }
            __syncthreads();
        } else if (have) {
            atomicMin(segMin + (size_t)seg * D + c, v);
            atomicMax(segMax + (size_t)seg * D + c, v);
        }
    }
}

// rotSplit: extents of SO(3) coefficients count rotSplit times their weight when the split axis is chosen.  With the
// cap bound a finer rotation partition prunes better than equal weighted extents (tools/knn_tree_shape_experiment.py:
// floor of the leaves a C5 query must visit 125 -> 104 at 1.4, 101 at 2.0).
__global__ void axisKernel(DevSpace<float> sp, const int* __restrict__ segMin, const int* __restrict__ segMax, uint32_t nSeg,
                           uint32_t* __restrict__ axis, float rotSplit) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nSeg) return;
    int best = 0;
    float bestExt = -1.0f;
    for (int p = 0; p < sp.nParts; ++p)
        for (int j = 0; j < sp.dim[p]; ++j) {
            const int c = sp.off[p] + j;
            float ext = (fromOrderedInt(segMax[(size_t)s * sp.D + c]) - fromOrderedInt(segMin[(size_t)s * sp.D + c])) * sp.weight[p];
            if (sp.kind[p] == MPTG_PART_SO3) ext *= rotSplit;
            if (ext > bestExt) bestExt = ext, best = c;
        }
    axis[s] = (uint32_t)best;
}

__global__ void keyKernel(const float* __restrict__ canon, const uint32_t* __restrict__ ids, const uint32_t* __restrict__ segOf,
                          const uint32_t* __restrict__ axis, uint32_t n, int D, unsigned long long* __restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t seg = segOf[i];
    keys[i] = ((unsigned long long)seg << 32) | orderedUint(canon[(size_t)ids[i] * D + axis[seg]]);
}

__global__ void splitKernel(SegLevel lv, uint32_t* __restrict__ segOf, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = segOf[i];
    segOf[i] = lv.childBase[s] + (i >= lv.mid[s] ? 1u : 0u);
}

// ---- emit the device image
// one warp per leaf: blocked points, perm, leaf box (SoA per level: lo[c][nNodes], hi[c][nNodes])
template <typename S>
__global__ void __launch_bounds__(256) leafEmitKernel(const S* __restrict__ canon, const uint32_t* __restrict__ ids, uint32_t n, int D,
                                                      uint32_t nLeaves, S* __restrict__ leafPts, uint32_t* __restrict__ perm,
                                                      S* __restrict__ lo, S* __restrict__ hi, uint32_t idBase = 0) {
    const uint32_t leaf = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (leaf >= nLeaves) return;
    const uint32_t p = leaf * 32u + lane;
    const bool real = p < n;
    const uint32_t src = ids[real ? p : n - 1];  // padding repeats the last point (perm marks it unused)
    perm[p] = real ? src + idBase : MPTG_NO_INDEX;
    for (int c = 0; c < D; ++c) {
        const S v = canon[(size_t)src * D + c];
        leafPts[((size_t)leaf * D + c) * 32u + lane] = v;
        S mn = v, mx = v;
        for (int o = 16; o > 0; o >>= 1) {
            const S a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
            mn = a < mn ? a : mn;
            mx = b > mx ? b : mx;
        }
        if (lane == 0) lo[(size_t)c * nLeaves + leaf] = mn, hi[(size_t)c * nLeaves + leaf] = mx;
    }
}

// one warp per parent: box of its (up to) 32 children
template <typename S>
__global__ void __launch_bounds__(256) parentBoxKernel(const S* __restrict__ clo, const S* __restrict__ chi, uint32_t nChild, int D,
                                                       uint32_t nParent, S* __restrict__ lo, S* __restrict__ hi) {
    const uint32_t parent = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (parent >= nParent) return;
    const uint32_t child = parent * 32u + lane;
    for (int c = 0; c < D; ++c) {
        S mn = child < nChild ? clo[(size_t)c * nChild + child] : (S)INFINITY;
        S mx = child < nChild ? chi[(size_t)c * nChild + child] : (S)-INFINITY;
        for (int o = 16; o > 0; o >>= 1) {
            const S a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
            mn = a < mn ? a : mn;
            mx = b > mx ? b : mx;
        }
        if (lane == 0) lo[(size_t)c * nParent + parent] = mn, hi[(size_t)c * nParent + parent] = mx;
    }
}

// SoA boxes -> blocked [block][2D][32]
template <typename S>
__global__ void boxBlockKernel(const S* __restrict__ lo, const S* __restrict__ hi, uint32_t nNodes, int D, S* __restrict__ box) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;  // node slot, padded to a multiple of 32
    const uint32_t nSlots = ((nNodes + 31u) / 32u) * 32u;
    if (j >= nSlots) return;
    const uint32_t b = j >> 5, ln = j & 31;
    for (int c = 0; c < D; ++c) {
        const S l = j < nNodes ? lo[(size_t)c * nNodes + j] : (S)INFINITY;
        const S h = j < nNodes ? hi[(size_t)c * nNodes + j] : (S)-INFINITY;
        box[((size_t)b * 2 * D + c) * 32u + ln] = l;
        box[((size_t)b * 2 * D + D + c) * 32u + ln] = h;
    }
}

// ---- SE(3)/f32: rotation caps.  Any centre gives a valid cap (the radius is measured against the centre as stored);
// the centre only decides how small the radius is.  q and -q are the same rotation, so members are sign-aligned to a
// reference before they are averaged.
// one warp per leaf: centre = normalised sum of the leaf's quaternions aligned to its first point
__global__ void __launch_bounds__(256) leafCapCentreKernel(const float* __restrict__ leafPts, uint32_t nLeaves, float4* __restrict__ centre) {
    const uint32_t leaf = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (leaf >= nLeaves) return;
    const float* pt = leafPts + ((size_t)leaf * 7u) * 32u + lane;
    float q[4];
    for (int c = 0; c < 4; ++c) q[c] = pt[c * 32];
    float d = 0.0f;
    for (int c = 0; c < 4; ++c) d += q[c] * __shfl_sync(0xffffffffu, q[c], 0);
    const float sg = d < 0.0f ? -1.0f : 1.0f;
    for (int c = 0; c < 4; ++c) {
        q[c] *= sg;
        for (int o = 16; o > 0; o >>= 1) q[c] += __shfl_xor_sync(0xffffffffu, q[c], o);
    }
    if (lane == 0) {
        const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        centre[leaf] = n > 0.0f && n < INFINITY ? make_float4(q[0] / n, q[1] / n, q[2] / n, q[3] / n) : make_float4(0.f, 0.f, 0.f, 1.f);
    }
}
// one warp per parent: centre = normalised sum of the children's centres aligned to the first child's
__global__ void __launch_bounds__(256) parentCapCentreKernel(const float4* __restrict__ child, uint32_t nChild, uint32_t nParent,
                                                             float4* __restrict__ centre) {
    const uint32_t parent = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (parent >= nParent) return;
    const uint32_t ch = parent * 32u + lane;
    const float4 v = ch < nChild ? child[ch] : make_float4(0.f, 0.f, 0.f, 0.f);
    float q[4] = {v.x, v.y, v.z, v.w};
    float d = 0.0f;
    for (int c = 0; c < 4; ++c) d += q[c] * __shfl_sync(0xffffffffu, q[c], 0);
    const float sg = d < 0.0f ? -1.0f : 1.0f;
    for (int c = 0; c < 4; ++c) {
        q[c] *= sg;
        for (int o = 16; o > 0; o >>= 1) q[c] += __shfl_xor_sync(0xffffffffu, q[c], o);
    }
    if (lane == 0) {
        const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        centre[parent] = n > 0.0f && n < INFINITY ? make_float4(q[0] / n, q[1] / n, q[2] / n, q[3] / n) : make_float4(0.f, 0.f, 0.f, 1.f);
    }
}
// one thread per stored point (padding repeats a real point): cosine of its angle to the centre of its level-l node, in
// double (exact to 1e-15 of the real geometry of the stored floats), rounded down to float; minimum per node through
// the bit pattern (non-negative floats order like unsigned integers).  Level 0 also takes the largest quaternion norm.
__global__ void __launch_bounds__(256) capRadiusKernel(const float* __restrict__ leafPts, uint32_t nPad, int shift, const float4* __restrict__ centre,
                                                       unsigned int* __restrict__ minCos, unsigned int* __restrict__ normMax) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nPad) return;  // nPad is a multiple of 32: whole warps leave together
    const uint32_t leaf = p >> 5, ln = p & 31;
    const float* pt = leafPts + ((size_t)leaf * 7u) * 32u + ln;
    const float4 c = centre[p >> shift];
    const double q0 = pt[0], q1 = pt[32], q2 = pt[64], q3 = pt[96];
    const double dot = fabs(q0 * c.x + q1 * c.y + q2 * c.z + q3 * c.w);
    const double n2 = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3, c2 = (double)c.x * c.x + (double)c.y * c.y + (double)c.z * c.z + (double)c.w * c.w;
    double cs = dot / sqrt(n2 * c2);
    if (!(cs >= 0.0)) cs = 0.0;  // zero norm or NaN: no rotation pruning for this node
    if (cs > 1.0) cs = 1.0;
    float cf = __double2float_rd(cs);
    float nf = __double2float_ru(sqrt(n2));
    if (!(nf >= 0.0f)) nf = INFINITY;  // NaN coordinates: the caller falls back to the box path
    for (int o = 16; o > 0; o >>= 1) {
        cf = fminf(cf, __shfl_xor_sync(0xffffffffu, cf, o));
        nf = fmaxf(nf, __shfl_xor_sync(0xffffffffu, nf, o));
    }
    if (ln == 0) {
        atomicMin(minCos + (p >> shift), __float_as_uint(cf));
        if (normMax) atomicMax(normMax, __float_as_uint(nf));
    }
}
__global__ void fillKernel(unsigned int* a, unsigned int v, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}
// blocked cap image of one level: [block][3][32] float4
__global__ void capPackKernel(const float4* __restrict__ centre, const unsigned int* __restrict__ minCos, const float* __restrict__ lo,
                              const float* __restrict__ hi, uint32_t nNodes, float4* __restrict__ cap) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nSlots = ((nNodes + 31u) / 32u) * 32u;
    if (j >= nSlots) return;
    const uint32_t b = j >> 5, ln = j & 31;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 1.f), v1 = make_float4(0.f, 1.f, INFINITY, -INFINITY), v2 = make_float4(INFINITY, -INFINITY, INFINITY, -INFINITY);
    if (j < nNodes) {
        v0 = centre[j];
        const double cs = (double)__uint_as_float(minCos[j]);
        const float sn = fminf(1.0f, __double2float_ru(sqrt(fmax(0.0, 1.0 - cs * cs))));
        v1 = make_float4((float)cs, sn, lo[(size_t)4 * nNodes + j], hi[(size_t)4 * nNodes + j]);
        v2 = make_float4(lo[(size_t)5 * nNodes + j], hi[(size_t)5 * nNodes + j], lo[(size_t)6 * nNodes + j], hi[(size_t)6 * nNodes + j]);
    }
    cap[((size_t)b * 3 + 0) * 32u + ln] = v0;
    cap[((size_t)b * 3 + 1) * 32u + ln] = v1;
    cap[((size_t)b * 3 + 2) * 32u + ln] = v2;
}

// SE(3): half2 words of the leaf points ([leaf][32] uint4, one 128-bit load per lane) + max conversion error of
// quaternion / translation coordinates
// power of two that brings the largest |translation coordinate| of the set into (1024, 2048]: the half copies then
// keep 11 significant bits whatever the unit of length is, and differences of in-range coordinates cannot overflow
__global__ void tScaleKernel(const float* __restrict__ lo, const float* __restrict__ hi, uint32_t nTop, float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float m = 0.0f;
    for (int c = 4; c < 7; ++c)
        for (uint32_t j = 0; j < nTop; ++j) m = fmaxf(m, fmaxf(fabsf(lo[(size_t)c * nTop + j]), fabsf(hi[(size_t)c * nTop + j])));
    float sc = 1.0f;
    if (m > 0.0f && m < 1e30f) {
        int e;
        frexpf(m, &e);  // m = f 2^e, f in [0.5, 1)
        sc = ldexpf(1.0f, 11 - e);
    }
    *out = sc;
}
__global__ void leafHalfKernel(const float* __restrict__ leafPts, uint32_t nPad, uint32_t* __restrict__ leafH, unsigned int* __restrict__ err) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    float eq = 0.0f, et = 0.0f;
    const float sc = __uint_as_float(err[3]);
    if (p < nPad) {
        const uint32_t leaf = p >> 5, ln = p & 31;
        unsigned short h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 7; ++c) {
            const float v = leafPts[((size_t)leaf * 7 + c) * 32u + ln] * (c < 4 ? 1.0f : sc);  // exact: sc is a power of two
            const __half hv = __float2half_rn(v);
            h[c] = __half_as_ushort(hv);
            const float e = fabsf(__half2float(hv) - v);
            if (c < 4) eq = fmaxf(eq, e);
            else et = fmaxf(et, e);
        }
        for (int r = 0; r < 4; ++r) leafH[(size_t)p * 4u + r] = (uint32_t)h[2 * r] | ((uint32_t)h[2 * r + 1] << 16);
    }
    for (int o = 16; o > 0; o >>= 1) {
        eq = fmaxf(eq, __shfl_xor_sync(0xffffffffu, eq, o));
        et = fmaxf(et, __shfl_xor_sync(0xffffffffu, et, o));
    }
    if ((threadIdx.x & 31) == 0) {  // non-negative floats order like unsigned ints
        atomicMax(err + 0, __float_as_uint(eq));
        atomicMax(err + 1, __float_as_uint(et));
    }
}

struct Levels {
    std::vector<uint32_t> begin, mid, end, childBase;  // concatenated over levels
    std::vector<uint32_t> levelOffset, levelCount;
};

// segment lists per binary level, same split rule as HostBuild::split (knn_bvh.cuh)
Levels makeLevels(uint32_t n) {
    Levels L;
    std::vector<std::pair<uint32_t, uint32_t>> cur{{0u, n}};
    for (;;) {
        bool any = false;
        std::vector<std::pair<uint32_t, uint32_t>> next;
        L.levelOffset.push_back((uint32_t)L.begin.size());
        L.levelCount.push_back((uint32_t)cur.size());
        for (auto [b, e] : cur) {
            const uint32_t len = e - b;
            L.begin.push_back(b);
            L.end.push_back(e);
            L.childBase.push_back((uint32_t)next.size());
            if (len <= 32) {
                L.mid.push_back(e);
                next.push_back({b, e});
            } else {
                uint32_t blk = 32;
                while ((uint64_t)blk * 32 < len) blk *= 32;
                const uint32_t nblk = (len + blk - 1) / blk;
                const uint32_t mid = b + ((nblk + 1) / 2) * blk;
                L.mid.push_back(mid);
                next.push_back({b, mid});
                next.push_back({mid, e});
                any = true;
            }
        }
        if (!any) break;
        cur.swap(next);
    }
    return L;
}


// ---- tail chunks (KnnTail): Morton order of one batch of new points, 32-point leaves with boxes
// the (up to) three coordinates of largest weighted extent, with offset and scale to 8 bits each (24-bit keys: three
// radix passes; a 256^3 grid is plenty for a tail of at most 65536 points)
__global__ void mortonAxesKernel(DevSpace<float> sp, const int* __restrict__ segMin, const int* __restrict__ segMax, int* __restrict__ axes,
                                 float* __restrict__ offScale) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float best[3] = {-1.0f, -1.0f, -1.0f};
    int ax[3] = {-1, -1, -1};
    for (int p = 0; p < sp.nParts; ++p)
        for (int j = 0; j < sp.dim[p]; ++j) {
            const int c = sp.off[p] + j;
            const float ext = (fromOrderedInt(segMax[c]) - fromOrderedInt(segMin[c])) * sp.weight[p];
            for (int r = 0; r < 3; ++r)
                if (ext > best[r]) {
                    for (int t = 2; t > r; --t) best[t] = best[t - 1], ax[t] = ax[t - 1];
                    best[r] = ext, ax[r] = c;
                    break;
                }
        }
    for (int r = 0; r < 3; ++r) {
        axes[r] = ax[r];
        const float mn = ax[r] >= 0 ? fromOrderedInt(segMin[ax[r]]) : 0.0f, mx = ax[r] >= 0 ? fromOrderedInt(segMax[ax[r]]) : 0.0f;
        offScale[2 * r] = mn;
        offScale[2 * r + 1] = mx > mn ? 255.0f / (mx - mn) : 0.0f;
    }
}
__device__ __forceinline__ uint32_t spread3(uint32_t v) {  // 10 bits -> every third bit
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void mortonKeyKernel(const float* __restrict__ canon, uint32_t n, int D, const int* __restrict__ axes, const float* __restrict__ offScale,
                                uint32_t* __restrict__ keys) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t key = 0;
    for (int r = 0; r < 3; ++r) {
        const int c = axes[r];
        if (c < 0) continue;
        float q = (canon[(size_t)i * D + c] - offScale[2 * r]) * offScale[2 * r + 1];
        q = fminf(fmaxf(q, 0.0f), 255.0f);
        key |= spread3((uint32_t)q) << r;
    }
    keys[i] = key;
}
// SoA leaf boxes of a chunk -> the tail's blocked box array at leaf offset `leaf0`
template <typename S>
__global__ void tailBoxKernel(const S* __restrict__ lo, const S* __restrict__ hi, uint32_t nNew, int D, uint32_t leaf0, S* __restrict__ box) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nNew) return;
    const uint32_t leaf = leaf0 + j, b = leaf >> 5, ln = leaf & 31;
    for (int c = 0; c < D; ++c) {
        box[((size_t)b * 2 * D + c) * 32u + ln] = lo[(size_t)c * nNew + j];
        box[((size_t)b * 2 * D + D + c) * 32u + ln] = hi[(size_t)c * nNew + j];
    }
}

}  // namespace

template <typename S>
int knnTailAppendT(mptg_ctx* ctx, KnnTail& tail, const mptg_space_desc& space, const S* ptsDev, uint32_t stride, uint32_t first, uint32_t count) {
    if (count == 0) return MPTG_OK;
    const DevSpace<float> sp = makeDevSpace<float>(space);
    const int D = sp.D;
    cudaStream_t st = ctx->stream;
    const uint32_t nNew = (count + 31) / 32;
    if (((size_t)tail.nLeaves + nNew) * 32 > TAIL_MAX_POINTS) return MPTG_OK;  // full: the caller scans what is not covered
    if (!tail.mem) {
        const size_t maxLeaves = TAIL_MAX_POINTS / 32, maxBlocks = (maxLeaves + 31) / 32;
        const size_t bPts = ((size_t)TAIL_MAX_POINTS * D * sizeof(S) + 255) & ~(size_t)255, bPerm = ((size_t)TAIL_MAX_POINTS * 4 + 255) & ~(size_t)255,
                     bBox = maxBlocks * 2 * D * 32 * sizeof(S);
        MPTG_CUDA(ctx, cudaMalloc(&tail.mem, bPts + bPerm + bBox));
        tail.leafPts = tail.mem;
        tail.perm = (uint32_t*)((char*)tail.mem + bPts);
        tail.box = (char*)tail.mem + bPts + bPerm;
    }
    size_t cubBytes = 0;
    cub::DoubleBuffer<uint32_t> kb(nullptr, nullptr), vb(nullptr, nullptr);
    MPTG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, cubBytes, kb, vb, (int)count, 0, 24, st));
    size_t wbytes = 0;
    auto wtake = [&](size_t b) {
        const size_t o = wbytes;
        wbytes += (b + 255) & ~(size_t)255;
        return o;
    };
    const size_t wCanon = wtake((size_t)count * D * 4), wExact = wtake(sizeof(S) == 8 ? (size_t)count * D * 8 : 0), wKey0 = wtake((size_t)count * 4),
                 wKey1 = wtake((size_t)count * 4), wId0 = wtake((size_t)count * 4), wId1 = wtake((size_t)count * 4), wSeg = wtake((size_t)count * 4),
                 wMin = wtake((size_t)D * 4), wMax = wtake((size_t)D * 4), wAxes = wtake(16), wOff = wtake(32), wCub = wtake(cubBytes),
                 wLo = wtake((size_t)D * nNew * sizeof(S)), wHi = wtake((size_t)D * nNew * sizeof(S));
    void* wbase;
    if (int rc = scratch(ctx, 7, wbytes, &wbase)) return rc;
    char* W = (char*)wbase;
    float* canon = (float*)(W + wCanon);
    S* exact = sizeof(S) == 8 ? (S*)(W + wExact) : (S*)canon;
    uint32_t* keys[2] = {(uint32_t*)(W + wKey0), (uint32_t*)(W + wKey1)};
    uint32_t* ids[2] = {(uint32_t*)(W + wId0), (uint32_t*)(W + wId1)};
    uint32_t* segOf = (uint32_t*)(W + wSeg);
    int* segMin = (int*)(W + wMin);
    int* segMax = (int*)(W + wMax);
    const uint32_t g256 = (count + 255) / 256;
    canonKernel<S><<<g256, 256, 0, st>>>(sp, ptsDev + first, stride, count, canon, exact);
    MPTG_LAUNCHED(ctx);
    iotaKernel<<<g256, 256, 0, st>>>(ids[0], segOf, count);
    MPTG_LAUNCHED(ctx);
    resetExtentKernel<<<1, 256, 0, st>>>(segMin, segMax, (uint32_t)D);
    MPTG_LAUNCHED(ctx);
    segExtentKernel<<<(count + EXT_THREADS - 1) / EXT_THREADS, EXT_THREADS, 0, st>>>(canon, ids[0], segOf, count, D, segMin, segMax);
    MPTG_LAUNCHED(ctx);
    mortonAxesKernel<<<1, 32, 0, st>>>(sp, segMin, segMax, (int*)(W + wAxes), (float*)(W + wOff));
    MPTG_LAUNCHED(ctx);
    mortonKeyKernel<<<g256, 256, 0, st>>>(canon, count, D, (const int*)(W + wAxes), (const float*)(W + wOff), keys[0]);
    MPTG_LAUNCHED(ctx);
    cub::DoubleBuffer<uint32_t> dk(keys[0], keys[1]), dv(ids[0], ids[1]);
    MPTG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(W + wCub, cubBytes, dk, dv, (int)count, 0, 24, st));
    ++ctx->launches;
    leafEmitKernel<S><<<(nNew * 32 + 255) / 256, 256, 0, st>>>(exact, dv.Current(), count, D, nNew, (S*)tail.leafPts + (size_t)tail.nLeaves * D * 32,
                                                              tail.perm + (size_t)tail.nLeaves * 32, (S*)(W + wLo), (S*)(W + wHi), first);
    MPTG_LAUNCHED(ctx);
    tailBoxKernel<S><<<(nNew + 127) / 128, 128, 0, st>>>((const S*)(W + wLo), (const S*)(W + wHi), nNew, D, tail.nLeaves, (S*)tail.box);
    MPTG_LAUNCHED(ctx);
    tail.nLeaves += nNew;
    tail.covered += count;
    return MPTG_OK;
}

int knnTailAppend(mptg_ctx* ctx, KnnTail& tail, const mptg_space_desc& space, const float* ptsDev, uint32_t stride, uint32_t first, uint32_t count) {
    return knnTailAppendT<float>(ctx, tail, space, ptsDev, stride, first, count);
}
int knnTailAppend(mptg_ctx* ctx, KnnTail& tail, const mptg_space_desc& space, const double* ptsDev, uint32_t stride, uint32_t first, uint32_t count) {
    return knnTailAppendT<double>(ctx, tail, space, ptsDev, stride, first, count);
}

namespace {
}  // namespace

template <typename S>
int knnBuildIndexGpuT(mptg_ctx* ctx, KnnIndex& ix, const mptg_space_desc& space, const S* ptsDev, uint32_t stride, uint32_t n) {
    if (n == 0) {
        ix.count = 0;
        return MPTG_OK;
    }
    const DevSpace<float> sp = makeDevSpace<float>(space);
    const int D = sp.D;
    cudaStream_t st = ctx->stream;
    const bool compressed = sizeof(S) == 4 && classifySpace(space) == SHAPE_SE3;  // half copies + caps; coordinates beyond the half range fall back below
    static const float rotSplit = [] {  // MPTG_KNN_ROT_SPLIT: tuning experiments
        const char* e = getenv("MPTG_KNN_ROT_SPLIT");
        const float v = e ? (float)atof(e) : 0.0f;
        return v > 0.0f ? v : 0.75f;
    }();

    // ---- geometry of the image
    KnnIndex nx;
    nx.count = n;
    nx.nNodes[0] = (n + 31) / 32;
    nx.nPad = nx.nNodes[0] * 32;
    nx.top = 0;
    while (nx.nNodes[nx.top] > 32) {
        if (nx.top + 1 >= BVH_MAXL) return fail(ctx, MPTG_ERR_CAPACITY, "kNN index: too many points for %d levels", BVH_MAXL);
        nx.nNodes[nx.top + 1] = (nx.nNodes[nx.top] + 31) / 32;
        ++nx.top;
    }
    size_t bytes = 0;
    auto take = [&](size_t b) {
        const size_t o = bytes;
        bytes += (b + 255) & ~(size_t)255;
        return o;
    };
    const size_t oPts = take((size_t)D * nx.nPad * sizeof(S));
    const size_t oPerm = take((size_t)nx.nPad * sizeof(uint32_t));
    size_t oBox[BVH_MAXL] = {0, 0, 0, 0, 0}, oCap[BVH_MAXL] = {0, 0, 0, 0, 0}, oLeafH = 0;
    uint32_t nBlocks[BVH_MAXL] = {0, 0, 0, 0, 0};
    for (int l = 0; l <= nx.top; ++l) {
        nBlocks[l] = (nx.nNodes[l] + 31) / 32;
        oBox[l] = take((size_t)nBlocks[l] * 2 * D * 32 * sizeof(S));
    }
    if (compressed) {
        oLeafH = take((size_t)nx.nNodes[0] * 4 * 32 * sizeof(uint32_t));
        for (int l = 0; l <= nx.top; ++l) oCap[l] = take((size_t)nBlocks[l] * 3 * 32 * sizeof(float4));
    }
    void* mem = ix.mem;
    size_t memBytes = ix.memBytes;
    if (memBytes < bytes) {
        MPTG_CUDA(ctx, cudaStreamSynchronize(st));
        if (mem) {
            // the handle must not keep a pointer to freed memory if the allocation below fails
            ix.mem = nullptr, ix.memBytes = 0, ix.count = 0, ix.leafH = nullptr;
            MPTG_CUDA(ctx, cudaFree(mem));
        }
        mem = nullptr;
        memBytes = bytes + bytes / 2;
        if (ix.capacityHint > n) {  // every term of `bytes` is linear in n up to padding
            double grow = (double)ix.capacityHint / (double)n;
            if (grow > 8.0) grow = 8.0;  // a nearly empty store of huge capacity grows in a few steps instead
            const size_t atCapacity = (size_t)((double)bytes * grow * 1.02) + (1u << 20);
            if (atCapacity > memBytes) memBytes = atCapacity;
        }
        MPTG_CUDA(ctx, cudaMalloc(&mem, memBytes));
    }
    unsigned long long* stats = ix.devStats;
    if (!stats) {
        MPTG_CUDA(ctx, cudaMalloc(&stats, 8 * sizeof(unsigned long long)));
        ix.devStats = stats;  // owned by the handle from here on: freed with it whatever happens below
        if (int rc = memsetSync(ctx, stats, 0, 8 * sizeof(unsigned long long))) return rc;
    }

    // ---- work space (context scratch slot 7): canon, keys x2, ids x2, segOf, extents, axis, level tables, SoA boxes, cub temp
    const Levels L = makeLevels(n);
    const uint32_t maxSeg = nx.nNodes[0] + 1;
    const size_t nTab = L.begin.size();
    size_t cubBytes = 0;
    cub::DoubleBuffer<unsigned long long> kb(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> vb(nullptr, nullptr);
    MPTG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(nullptr, cubBytes, kb, vb, (int)n, 0, 64, st));
    size_t wbytes = 0;
    auto wtake = [&](size_t b) {
        const size_t o = wbytes;
        wbytes += (b + 255) & ~(size_t)255;
        return o;
    };
    const size_t wCanon = wtake((size_t)n * D * 4), wExact = wtake(sizeof(S) == 8 ? (size_t)n * D * 8 : 0), wKey0 = wtake((size_t)n * 8), wKey1 = wtake((size_t)n * 8), wId0 = wtake((size_t)n * 4),
                 wId1 = wtake((size_t)n * 4), wSeg = wtake((size_t)n * 4), wMin = wtake((size_t)maxSeg * D * 4),
                 wMax = wtake((size_t)maxSeg * D * 4), wAxis = wtake((size_t)maxSeg * 4), wTab = wtake(nTab * 4 * 4), wCub = wtake(cubBytes),
                 wErr = wtake(16);
    size_t wLo[BVH_MAXL], wHi[BVH_MAXL], wCentre[BVH_MAXL] = {0, 0, 0, 0, 0}, wMinCos[BVH_MAXL] = {0, 0, 0, 0, 0};
    for (int l = 0; l <= nx.top; ++l) wLo[l] = wtake((size_t)D * nx.nNodes[l] * sizeof(S)), wHi[l] = wtake((size_t)D * nx.nNodes[l] * sizeof(S));
    if (compressed)
        for (int l = 0; l <= nx.top; ++l) wCentre[l] = wtake((size_t)nx.nNodes[l] * sizeof(float4)), wMinCos[l] = wtake((size_t)nx.nNodes[l] * 4);
    void* wbase;
    {
        size_t want = wbytes;  // sized for the store's capacity the first time (see KnnIndex::capacityHint)
        if (ix.capacityHint > n) {
            double grow = (double)ix.capacityHint / (double)n;
            if (grow > 8.0) grow = 8.0;
            want = (size_t)((double)wbytes * grow * 1.02) + (1u << 20);
        }
        if (ctx->scratchBytes[7] >= wbytes) want = wbytes;
        if (int rc = scratch(ctx, 7, want, &wbase)) return rc;
    }
    char* W = (char*)wbase;
    float* canon = (float*)(W + wCanon);
    S* exact = sizeof(S) == 8 ? (S*)(W + wExact) : (S*)canon;  // what the emitted leaves and boxes are made of
    unsigned long long* keys[2] = {(unsigned long long*)(W + wKey0), (unsigned long long*)(W + wKey1)};
    uint32_t* ids[2] = {(uint32_t*)(W + wId0), (uint32_t*)(W + wId1)};
    uint32_t* segOf = (uint32_t*)(W + wSeg);
    int* segMin = (int*)(W + wMin);
    int* segMax = (int*)(W + wMax);
    uint32_t* axis = (uint32_t*)(W + wAxis);
    uint32_t* tab = (uint32_t*)(W + wTab);
    unsigned int* err = (unsigned int*)(W + wErr);
    {  // level tables: begin | mid | end | childBase, each nTab entries
        std::vector<uint32_t> host(nTab * 4);
        std::copy(L.begin.begin(), L.begin.end(), host.begin());
        std::copy(L.mid.begin(), L.mid.end(), host.begin() + nTab);
        std::copy(L.end.begin(), L.end.end(), host.begin() + 2 * nTab);
        std::copy(L.childBase.begin(), L.childBase.end(), host.begin() + 3 * nTab);
        if (int rc = uploadSync(ctx, tab, host.data(), host.size() * 4)) return rc;
    }

    // ---- 1. canonical copy, 2. one sort per binary level
    const uint32_t g256 = (n + 255) / 256;
    canonKernel<S><<<g256, 256, 0, st>>>(sp, ptsDev, stride, n, canon, exact);
    MPTG_LAUNCHED(ctx);
    iotaKernel<<<g256, 256, 0, st>>>(ids[0], segOf, n);
    MPTG_LAUNCHED(ctx);
    int cur = 0;
    const size_t nLevels = L.levelCount.size();
    for (size_t lev = 0; lev + 1 < nLevels; ++lev) {  // the last level has nothing left to split
        const uint32_t nSeg = L.levelCount[lev], off = L.levelOffset[lev];
        SegLevel lv{tab + off, tab + nTab + off, tab + 2 * nTab + off, tab + 3 * nTab + off, nSeg};
        resetExtentKernel<<<(nSeg * D + 255) / 256, 256, 0, st>>>(segMin, segMax, nSeg * D);
        MPTG_LAUNCHED(ctx);
        segExtentKernel<<<(n + EXT_THREADS - 1) / EXT_THREADS, EXT_THREADS, 0, st>>>(canon, ids[cur], segOf, n, D, segMin, segMax);
        MPTG_LAUNCHED(ctx);
        axisKernel<<<(nSeg + 255) / 256, 256, 0, st>>>(sp, segMin, segMax, nSeg, axis, rotSplit);
        MPTG_LAUNCHED(ctx);
        keyKernel<<<g256, 256, 0, st>>>(canon, ids[cur], segOf, axis, n, D, keys[cur]);
        MPTG_LAUNCHED(ctx);
        cub::DoubleBuffer<unsigned long long> dk(keys[cur], keys[cur ^ 1]);
        cub::DoubleBuffer<uint32_t> dv(ids[cur], ids[cur ^ 1]);
        int segBits = 1;
        while ((1u << segBits) < nSeg) ++segBits;
        MPTG_CUDA(ctx, cub::DeviceRadixSort::SortPairs(W + wCub, cubBytes, dk, dv, (int)n, 0, 32 + segBits, st));
        ++ctx->launches;
        cur = dv.Current() == ids[0] ? 0 : 1;  // the key buffers are scratch: fresh keys are written every level
        splitKernel<<<g256, 256, 0, st>>>(lv, segOf, n);
        MPTG_LAUNCHED(ctx);
    }

    // ---- 3. device image
    char* M = (char*)mem;
    S* lo0 = (S*)(W + wLo[0]);
    S* hi0 = (S*)(W + wHi[0]);
    leafEmitKernel<S><<<(nx.nNodes[0] * 32 + 255) / 256, 256, 0, st>>>(exact, ids[cur], n, D, nx.nNodes[0], (S*)(M + oPts),
                                                                      (uint32_t*)(M + oPerm), lo0, hi0);
    MPTG_LAUNCHED(ctx);
    for (int l = 1; l <= nx.top; ++l) {
        parentBoxKernel<S><<<(nx.nNodes[l] * 32 + 255) / 256, 256, 0, st>>>((const S*)(W + wLo[l - 1]), (const S*)(W + wHi[l - 1]),
                                                                           nx.nNodes[l - 1], D, nx.nNodes[l], (S*)(W + wLo[l]), (S*)(W + wHi[l]));
        MPTG_LAUNCHED(ctx);
    }
    for (int l = 0; l <= nx.top; ++l) {
        boxBlockKernel<S><<<(nBlocks[l] * 32 + 255) / 256, 256, 0, st>>>((const S*)(W + wLo[l]), (const S*)(W + wHi[l]), nx.nNodes[l], D,
                                                                        (S*)(M + oBox[l]));
        MPTG_LAUNCHED(ctx);
    }
    float errQ = 0.f, errT = 0.f, normMax = 1.f, tScale = 1.f;
    bool useHalf = compressed;
    if (compressed) {
        MPTG_CUDA(ctx, cudaMemsetAsync(err, 0, 16, st));
        tScaleKernel<<<1, 32, 0, st>>>((const float*)(W + wLo[nx.top]), (const float*)(W + wHi[nx.top]), nx.nNodes[nx.top], (float*)(err + 3));
        MPTG_LAUNCHED(ctx);
        leafHalfKernel<<<(nx.nPad + 255) / 256, 256, 0, st>>>((const float*)(M + oPts), nx.nPad, (uint32_t*)(M + oLeafH), err);
        MPTG_LAUNCHED(ctx);
        // rotation caps of every level: centres bottom-up, radii against ALL member points, then the blocked image
        const float* lp = (const float*)(M + oPts);
        for (int l = 0; l <= nx.top; ++l) {
            float4* centre = (float4*)(W + wCentre[l]);
            unsigned int* minCos = (unsigned int*)(W + wMinCos[l]);
            if (l == 0) leafCapCentreKernel<<<(nx.nNodes[0] * 32 + 255) / 256, 256, 0, st>>>(lp, nx.nNodes[0], centre);
            else parentCapCentreKernel<<<(nx.nNodes[l] * 32 + 255) / 256, 256, 0, st>>>((const float4*)(W + wCentre[l - 1]), nx.nNodes[l - 1], nx.nNodes[l], centre);
            MPTG_LAUNCHED(ctx);
            fillKernel<<<(nx.nNodes[l] + 255) / 256, 256, 0, st>>>(minCos, 0x3f800000u, nx.nNodes[l]);
            MPTG_LAUNCHED(ctx);
            capRadiusKernel<<<(nx.nPad + 255) / 256, 256, 0, st>>>(lp, nx.nPad, 5 * (l + 1), centre, minCos, l == 0 ? err + 2 : nullptr);
            MPTG_LAUNCHED(ctx);
            capPackKernel<<<(nBlocks[l] * 32 + 255) / 256, 256, 0, st>>>(centre, minCos, (const float*)(W + wLo[l]), (const float*)(W + wHi[l]), nx.nNodes[l],
                                                                        (float4*)(M + oCap[l]));
            MPTG_LAUNCHED(ctx);
        }
        unsigned int herr[4];
        MPTG_CUDA(ctx, cudaMemcpyAsync(herr, err, 16, cudaMemcpyDeviceToHost, st));
        MPTG_CUDA(ctx, cudaStreamSynchronize(st));
        memcpy(&errQ, &herr[0], 4);
        memcpy(&errT, &herr[1], 4);
        memcpy(&normMax, &herr[2], 4);
        memcpy(&tScale, &herr[3], 4);
        // a coordinate overflowed the half range, or a quaternion is not finite: keep the float box path
        if (!(errT < 1e30f) || !(errQ < 1e30f) || !(normMax < 1e18f)) useHalf = false;
    }

    nx.mem = mem;
    nx.memBytes = memBytes;
    nx.devStats = stats;
    nx.builds = ix.builds + 1;
    nx.capacityHint = ix.capacityHint;
    nx.leafPts = M + oPts;
    nx.perm = (uint32_t*)(M + oPerm);
    for (int l = 0; l <= nx.top; ++l) nx.box[l] = M + oBox[l];
    if (useHalf) {
        nx.leafH = (uint32_t*)(M + oLeafH);
        for (int l = 0; l <= nx.top; ++l) nx.cap[l] = M + oCap[l];
        nx.errQ = errQ * 1.0001f + 1e-30f;
        nx.errT = errT * 1.0001f + 1e-30f;
        nx.normMax = normMax;
        nx.tScale = tScale;
    }
    ix = nx;
    return MPTG_OK;
}

int knnBuildIndexGpu(mptg_ctx* ctx, KnnIndex& ix, const mptg_space_desc& space, const float* ptsDev, uint32_t stride, uint32_t n) {
    return knnBuildIndexGpuT<float>(ctx, ix, space, ptsDev, stride, n);
}
int knnBuildIndexGpu(mptg_ctx* ctx, KnnIndex& ix, const mptg_space_desc& space, const double* ptsDev, uint32_t stride, uint32_t n) {
    return knnBuildIndexGpuT<double>(ctx, ix, space, ptsDev, stride, n);
}

}  // namespace mptg
