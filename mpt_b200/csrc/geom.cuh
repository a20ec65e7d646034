// geom.cuh -- scenario geometry handle shared by geom.cu (grid, shapes, link arm) and mesh.cu.
#pragma once

#include "common.cuh"
#include "../../include/mptg/mptg_space.h"

struct MeshData;  // mesh.cu

struct mptg_geom {
    mptg_ctx* ctx = nullptr;
    int kind = 0;
    int scalar = MPTG_F32;
    int D = 0;  // scalars per state
    // grid
    int width = 0, height = 0;
    uint32_t* gridBits = nullptr;  // 1 bit per cell, row-major linear index, bit = obstacle
    // shapes
    int dim = 0, nBalls = 0, nRects = 0;
    void* balls = nullptr;  // nBalls * (dim + 1) scalars: centre..., radius
    void* rects = nullptr;  // nRects * 4 scalars
    // link arm
    int nLinks = 0, nCircles = 0;
    void* lengths = nullptr;  // nLinks scalars
    void* circles = nullptr;  // nCircles * 3 scalars (cx, cy, r)
    double linkRadius = 0;
    // mesh
    MeshData* mesh = nullptr;
    // Nao-cup scenario: nao::Model<float|double> (host; passed to the kernels by value)
    void* naoModel = nullptr;
    // flat edge check (geom.cu flatLink): per-edge item counts, their prefix sums, scan work space
    void* flatBuf = nullptr;
    size_t flatBytes = 0;
    // per-call status / counters (device), mirrored on demand
    unsigned long long* devStats = nullptr;  // [0]=states [1]=bv tests [2]=primitive tests [3]=items [4]=error flags
    uint64_t stats[4] = {0, 0, 0, 0};
};

namespace mptg {
constexpr unsigned FULL_MASK_ = 0xffffffffu;
enum GeomError : unsigned long long { GEOM_ERR_STACK = 1ull, GEOM_ERR_STEPS = 2ull, GEOM_ERR_SCHED = 4ull };

// implemented in mesh.cu
int meshCreate(mptg_ctx* ctx, int scalar, uint32_t nr, const float* robotTris, uint32_t ne, const float* envTris, MeshData** out);
void meshDestroy(MeshData* m);
int meshValidDev(mptg_geom* g, const void* statesDev, uint32_t n, uint8_t* okDev, uint8_t* nearDev);
int meshLinkDev(mptg_geom* g, const mptg_space_desc* space, const void* fromDev, const void* toDev, uint32_t n, double step,
                uint8_t* okDev, uint8_t* nearDev);
double meshBand(const MeshData* m);
}  // namespace mptg
