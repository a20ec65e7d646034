// ref_nao.cpp -- the REFERENCE'S OWN Nao-cup scenario code (demo/nao_cup/src/{naocup,collide,linear}.hpp), compiled from
// where it lies under /root/reference, behind a small C interface.  TEST INFRASTRUCTURE ONLY (built into oracle/_ref/).
//
// Nothing is copied.  Its one external dependency, Eigen (Transform / AngleAxis / Translation / small vectors), is
// absent here and is satisfied by the stand-ins of OURS under oracle/shim/Eigen/{Dense,Geometry}, whose product
// structure follows Eigen's and whose sums run left to right (stated in that header).  So nao_clear() and nao_link()
// -- forward kinematics of both arms, the 209 sphere / capsule pair tests, the cup-upright condition and the 1-degree
// midpoint bisection -- execute exactly as the reference wrote them, with libm sin/cos, -ffp-contract=off.
#include <cstdint>

#include <nao_cup/src/naocup.hpp>

namespace {
template <typename S>
int clearBatch(const S* q, uint32_t n, uint8_t* ok, uint8_t* collision) {
    void* w = nao_cup::nao_system_data_alloc<S>(0, nullptr, nullptr);
    for (uint32_t i = 0; i < n; ++i) {
        ok[i] = nao_cup::nao_clear<S>(w, q + (size_t)i * nao_cup::DIMENSIONS) ? 1 : 0;
        auto* world = static_cast<nao_cup::nao_world<S>*>(w);
        if (collision) collision[i] = world->in_collision ? 1 : 0;
    }
    nao_cup::nao_system_data_free<S>(w);
    return 0;
}
template <typename S>
int linkBatch(const S* a, const S* b, uint32_t n, uint8_t* ok) {
    void* w = nao_cup::nao_system_data_alloc<S>(0, nullptr, nullptr);
    for (uint32_t i = 0; i < n; ++i)
        ok[i] = nao_cup::nao_link<S>(w, a + (size_t)i * nao_cup::DIMENSIONS, b + (size_t)i * nao_cup::DIMENSIONS) ? 1 : 0;
    nao_cup::nao_system_data_free<S>(w);
    return 0;
}
}  // namespace

extern "C" {
// scalar: 4 = float, 8 = double.  collision (optional) = the world's in_collision flag of each state.
int ref_nao_clear(int scalar, const void* q, uint32_t n, uint8_t* ok, uint8_t* collision) {
    return scalar == 4 ? clearBatch<float>((const float*)q, n, ok, collision)
                       : clearBatch<double>((const double*)q, n, ok, collision);
}
int ref_nao_link(int scalar, const void* a, const void* b, uint32_t n, uint8_t* ok) {
    return scalar == 4 ? linkBatch<float>((const float*)a, (const float*)b, n, ok) : linkBatch<double>((const double*)a, (const double*)b, n, ok);
}
// the reference's configuration constants (start, goal, bounds), 10 scalars each as doubles
void ref_nao_configs(int scalar, double* init, double* lo, double* hi, double* target) {
    for (unsigned i = 0; i < nao_cup::DIMENSIONS; ++i) {
        if (scalar == 4) {
            init[i] = nao_cup::nao_init_config<float>()[i], lo[i] = nao_cup::nao_min_config<float>()[i];
            hi[i] = nao_cup::nao_max_config<float>()[i], target[i] = nao_cup::nao_target_config<float>()[i];
        } else {
            init[i] = nao_cup::nao_init_config<double>()[i], lo[i] = nao_cup::nao_min_config<double>()[i];
            hi[i] = nao_cup::nao_max_config<double>()[i], target[i] = nao_cup::nao_target_config<double>()[i];
        }
    }
}
}
