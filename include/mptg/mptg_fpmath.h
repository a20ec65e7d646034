/* mptg_fpmath.h -- transcendental primitives with a WRITTEN operation order.
 *
 * Why this exists: the reference (UNC-Robotics/mpt) calls std::acos / std::sin / std::cos from the
 * host libm (src/mpt/so3_space.hpp:61-66, demo/link_manipulator_scenario.hpp:107-108; the SO(3)
 * distance acos lives in Nigh, pinned by test/so3_space_test.cpp:53-55).  glibc's and CUDA's
 * implementations differ in the last ulp, so "bit-exact CPU oracle vs sm_100a kernel" is only
 * possible if both sides evaluate the SAME sequence of IEEE-754 basic operations.  Every function
 * here uses only + - * / sqrt fma (all correctly rounded on x86-64 and on sm_100a) in a fixed order,
 * and is compiled by g++ (oracle, host layer) and by nvcc (kernels).  Build flags that keep the
 * compilers from re-associating or contracting: g++ -ffp-contract=off, nvcc --fmad=false
 * (explicit fma calls below are kept; implicit contraction is what is disabled).
 *
 * Algorithms: double acos / sin / cos restate the published fdlibm algorithms (e_acos.c, k_sin.c,
 * k_cos.c, medium-argument path of e_rem_pio2.c; Sun Microsystems, freely redistributable) --
 * coefficients verified against mpmath to < 1e-17 relative.  float acos on [0,1] is our own fit,
 * acos(x) = sqrt(1-x) * P8(x), derived by tools/fit_fpmath.py (max real-valued rel. error 3.9e-8).
 * Accuracy is checked exhaustively / by dense sampling against libm in oracle/fpmath_check.cpp.
 */
#ifndef MPTG_FPMATH_H
#define MPTG_FPMATH_H

#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define MPTG_HD __host__ __device__ __forceinline__
#else
#define MPTG_HD inline
#endif

namespace mptg {
namespace fp {

/* ---- correctly rounded basic ops, immune to fast-math style flags on the device ---- */
MPTG_HD float fma_(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}
MPTG_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
MPTG_HD float sqrt_(float x) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return __builtin_sqrtf(x);
#endif
}
MPTG_HD double sqrt_(double x) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(x);
#else
    return __builtin_sqrt(x);
#endif
}
MPTG_HD float div_(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
MPTG_HD double div_(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
MPTG_HD float abs_(float x) { return fabsf(x); }
MPTG_HD double abs_(double x) { return fabs(x); }

/* clear the low 32 bits of a double (fdlibm SET_LOW_WORD(x,0)) */
MPTG_HD double clear_low32(double x) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(x), 0);
#else
    uint64_t u;
    memcpy(&u, &x, 8);
    u &= 0xffffffff00000000ull;
    memcpy(&x, &u, 8);
    return x;
#endif
}

/* round-to-nearest-even integer valued double */
MPTG_HD double rint_(double x) {
#if defined(__CUDA_ARCH__)
    return rint(x);
#else
    return __builtin_rint(x); /* default rounding mode = nearest-even */
#endif
}

/* ---- acos on [0,1], double (fdlibm e_acos.c, x >= 0 branches) ---- */
MPTG_HD double acos01(double x) {
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
    const double pS0 = 1.66666666666666657415e-01, pS1 = -3.25565818622400915405e-01,
                 pS2 = 2.01212532134862925881e-01, pS3 = -4.00555345006794114027e-02,
                 pS4 = 7.91534994289814532176e-04, pS5 = 3.47933107596021167570e-05,
                 qS1 = -2.40339491173441421878e+00, qS2 = 2.02094576023350569471e+00,
                 qS3 = -6.88283971605453293030e-01, qS4 = 7.70381505559019352791e-02;
    if (x >= 1.0) return 0.0;
    if (x < 0.5) {
        if (x < 6.938893903907228e-18) return pio2_hi + pio2_lo; /* |x| < 2^-57 */
        double z = x * x;
        double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        double r = div_(p, q);
        return pio2_hi - (x - (pio2_lo - x * r));
    }
    double z = (1.0 - x) * 0.5;
    double s = sqrt_(z);
    double df = clear_low32(s);
    double c = div_(z - df * df, s + df);
    double p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    double q = 1.0 + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    double r = div_(p, q);
    double w = r * s + c;
    return 2.0 * (df + w);
}

/* ---- acos on [0,1], float: sqrt(1-x) * P8(x), Horner with explicit fma ---- */
MPTG_HD float acos01(float x) {
    float t = 1.0f - x;
    float s = sqrt_(t);
    float p = 6.845318130e-04f;
    p = fma_(p, x, -3.974577878e-03f);
    p = fma_(p, x, 1.102838106e-02f);
    p = fma_(p, x, -2.072766609e-02f);
    p = fma_(p, x, 3.257117048e-02f);
    p = fma_(p, x, -5.059357360e-02f);
    p = fma_(p, x, 8.903013915e-02f);
    p = fma_(p, x, -2.146011591e-01f);
    p = fma_(p, x, 1.570796371e+00f);
    return s * p;
}

/* ---- sin & cos, double, |x| < 2^20 (fdlibm medium path, always two Cody-Waite iterations) ---- */
MPTG_HD void sincos_(double x, double* sn, double* cs) {
    const double invpio2 = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00;  /* first 33 bits of pi/2 */
    const double pio2_2 = 6.07710050630396597660e-11;  /* second 33 bits */
    const double pio2_2t = 2.02226624879595063154e-21; /* pi/2 - (pio2_1+pio2_2) */
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double fn = rint_(x * invpio2);
    double t = x - fn * pio2_1;
    double w = fn * pio2_2;
    double r = t - w;
    w = fn * pio2_2t - ((t - r) - w);
    double y0 = r - w;
    double y1 = (r - y0) - w;
    /* kernels on [-pi/4, pi/4] with tail y1 */
    double z = y0 * y0;
    double v = z * y0;
    double rs = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
    double ks = y0 - ((z * (0.5 * y1 - v * rs) - y1) - v * S1);
    double rc = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
    double hz = 0.5 * z;
    double wc = 1.0 - hz;
    double kc = wc + (((1.0 - wc) - hz) + (z * rc - y0 * y1));
    long long n = (long long)fn;
    switch (n & 3) {
        case 0: *sn = ks; *cs = kc; break;
        case 1: *sn = kc; *cs = -ks; break;
        case 2: *sn = -ks; *cs = -kc; break;
        default: *sn = -kc; *cs = ks; break;
    }
}
MPTG_HD double sin_(double x) { double s, c; sincos_(x, &s, &c); return s; }
MPTG_HD double cos_(double x) { double s, c; sincos_(x, &s, &c); return c; }

/* float sin/cos: evaluate in double, round once.  (float)double is round-to-nearest on both sides. */
MPTG_HD void sincos_(float x, float* sn, float* cs) {
    double s, c;
    sincos_((double)x, &s, &c);
    *sn = (float)s;
    *cs = (float)c;
}
MPTG_HD float sin_(float x) { return (float)sin_((double)x); }
MPTG_HD float cos_(float x) { return (float)cos_((double)x); }

template <typename S> struct consts;
template <> struct consts<float> {
    static MPTG_HD float pi() { return 3.14159274101257324219f; }
    static MPTG_HD float eps() { return 1.1920928955078125e-07f; }
    static MPTG_HD float inf() { return __builtin_huge_valf(); }
};
template <> struct consts<double> {
    static MPTG_HD double pi() { return 3.14159265358979323846; }
    static MPTG_HD double eps() { return 2.220446049250313e-16; }
    static MPTG_HD double inf() { return __builtin_huge_val(); }
};

}  // namespace fp
}  // namespace mptg

#endif /* MPTG_FPMATH_H */
