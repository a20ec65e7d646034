// fpmath_check.cpp -- accuracy and monotonicity of include/mptg/mptg_fpmath.h against libm.
// TEST INFRASTRUCTURE ONLY.  Usage: fpmath_check [stride]   (stride 1 = every float in [0,1])
// Prints: max ulp error of acos01(float) vs (float)acos((double)x), number of monotonicity
// violations, max relative error of the double acos / sin / cos on dense samples.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../include/mptg/mptg_fpmath.h"

namespace fp = mptg::fp;

static float bitsToFloat(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static int32_t ulpDiff(float a, float b) {
    int32_t ia, ib;
    memcpy(&ia, &a, 4);
    memcpy(&ib, &b, 4);
    return ia > ib ? ia - ib : ib - ia;
}

int main(int argc, char** argv) {
    uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;
    // floats in [2^-30, 1]: bit patterns are monotone for positive floats
    uint32_t lo, hi;
    float flo = 9.3132257e-10f, fhi = 1.0f;
    memcpy(&lo, &flo, 4);
    memcpy(&hi, &fhi, 4);
    int32_t maxUlp = 0;
    uint64_t nonMono = 0, count = 0;
    float prev = fp::acos01(bitsToFloat(lo));
    for (uint64_t u = lo; u <= hi; u += stride) {
        float x = bitsToFloat((uint32_t)u);
        float got = fp::acos01(x);
        float want = (float)std::acos((double)x);
        int32_t d = ulpDiff(got, want);
        if (want > 1e-3f && d > maxUlp) maxUlp = d;  // ulp metric is meaningless right at acos(1)=0
        if (got > prev) ++nonMono;                   // acos is decreasing
        prev = got;
        ++count;
    }
    // absolute error near x -> 1
    double maxAbsNear1 = 0;
    for (uint32_t u = hi - 4096; u <= hi; ++u) {
        float x = bitsToFloat(u);
        maxAbsNear1 = std::fmax(maxAbsNear1, std::fabs((double)fp::acos01(x) - std::acos((double)x)));
    }
    std::printf("acos01f: %llu samples, max ulp err %d, monotonicity violations %llu, max abs err near 1: %.3g\n",
                (unsigned long long)count, maxUlp, (unsigned long long)nonMono, maxAbsNear1);

    double maxRelAcos = 0, maxRelSin = 0, maxRelCos = 0;
    for (int i = 0; i <= 2000000; ++i) {
        double x = i / 2000000.0;
        double a = fp::acos01(x), b = std::acos(x);
        if (b > 0) maxRelAcos = std::fmax(maxRelAcos, std::fabs(a - b) / b);
    }
    for (int i = -2000000; i <= 2000000; ++i) {
        double x = i * (110.0 / 2000000.0);  // |x| <= 110 covers 32 links * pi
        double s, c;
        fp::sincos_(x, &s, &c);
        double es = std::fabs(s - std::sin(x)), ec = std::fabs(c - std::cos(x));
        maxRelSin = std::fmax(maxRelSin, es);
        maxRelCos = std::fmax(maxRelCos, ec);
    }
    std::printf("acos01d: max rel err %.3g; sincos_d on [-110,110]: max abs err sin %.3g cos %.3g\n", maxRelAcos,
                maxRelSin, maxRelCos);
    bool ok = maxUlp <= 4 && maxRelAcos < 1e-15 && maxRelSin < 5e-16 && maxRelCos < 5e-16;
    std::printf("%s\n", ok ? "OK" : "FAILED");
    return ok ? 0 : 1;
}
