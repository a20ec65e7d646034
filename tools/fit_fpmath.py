#!/usr/bin/env python3
"""Derive the float32 polynomial used by mptg_acos01f (include/mptg/mptg_fpmath.h).

acos(x) = sqrt(1-x) * P(x) on [0,1]; g(x) = acos(x)/sqrt(1-x) is analytic on [0,1], so a Chebyshev
interpolant of modest degree is near-minimax.  Prints the coefficients (as float32 hex-exact decimals)
and the max abs/rel error of the real-valued approximation.  The exhaustive float32 ulp check against
libm lives in oracle/fpmath_check.cpp.
"""
import numpy as np
import mpmath as mp

mp.mp.dps = 40
DEG = 8


def g(x):
    x = mp.mpf(x)
    if x == 1:
        return mp.sqrt(2)
    return mp.acos(x) / mp.sqrt(1 - x)


def main():
    n = DEG + 1
    # Chebyshev nodes on [0,1]
    k = np.arange(n)
    nodes = 0.5 + 0.5 * np.cos((2 * k + 1) * np.pi / (2 * n))
    vals = [g(float(t)) for t in nodes]
    # solve Vandermonde in high precision
    A = mp.matrix(n, n)
    for i, t in enumerate(nodes):
        for j in range(n):
            A[i, j] = mp.mpf(float(t)) ** j
    c = mp.lu_solve(A, mp.matrix(vals))
    coef = [float(np.float32(float(ci))) for ci in c]
    xs = np.linspace(0, 1, 20001)
    err = 0
    rel = 0
    for x in xs:
        p = sum(mp.mpf(cj) * mp.mpf(float(x)) ** j for j, cj in enumerate(coef))
        a = mp.sqrt(1 - mp.mpf(float(x))) * p
        e = abs(a - mp.acos(float(x)))
        err = max(err, e)
        if x < 1:
            rel = max(rel, e / mp.acos(float(x)))
    print("degree", DEG, "max abs err", mp.nstr(err, 5), "max rel err", mp.nstr(rel, 5))
    for j, cj in enumerate(coef):
        print(f"    c{j} = {np.float32(cj)!r:>16}  // {float(cj).hex()}")
    print("static const float MPTG_ACOSF_C[] = {" + ", ".join(f"{cj:.9e}f" for cj in coef) + "};")


if __name__ == "__main__":
    main()
