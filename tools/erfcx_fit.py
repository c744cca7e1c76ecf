"""Coefficients of the pair kernel's erfc: erfc(x) = exp(-x^2) * t * P(t), t = 1/(1 + p x), x in [0, 6].

Iteratively reweighted least squares on the RELATIVE error of erfcx(x)/t (a minimax-like fit), degree 7, p = 0.3275911:
max relative error 5.4e-8 in exact arithmetic, ~2.5e-7 when evaluated with FP32 Horner steps (rounding dominated).
One MUFU.RCP + 7 FFMA + 1 FMUL, and the exp(-x^2) factor is the one the force expression needs anyway.

    python tools/erfcx_fit.py        # prints the C initialiser used in csrc/direct.cu
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
from scipy.special import erfcx

P_COEF, DEGREE, XMAX = 0.3275911, 7, 6.0


def fit():
    x = np.linspace(0, XMAX, 400001)
    t = 1 / (1 + P_COEF * x)
    f = erfcx(x) / t
    tmin, tmax = t.min(), t.max()
    u = (2 * t - (tmin + tmax)) / (tmax - tmin)
    w = np.ones_like(x)
    for _ in range(60):
        c = C.chebfit(u, f, DEGREE, w=w / f)
        err = np.abs(C.chebval(u, c) / f - 1)
        w = w * (1 + 5 * err / err.max())
    a, b = 2 / (tmax - tmin), -(tmin + tmax) / (tmax - tmin)
    pt, base = np.zeros(1), np.array([1.0])
    for ck in C.cheb2poly(c):
        pt = P.polyadd(pt, ck * base)
        base = P.polymul(base, np.array([b, a]))
    t32 = t.astype(np.float32)
    acc = np.float32(pt[-1]) * np.ones_like(t32)
    for ck in pt[-2::-1]:
        acc = acc * t32 + np.float32(ck)
    return pt, err.max(), np.abs(acc.astype(np.float64) / f - 1).max()


if __name__ == "__main__":
    pt, e64, e32 = fit()
    print("// tools/erfcx_fit.py: p = %.7f, degree %d, x in [0, %g]: max rel err %.2e (exact), %.2e (FP32 Horner)" % (P_COEF, DEGREE, XMAX, e64, e32))
    print("// P(t) = c0 + c1 t + ... + c7 t^7")
    print(", ".join("%.9ef" % v for v in pt))
