"""Fit the fp32 approximations used by csrc/kl.cuh for the exact complex-VD KL

    penalty(la) = gamma - la - Ei(-exp(-la)) = Ein(t),  t = exp(-la)
    (reference: cplxmodule/nn/relevance/complex/vd.py:95-99)

  t <= 1 :  Ein(t) = t * P(t)                       (entire function, no cancellation)
  t >  1 :  Ein(t) = gamma + ln t + E1(t),  E1(t) = exp(-t)/t * H(1/t)

Truth from mpmath at 50 digits.  Prints C arrays (Horner order, highest first).
"""
import numpy as np
import mpmath as mp

mp.mp.dps = 50


def ein(t):
    t = mp.mpf(t)
    return mp.euler + mp.log(t) + mp.e1(t)


def fit_small(deg):
    # Chebyshev nodes on [0,1]; fit P(t) = Ein(t)/t, P(0)=1
    n = 400
    x = 0.5 - 0.5 * np.cos(np.pi * (np.arange(n) + 0.5) / n)
    y = np.array([float(ein(v) / mp.mpf(v)) for v in x])
    c = np.polynomial.chebyshev.Chebyshev.fit(x, y, deg, domain=[0, 1])
    p = c.convert(kind=np.polynomial.Polynomial, domain=[0, 1], window=[0, 1])
    coef = p.coef  # ascending
    xs = np.linspace(1e-6, 1, 5000)
    approx = np.polyval(coef[::-1].astype(np.float32).astype(np.float64), xs)
    truth = np.array([float(ein(v) / mp.mpf(v)) for v in xs])
    return coef, np.max(np.abs(approx / truth - 1))


def fit_large(deg, tmax):
    # H(u) = t e^t E1(t), u = 1/t in [1/tmax, 1]
    n = 600
    lo, hi = 1.0 / tmax, 1.0
    x = 0.5 * (lo + hi) - 0.5 * (hi - lo) * np.cos(np.pi * (np.arange(n) + 0.5) / n)
    y = np.array([float(mp.e1(1 / mp.mpf(u)) * mp.exp(1 / mp.mpf(u)) / mp.mpf(u)) for u in x])
    c = np.polynomial.chebyshev.Chebyshev.fit(x, y, deg, domain=[lo, hi])
    p = c.convert(kind=np.polynomial.Polynomial, domain=[lo, hi], window=[lo, hi])
    coef = p.coef
    us = np.linspace(lo, hi, 5000)
    approx = np.polyval(coef[::-1].astype(np.float32).astype(np.float64), us)
    truth = np.array([float(mp.e1(1 / mp.mpf(u)) * mp.exp(1 / mp.mpf(u)) / mp.mpf(u)) for u in us])
    # what matters is the absolute error of E1 relative to Ein
    t = 1 / us
    e1_err = np.abs(approx - truth) * np.exp(-t) / t
    einv = np.array([float(ein(v)) for v in t])
    return coef, np.max(np.abs(approx / truth - 1)), np.max(e1_err / einv)


if __name__ == "__main__":
    for deg in (6, 7, 8, 9):
        coef, err = fit_small(deg)
        print("small deg", deg, "max rel err", err)
    coef, err = fit_small(8)
    print("static const float EIN_SMALL[] = {" + ", ".join(f"{v:.9e}f" for v in coef[::-1]) + "};")
    for deg in (8, 10, 12, 14):
        coef, rel, einrel = fit_large(deg, 32.0)
        print("large deg", deg, "rel err H", rel, "rel err wrt Ein", einrel)
    coef, rel, einrel = fit_large(12, 32.0)
    print("static const float E1_LARGE[] = {" + ", ".join(f"{v:.9e}f" for v in coef[::-1]) + "};")
