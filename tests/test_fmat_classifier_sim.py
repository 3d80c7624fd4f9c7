"""CPU: the error bounds of the filter kernel's fast inlier classification, checked by simulation.

monocular_slam_b200/csrc/fmat.cu decides most points of a candidate matrix in single precision (cand32_init / side32 /
classify32) and the rest with a division-free double-precision test (classify); both claim to decide a point only when
OpenCV's own expression -- (float)max(d1*d1/den1, d2*d2/den2) <= t2, evaluated in double without FMA -- would decide it the same
way.  This file restates that arithmetic in numpy (float32 FMA = one rounding of the exact double product-sum) and checks
the claim on ~1.5 million (candidate, point) pairs: realistic and badly scaled candidates, coordinates from 0.01 to 4K,
thresholds 0.5-10 px.  The constants below mirror the kernel's; change them together.  Also simulated: the
single-precision screen of the iteration-budget formula (iter_limit_of) against the exact expression."""
import math

import numpy as np
import pytest

import oracle
from monocular_slam_b200 import synthetic as syn

f32, f64 = np.float32, np.float64


def fma32(a, b, c): return (a.astype(f64) * b.astype(f64) + c.astype(f64)).astype(f32)
U4, U8 = f32(2.384185791015625e-07), f32(4.76837158203125e-07)

def cand32(F, cmax):
    Ff = F.astype(f32).ravel()
    X1, Y1, X2, Y2 = [f32(v) for v in cmax]; tiny = f32(1e-30)
    a = np.abs(Ff)
    def mag(i, j, k, X, Y): return f32(f32(f64(a[i]) * f64(X) + f64(f32(f64(a[j]) * f64(Y) + f64(a[k])))) + tiny)
    A2, B2, C2 = mag(0, 1, 2, X1, Y1), mag(3, 4, 5, X1, Y1), mag(6, 7, 8, X1, Y1)
    A1, B1, C1 = mag(0, 3, 6, X2, Y2), mag(1, 4, 7, X2, Y2), mag(2, 5, 8, X2, Y2)
    E2 = f32(U8 * f32(f64(X2) * f64(A2) + f64(f32(f64(Y2) * f64(B2) + f64(C2)))))
    E1 = f32(U8 * f32(f64(X1) * f64(A1) + f64(f32(f64(Y1) * f64(B1) + f64(C1)))))
    return Ff, f32(U4 * A2), f32(U4 * B2), E2, f32(U4 * A1), f32(U4 * B1), E1

def side32(a, b, d, ea, eb, E, tlo32, thi32):
    la = np.maximum(np.abs(a) - ea, f32(0)); lb = np.maximum(np.abs(b) - eb, f32(0)); ha = np.abs(a) + ea; hb = np.abs(b) + eb
    den_lo = fma32(la, la, lb * lb); den_hi = fma32(ha, ha, hb * hb)
    dhi = np.abs(d) + E; dlo = np.maximum(np.abs(d) - E, f32(0))
    with np.errstate(over='ignore', invalid='ignore'):
        regular = (den_lo > f32(1e-30)) & (den_hi < f32(1e30)) & (dhi < f32(1e15))
        inn = dhi * dhi <= tlo32 * den_lo; out = dlo * dlo >= thi32 * den_hi
    return np.where(regular, np.where(inn, 1, np.where(out, 0, -1)), -1)

def classify32(F, p1, p2, cmax, t2):
    Ff, ea2, eb2, E2, ea1, eb1, E1 = cand32(F, cmax)
    tlo = f64(t2) * (1 - 1e-9); thi = f64(t2) * (1 + 1e-6)
    tlo32 = np.nextafter(f32(tlo * (1 - 4e-6)), f32(0)); thi32 = np.nextafter(f32(thi * (1 + 4e-6)), f32(np.inf))
    x1, y1, x2, y2 = p1[:, 0], p1[:, 1], p2[:, 0], p2[:, 1]
    c = lambda i: np.full_like(x1, Ff[i])
    with np.errstate(over='ignore', invalid='ignore'):
        a2 = fma32(c(0), x1, fma32(c(1), y1, c(2))); b2 = fma32(c(3), x1, fma32(c(4), y1, c(5))); c2 = fma32(c(6), x1, fma32(c(7), y1, c(8)))
        d2 = fma32(x2, a2, fma32(y2, b2, c2))
        a1 = fma32(c(0), x2, fma32(c(3), y2, c(6))); b1 = fma32(c(1), x2, fma32(c(4), y2, c(7))); c1 = fma32(c(2), x2, fma32(c(5), y2, c(8)))
        d1 = fma32(x1, a1, fma32(y1, b1, c1))
        s2 = side32(a2, b2, d2, ea2, eb2, E2, tlo32, thi32); s1 = side32(a1, b1, d1, ea1, eb1, E1, tlo32, thi32)
    return np.where((s1 == 0) | (s2 == 0), 0, np.where((s1 == 1) & (s2 == 1), 1, -1))

def classify64(F, p1, p2, t2):
    F = F.ravel(); x1, y1, x2, y2 = [v.astype(f64) for v in (p1[:, 0], p1[:, 1], p2[:, 0], p2[:, 1])]
    tlo = f64(t2) * (1 - 1e-9); thi = f64(t2) * (1 + 1e-6)
    with np.errstate(over='ignore', invalid='ignore'):
        a = F[0]*x1 + F[1]*y1 + F[2]; b = F[3]*x1 + F[4]*y1 + F[5]; c = F[6]*x1 + F[7]*y1 + F[8]
        den2 = a*a + b*b; d2 = x2*a + y2*b + c
        a = F[0]*x2 + F[3]*y2 + F[6]; b = F[1]*x2 + F[4]*y2 + F[7]; c = F[2]*x2 + F[5]*y2 + F[8]
        den1 = a*a + b*b; d1 = x1*a + y1*b + c
        q1, q2 = d1*d1, d2*d2
        regular = (den1 > 0) & (den1 < 1e300) & (den2 > 0) & (den2 < 1e300)
        inn = (q1 <= tlo*den1) & (q2 <= tlo*den2); out = (q1 >= thi*den1) | (q2 >= thi*den2)
    return np.where(regular & (inn | out), np.where(inn, 1, 0), -1)


def test_fast_classification_never_disagrees_with_the_exact_test():
    r = np.random.default_rng(5)
    tot = dec32 = dec64 = bad32 = bad64 = 0
    for trial in range(60):
        size = [(640, 480), (1241, 376), (1920, 1080), (3840, 2160)][trial % 4]
        n = int(r.integers(200, 1500))
        p1, p2 = syn.two_view_matches(int(r.integers(1 << 30)), n, float(r.uniform(0.2, 0.95)), float(r.uniform(0, 2)), size)
        if trial % 7 == 0:                                   # tiny coordinates
            p1, p2 = (p1 * f32(0.01)).astype(f32), (p2 * f32(0.01)).astype(f32)
        thr = float(r.choice([0.5, 1, 3, 5, 10]))
        t2 = f32(thr * thr)
        cmax = [np.abs(p1[:, 0]).max(), np.abs(p1[:, 1]).max(), np.abs(p2[:, 0]).max(), np.abs(p2[:, 1]).max()]
        for s in range(12):
            idx = r.choice(n, 7, replace=False)
            for F in oracle.fm_7point(p1[idx], p2[idx]):
                if s % 4 == 3:
                    F = F * float(10.0 ** r.integers(-12, 12))                  # badly scaled candidates
                exact = oracle.fm_errors(p1, p2, F) <= t2
                c32, c64 = classify32(F, p1, p2, cmax, t2), classify64(F, p1, p2, t2)
                tot += n
                dec32 += int((c32 >= 0).sum())
                dec64 += int((c64 >= 0).sum())
                bad32 += int(((c32 >= 0) & ((c32 == 1) != exact)).sum())
                bad64 += int(((c64 >= 0) & ((c64 == 1) != exact)).sum())
    assert bad32 == 0 and bad64 == 0
    assert dec32 > 0.99 * tot and dec64 > 0.9999 * tot       # and they are worth having: > 99 % decided in single precision


DBL_MIN = 2.2250738585072014e-308


def _budget_exact(conf, n, good, n0):
    ep = (n - good) / n
    num, denom = max(1 - conf, DBL_MIN), 1 - (1 - ep) ** 7
    if denom < DBL_MIN:
        return 0
    num, denom = math.log(num), math.log(denom)
    if denom >= 0:
        return n0
    q = num / denom
    return n0 if q >= n0 else min(n0, int(round(q)))


def _budget_screened(conf, n, good, n0):
    """iter_limit_of of fmat.cu"""
    lognum = math.log(max(1 - conf, DBL_MIN))
    w = f32(good) / f32(n)
    w2 = f32(w * w)
    df = f32(f32(1) - f32(f32(f32(w2 * w2) * w2) * w))
    lim = f32(f32(2) * f32(n0) + f32(2))
    if df >= 1:
        if f32(-lognum) >= f32(lim * f32(1.2e-7)):
            return n0
    else:
        with np.errstate(divide="ignore", invalid="ignore"):
            qf = f32(lognum) / np.log(df, dtype=np.float32) if df > 0 else f32(-0.0)
        if not (qf < lim):
            return n0
    return _budget_exact(conf, n, good, n0)


@pytest.mark.parametrize("conf", [1e-15, 1e-9, 1e-5, 1e-3, 0.1, 0.5, 0.85, 0.99, 0.999999, 1 - 2.3e-16])
def test_budget_screen_equals_exact_formula(conf):
    for n in (15, 50, 84, 333, 1000, 9600, 100000):
        goods = set(range(7, min(n, 120) + 1)) | set(range(max(7, n - 120), n + 1)) | set(np.linspace(7, n, 80).astype(int))
        for n0 in (1000, 589, 19, 3):
            for g in goods:
                assert _budget_screened(conf, n, g, n0) == _budget_exact(conf, n, g, n0), (conf, n, g, n0)
