"""CPU property tests of the closed-form bounds the device samplers rely on (csrc/samplers.cuh).

btrs_trial() accepts / rejects a candidate k without evaluating ln f(k)/f(m) whenever ln v' falls outside
[t - rho, t + rho], a quartic in k - m with a fifth-order remainder bound.  That is only exact if the interval
really brackets ln f(k)/f(m) for every k in the stated range.  Checked here against the exact log-pmf ratio."""
import numpy as np
import pytest
from scipy.special import gammaln


def logf(n, p, k):
    return gammaln(n + 1) - gammaln(k + 1) - gammaln(n - k + 1) + k * np.log(p) + (n - k) * np.log1p(-p)


def _fc(z):
    return gammaln(z + 1) - ((z + 0.5) * np.log(z + 1) - (z + 1) + 0.5 * np.log(2 * np.pi))


def _log_ratio_stable(n, p, k, m):
    """ln f(k)/f(m) in fp64 without the cancellation of the gammaln form (deviance arrangement, exact fc)."""
    q = 1 - p
    d = k - m
    a0, b0, a, b = m + 1, n - m + 1, k + 1, n - k + 1
    c0 = np.log1p(((n + 1) * p - m - q) / (q * a0))
    return (-(a * np.log1p(d / a0) - d) - (b * np.log1p(-d / b0) + d) + 0.5 * (np.log1p(d / a0) + np.log1p(-d / b0))
            + d * c0 + _fc(m) + _fc(n - m) - _fc(k) - _fc(n - k))


def _squeeze_consts(n, p, dtype=np.float64):
    """The per-draw constants of btrs_setup() (csrc/samplers.cuh)."""
    f = dtype
    # m and frac = (n+1) p - m: the device forms n p as an exact fp32 pair (hi + lo), so they carry ~1e-7 absolute
    # error, not the 4e-3 of a plain fp32 product at n p = 4e4
    m = f(np.floor((float(n) + 1) * float(p)))
    frac = f((float(n) + 1) * float(p) - float(m))
    n, p = f(n), f(p)
    q = f(1) - p
    a0, b0 = m + f(1), n - m + f(1)
    ra, rb = f(1) / a0, f(1) / b0
    eps = (frac - q) / (q * a0)
    w = eps / (f(2) + eps)
    w2 = w * w
    c0 = f(2) * w * (f(1) + w2 * (f(1 / 3) + w2 * (f(1 / 5) + w2 * (f(1 / 7) + w2 * f(1 / 9)))))
    L = f(1 / 12) * (ra * ra - rb * rb) + f(0.5) * (ra - rb) + c0
    Q = f(1 / 12) * (ra ** 3 + rb ** 3) + f(0.25) * (ra * ra + rb * rb) + f(0.5) * (ra + rb)
    C3 = f(1 / 12) * (ra ** 4 - rb ** 4) + f(1 / 6) * ((ra * ra - rb * rb) + (ra ** 3 - rb ** 3))
    C4 = f(-0.125) * (ra ** 4 + rb ** 4) + f(-1 / 12) * (ra ** 3 + rb ** 3)
    R5 = f(0.25) * (ra ** 4 + rb ** 4)
    dmax = np.floor(f(0.9) * min(a0, b0))
    return m, L, Q, C3, C4, R5, dmax


def test_quartic_squeeze_brackets_the_log_pmf_ratio():
    """btrs_trial() accepts / rejects a candidate without the exact test whenever ln v' falls outside
    [t - rho, t + rho], t = L d - Q d^2 + C3 d^3 + C4 d^4, rho = 0.25 |d|^5 (ra^4 + rb^4), |d| <= 0.9 min(a0, b0).
    That is only exact if the interval brackets ln f(k)/f(m) for every candidate in that range -- checked here, in
    fp64 and with the constants and the polynomial evaluated in fp32 plus the kernel's 1e-5 (1 + |t|) margin, against
    the log-pmf ratio (cancellation-free fp64 form; mpmath for n >= 1.5e5)."""
    import mpmath as mp
    mp.mp.dps = 40
    checked = 0
    worst = 0.0
    ns = [11, 12, 20, 21, 25, 37, 40, 55, 64, 100, 150, 333, 700, 1000, 2001, 5000, 10000, 30000, 77777]
    ps = [0.5, 0.499, 0.49, 0.45, 0.4, 0.33, 0.25, 0.2, 0.15, 0.1, 0.07, 0.05, 0.0123, 0.01, 0.003, 0.0004, 0.00001]
    for n in ns:
        for p in ps:
            if n * p < 10:
                continue
            m, L, Q, C3, C4, R5, dmax = _squeeze_consts(n, p)
            ks = np.arange(max(0, m - dmax), min(n, m + dmax) + 1)
            d = ks - m
            exact = _log_ratio_stable(n, p, ks, m)
            t = d * (L + d * (-Q + d * (C3 + d * C4)))
            rho = np.abs(d) ** 5 * R5
            assert (np.abs(exact - t) <= rho + 1e-9).all(), (n, p)
            worst = max(worst, (np.abs(exact - t) / (rho + 1e-9)).max())
            # fp32 evaluation as on the device
            m32, L32, Q32, C332, C432, R532, dmax32 = _squeeze_consts(n, p, np.float32)
            assert m32 == m and dmax32 == dmax
            d32 = d.astype(np.float32)
            t32 = d32 * (L32 + d32 * (-Q32 + d32 * (C332 + d32 * C432)))
            rho32 = np.abs(d32) ** 5 * R532
            eps = np.float32(1e-5) * (1 + np.abs(t32))
            assert (exact <= (t32 + rho32 + eps).astype(float)).all(), (n, p)
            assert (exact >= (t32 - rho32 - eps).astype(float)).all(), (n, p)
            checked += len(ks)
    assert checked > 300000 and worst < 0.75, (checked, worst)   # headroom of the remainder constant: >= 25 %
    rs = np.random.RandomState(0)
    for n in (150000, 1200000, 5000000):
        for p in (0.5, 0.25, 0.1, 0.05, 0.01, 0.0004, 0.00001):
            if n * p < 10:
                continue
            m, L, Q, C3, C4, R5, dmax = _squeeze_consts(n, p)
            m32, L32, Q32, C332, C432, R532, dmax32 = _squeeze_consts(n, p, np.float32)
            sd = np.sqrt(n * p * (1 - p))
            ds = np.unique(np.clip(np.round(np.concatenate([rs.randn(12) * sd * 3, [-dmax, dmax, -1, 1, 2, -2],
                                                            rs.uniform(-dmax, dmax, 4)])), -dmax, dmax))
            N, P = mp.mpf(n), mp.mpf(p)

            def lf(x):
                return (mp.loggamma(N + 1) - mp.loggamma(x + 1) - mp.loggamma(N - x + 1) + x * mp.log(P) +
                        (N - x) * mp.log(1 - P))
            for d in ds:
                exact = float(lf(mp.mpf(m + d)) - lf(mp.mpf(m)))
                t = d * (L + d * (-Q + d * (C3 + d * C4)))
                assert abs(exact - t) <= abs(d) ** 5 * R5 + 1e-11, (n, p, d)
                d32 = np.float32(d)
                t32 = float(d32 * (L32 + d32 * (-Q32 + d32 * (C332 + d32 * C432))))
                band = float(np.abs(d32) ** 5 * R532) + 1e-5 * (1 + abs(t32))
                assert abs(exact - t32) <= band, (n, p, d, exact, t32, band)
                checked += 1
    # beyond dmax the upper bound at the clamped distance still holds (the pmf is unimodal)
    for n, p in ((30, 0.4), (200, 0.05), (5000, 0.003)):
        m, L, Q, C3, C4, R5, dmax = _squeeze_consts(n, p)
        for sgn in (-1, 1):
            dc = sgn * dmax
            top = dc * (L + dc * (-Q + dc * (C3 + dc * C4))) + abs(dc) ** 5 * R5
            ks = np.arange(m + dmax + 1, n + 1) if sgn > 0 else np.arange(0, m - dmax)
            if len(ks):
                assert (_log_ratio_stable(n, p, ks, m) <= top + 1e-9).all(), (n, p, sgn)


def test_stirling_tail_series():
    """fc(k) = ln k! - [(k+1/2) ln(k+1) - (k+1) + ln(2 pi)/2] against its 3-term series."""
    k = np.arange(0, 2000, dtype=np.float64)
    exact = gammaln(k + 1) - ((k + 0.5) * np.log(k + 1) - (k + 1) + 0.5 * np.log(2 * np.pi))
    r = 1.0 / (k + 1)
    series = r * (1 / 12 - r * r * (1 / 360 - r * r / 1260))
    err = np.abs(series - exact)
    assert err[10:].max() < 1e-9          # the regime BTRS evaluates it in (m, n-m, n-k >= 10)
    assert err[0] < 4e-4 and err[1] < 2e-5  # k < 10 only in the far lower tail (pmf < e^-10)


def test_ptrs_log_pmf_rearrangement():
    """ln Poisson pmf as used by pois_trial: k (log1p(y) - y) - ln(2 pi k)/2 - 1/(12k) + 1/(360k^3)."""
    for lam in (10.0, 37.5, 1e3, 4.1e5, 1.5e6):
        sd = np.sqrt(lam)
        k = np.arange(max(10, int(lam - 8 * sd)), int(lam + 8 * sd) + 1, dtype=np.float64)
        exact = -lam + k * np.log(lam) - gammaln(k + 1)
        y = (lam - k) / k
        approx = k * (np.log1p(y) - y) - 0.5 * np.log(2 * np.pi * k) - (1 / (12 * k) - 1 / (360 * k ** 3))
        assert np.abs(approx - exact).max() < 5e-7 * max(1.0, lam ** 0.5), lam


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def _stirling_tail32(k):
    r = np.float32(1) / (_f32(k) + np.float32(1))
    r2 = r * r
    return r * (np.float32(0.083333333333) - r2 * (np.float32(0.002777777778) - r2 * np.float32(0.000793650794)))


def _deviance_term32(x, x0, dx):
    """fp32 restatement of deviance_term() in csrc/samplers.cuh: x ln(x/x0) - dx."""
    v = dx / (x + x0)
    v2 = v * v
    s = np.float32(1 / 19)
    for j in (17, 15, 13, 11, 9, 7, 5, 3):
        s = s * v2 + np.float32(1 / j)
    series = np.float32(2) * x * (s * v2 * v) + dx * v
    direct = x * np.log(x / x0).astype(np.float32) - dx
    return np.where(np.abs(v) > np.float32(0.5), direct, series).astype(np.float32)


def _btrs_full_h32(n, p, k, m, frac):
    n, p, k, m, frac = _f32(n), _f32(p), _f32(k), _f32(m), _f32(frac)
    one = np.float32(1)
    q = one - p
    d = k - m
    a0, b0, a, b = m + one, n - m + one, k + one, n - k + one
    ga = _deviance_term32(a, a0, d)
    gb = _deviance_term32(b, b0, -d)
    lg = np.float32(0.5) * np.log((a / a0) * (b / b0)).astype(np.float32)
    eps = (frac - q) / (q * a0)
    w = eps / (np.float32(2) + eps)
    w2 = w * w
    c0 = np.float32(2) * w * (one + w2 * (np.float32(1 / 3) + w2 * (np.float32(1 / 5) + w2 * (
        np.float32(1 / 7) + w2 * np.float32(1 / 9)))))
    return (lg - (ga + gb) + d * c0 + _stirling_tail32(m) + _stirling_tail32(n - m) - _stirling_tail32(k) -
            _stirling_tail32(n - k))


def test_exact_test_in_deviance_form_is_accurate_in_fp32():
    """btrs_full(): ln f(k)/f(m) = -g(a;a0) - g(b;b0) + ln(ab/(a0 b0))/2 + d c0 + Stirling tails, evaluated in fp32
    as the device does, against the fp64 log-pmf ratio over 8 sigma around the mode for n up to 5e6 (where the
    textbook three-logarithm form loses 3e-3 to cancellation)."""
    worst = worst_core = 0.0
    checked = 0
    for n in (25, 40, 100, 333, 1000, 5000, 30000, 150000, 1200000, 5000000):
        for p in (0.5, 0.4, 0.25, 0.1, 0.05, 0.01, 0.003, 0.0004, 0.00001):
            if n * p < 10:
                continue
            m = np.floor((n + 1) * p)
            sd = np.sqrt(n * p * (1 - p))
            ks = np.arange(max(0, int(m - 8 * sd)), min(n, int(m + 8 * sd)) + 1).astype(float)
            exact = logf(n, p, ks) - logf(n, p, m)
            keep = exact > -40          # candidates with ln f(k)/f(m) < -40 are accepted with probability e^-40
            ks, exact = ks[keep], exact[keep]
            h = _btrs_full_h32(n, p, ks, m, (n + 1) * p - m).astype(float)
            err = np.abs(h - exact)
            worst = max(worst, err.max())
            core = (ks >= 10) & (n - ks >= 10)
            worst_core = max(worst_core, err[core].max())
            checked += len(ks)
    assert checked > 100000
    assert worst_core < 2e-5, worst_core   # the Stirling series is the limit where k (or n - k) < 10:
    assert worst < 4e-4, worst             # fc(0) is off by 3e-4, fc(1) by 2e-5 (test_stirling_tail_series)


def _cf_cdf(n, p, ks):
    """cdf of the approximate sampler (samplers.cuh: binomial_normal) in exact arithmetic: the normal deviate z with
    n p + sqrt(n p q - 1/12) z + (q - p)(z^2 - 1)/6 < k + 1/2."""
    import scipy.stats as st
    q = 1.0 - p
    s, g, mu = np.sqrt(n * p * q - 1.0 / 12.0), (q - p) / 6.0, n * p
    b = ks + 0.5
    if abs(g) < 1e-12:
        z = (b - mu) / s
    else:
        disc = s * s - 4.0 * g * (mu - g - b)
        z = np.where(disc > 0, (-s + np.sqrt(np.maximum(disc, 0.0))) / (2.0 * g), -np.inf)
    return st.norm.cdf(z)


@pytest.mark.parametrize("npq", [20, 100, 1000, 1e5])
def test_fast_sampler_formula_error_bound(npq):
    """The opt-in approximate sampler (SONICS_SAMPLERS_FAST): |cdf - binomial cdf| <= 1.2e-2 / (n p q), pmf within
    0.9 % / 0.08 % inside three sigma at n p q = 100 / 1000 -- the numbers quoted in samplers.cuh and the header."""
    import scipy.stats as st
    for p in (0.5, 0.3, 0.1, 0.01, 1e-4):
        n = int(round(npq / (p * (1 - p))))
        s, mu = np.sqrt(n * p * (1 - p)), n * p
        ks = np.arange(max(0, int(mu - 8 * s)), min(n, int(mu + 8 * s)) + 1).astype(np.float64)
        ap = _cf_cdf(n, p, ks)
        assert np.abs(ap - st.binom.cdf(ks, n, p)).max() <= 1.2e-2 / npq, (npq, p)
        pa = np.diff(np.concatenate([_cf_cdf(n, p, np.array([ks[0] - 1.0])), ap]))
        pe = st.binom.pmf(ks, n, p)
        c = np.abs(ks - mu) <= 3 * s
        rel = np.abs(pa[c] / pe[c] - 1).max()
        assert rel <= {20: 0.10, 100: 0.009, 1000: 0.0008, 1e5: 1e-5}[npq], (npq, p, rel)
