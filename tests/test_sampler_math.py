"""CPU property tests of the closed-form bounds the device samplers rely on (csrc/samplers.cuh).

btrs_trial() accepts / rejects a candidate k without evaluating ln f(k)/f(m) whenever
ln v' falls outside [t - rho, t + rho] (BTPE step 5.2).  That is only exact if the interval really
brackets ln f(k)/f(m) for every k the sampler can propose -- including |k - m| <= 20, where BTPE itself
switches to an explicit product.  Checked here against scipy's exact log-pmf."""
import numpy as np
from scipy.special import gammaln


def logf(n, p, k):
    return gammaln(n + 1) - gammaln(k + 1) - gammaln(n - k + 1) + k * np.log(p) + (n - k) * np.log1p(-p)


def test_btpe_squeeze_brackets_the_log_pmf_ratio():
    checked = 0
    for n in (25, 40, 100, 333, 1000, 5000, 30000, 150000, 1200000, 5000000):
        for p in (0.5, 0.4, 0.25, 0.1, 0.05, 0.01, 0.003, 0.0004, 0.00001):
            if n * p < 10:
                continue
            q = 1 - p
            npq = n * p * q
            m = np.floor((n + 1) * p)
            sd = np.sqrt(npq)
            ks = np.arange(max(0, int(m - 12 * sd)), min(n, int(m + 12 * sd)) + 1).astype(float)
            d = np.abs(ks - m)
            ok = d < npq / 2 - 1
            ks, d = ks[ok], d[ok]
            exact = logf(n, p, ks) - logf(n, p, m)
            t = -d * d / (2 * npq)
            rho = (d / npq) * ((d * (d / 3 + 0.625) + 1 / 6) / npq + 0.5)
            assert (exact >= t - rho - 1e-9).all() and (exact <= t + rho + 1e-9).all(), (n, p)
            checked += len(ks)
    assert checked > 100000


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
