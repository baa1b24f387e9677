"""GPU unit tests of the building blocks of K1: Philox stream, exact samplers, device lnL."""
import numpy as np
import pytest
import scipy.stats as st

from oracle import oracle as orc
from tests.gpu_util import engine, philox2x32_10, philox4x32_10
from tests.util import load

pytestmark = pytest.mark.gpu


def test_philox_reference_vectors():
    # Random123 known-answer vectors for philox4x32-10
    assert philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox2x32_reference_vectors_and_device_stream():
    # Random123 known-answer vectors for philox2x32-10
    assert philox2x32_10([0, 0], 0) == [0xff1dae59, 0x6cd10df2]
    assert philox2x32_10([0xffffffff, 0xffffffff], 0xffffffff) == [0x2c3f628b, 0xab4fd7ad]
    assert philox2x32_10([0x243f6a88, 0x85a308d3], 0x13198a2e) == [0xdd7ce038, 0xf62a4c12]
    e = engine()
    assert list(e.debug_pairs(0, 0, 0, 1)) == [0xff1dae59, 0x6cd10df2]
    seed, sim, uid = 0x299f31d0a4093822, (5 << 32) | 0x85a308d3, 0x13198a2e
    exp = []
    for i in range(4):
        exp += philox2x32_10([i | (5 << 24), uid ^ (seed >> 32)], (sim ^ seed) & 0xffffffff)
    assert list(e.debug_pairs(seed, sim, uid, 4)) == exp
    # distinct simulations / genotypes never share a stream (the keying is a bijection of (uid, sim))
    a = e.debug_pairs(7, 1, 2, 2)
    assert list(a) != list(e.debug_pairs(7, 2, 1, 2)) and list(a) != list(e.debug_pairs(7, 1, 3, 2))


def test_device_philox_stream_layout():
    e = engine()
    assert list(e.debug_philox(0, 0, 0, 0, 1)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    seed, sim, uid, stream = 0x299f31d0a4093822, (5 << 32) | 0x85a308d3, 0x13198a2e, 1
    got = e.debug_philox(seed, sim, uid, stream, 3)
    exp = []
    for blk in range(3):
        exp += philox4x32_10([blk, sim & 0xffffffff, uid, stream | ((sim >> 32) << 8)],
                             [seed & 0xffffffff, seed >> 32])
    assert list(got) == exp


BINOM_CASES = [(1, 0.3), (5, 0.5), (20, 0.3), (100, 0.05), (1000, 0.005), (300000, 2e-5), (40, 0.9), (12, 0.97),
               (100, 0.1), (200, 0.08), (1000, 0.0373), (5000, 0.5), (150000, 0.05), (14000, 0.4), (20000, 0.7),
               (1200000, 0.0123), (3000000, 0.1), (7503, 0.9991), (37, 0.4999), (1000000, 0.5),
               # either side of the inversion / BTRS switch at n min(p, q) = 20 (SONICS_INV_BELOW)
               (150000, 1.3e-4), (2000000, 9.9e-6), (45, 0.44), (1999, 0.01), (2001, 0.01), (50, 0.6)]


def _k1_request_mix():
    """Ten (n, p) requests taken from what K1 actually asks for: the draws of oracle-recorded cfg2 / cfg3
    simulations, at the deciles of n * min(p, 1-p) over the requests with n p >= 1."""
    from oracle import oracle as orc
    from tests.util import CFG2, CFG3
    req = []
    for gt, kw, seeds in ((CFG2, {}, range(6)), (CFG3, dict(n_cycles=30, capture_cycle=16), range(2))):
        reads = orc.parse_reads(gt)
        for sd in seeds:
            _, _, script = orc.record_script(orc.make_cfg(rng_mode=1, with_readout=False, **kw), reads, 4242 + sd)
            b = script[script["kind"] == 1]
            req += [(int(n), float(p)) for n, p in zip(b["n"], b["p"])]
    req = sorted(set(r for r in req if r[0] * min(r[1], 1 - r[1]) >= 1.0), key=lambda r: r[0] * min(r[1], 1 - r[1]))
    return [req[int(q * (len(req) - 1))] for q in np.linspace(0.02, 0.98, 10)]


K1_MIX = _k1_request_mix()


@pytest.mark.parametrize("n,p", BINOM_CASES + K1_MIX)
def test_binomial_sampler_matches_exact_pmf(n, p):
    e = engine()
    N = 20_000_000
    x = e.debug_sample("binomial", n, p, N, seed=1234 + n)
    p32 = float(np.float32(p))  # the kernel receives p as a float
    assert x.min() >= 0 and x.max() <= n
    mu, sd = n * p32, np.sqrt(n * p32 * (1 - p32))
    assert abs(x.mean() - mu) < 5 * sd / np.sqrt(N) + 1e-12
    assert abs(x.var() - sd * sd) < 6 * sd * sd * np.sqrt(2.0 / N) + 5 * abs(sd * (1 - 2 * p32)) / np.sqrt(N) + 1e-9
    lo, hi = int(max(0, np.floor(mu - 7 * sd - 2))), int(min(n, np.ceil(mu + 7 * sd + 2)))
    ks = np.arange(lo, hi + 1)
    pmf = st.binom.pmf(ks, n, p32)
    obs = np.bincount(np.clip(x, lo, hi) - lo, minlength=len(ks)).astype(np.float64)
    expc = pmf * N
    # pool the tails so every cell expects >= 25
    keep = expc >= 25
    o = np.append(obs[keep], obs[~keep].sum())
    ex = np.append(expc[keep], N - expc[keep].sum())
    if ex[-1] < 5:
        o, ex = o[:-1], ex[:-1] * (o[:-1].sum() / ex[:-1].sum())
    chi2 = ((o - ex) ** 2 / ex).sum()
    pval = st.chi2.sf(chi2, len(o) - 1)
    assert pval > 1e-4, (n, p, chi2, len(o), pval)


@pytest.mark.parametrize("lam", [0.3, 3.0, 9.9, 10.0, 17.5, 250.0, 41234.5, 1.5e6])
def test_poisson_sampler_matches_exact_pmf(lam):
    e = engine()
    N = 20_000_000
    x = e.debug_sample("poisson", 0, lam, N, seed=77)
    sd = np.sqrt(lam)
    assert abs(x.mean() - lam) < 5 * sd / np.sqrt(N)
    lo, hi = int(max(0, np.floor(lam - 7 * sd - 2))), int(np.ceil(lam + 7 * sd + 2))
    ks = np.arange(lo, hi + 1)
    expc = st.poisson.pmf(ks, lam) * N
    obs = np.bincount(np.clip(x, lo, hi) - lo, minlength=len(ks)).astype(np.float64)
    keep = expc >= 25
    o = np.append(obs[keep], obs[~keep].sum())
    ex = np.append(expc[keep], N - expc[keep].sum())
    if ex[-1] < 5:
        o, ex = o[:-1], ex[:-1] * (o[:-1].sum() / ex[:-1].sum())
    pval = st.chi2.sf(((o - ex) ** 2 / ex).sum(), len(o) - 1)
    assert pval > 1e-4, (lam, pval)


def test_sampler_edge_cases():
    e = engine()
    assert (e.debug_sample("binomial", 0, 0.3, 1000, 1) == 0).all()
    assert (e.debug_sample("binomial", 50, 0.0, 1000, 1) == 0).all()
    assert (e.debug_sample("binomial", 50, 1.0, 1000, 1) == 50).all()
    assert (e.debug_sample("poisson", 0, 0.0, 1000, 1) == 0).all()


def _rows(kats):
    x = np.zeros((len(kats), 500), dtype=np.int64)
    c = np.zeros((len(kats), 500), dtype=np.int64)
    for i, kat in enumerate(kats):
        for a, n in kat["reads"].items():
            x[i, int(a)] = n
        for a, n in kat["pool"].items():
            c[i, int(a)] = n
    return x, c


def test_device_lnl_known_answers():
    """Deterministic lnL gate (north star): fixed pool, |gpu - reference| <= 1e-9 relative."""
    e = engine()
    kats = load("lnl_kat.json")["kats"]
    x, c = _rows(kats)
    got = e.mvhyperg(x, c)
    for i, kat in enumerate(kats):
        ref = orc.mvhyperg(x[i], c[i])
        assert abs(got[i] - ref) <= 1e-9 * abs(ref), (kat["name"], got[i], ref)
        if "survey_lnL" in kat:
            assert abs(got[i] - kat["survey_lnL"]) <= 1e-9 * abs(kat["survey_lnL"])
    assert e.mvhyperg(x[0], c[0]) == got[0]


def test_device_lnl_random_pools_vs_oracle():
    e = engine()
    rs = np.random.RandomState(11)
    n = 20000
    x = np.zeros((n, 64), dtype=np.int64)
    c = np.zeros((n, 64), dtype=np.int64)
    for i in range(n):
        lo = rs.randint(1, 30)
        nb = rs.randint(3, 20)
        peak = lo + rs.randint(0, nb)
        for b in range(nb):
            c[i, lo + b] = int(10 ** rs.uniform(0, 6.3) * np.exp(-0.5 * ((lo + b - peak) / 2.0) ** 2)) + rs.randint(0, 3)
        reads_bins = rs.choice(np.arange(lo, lo + nb), size=rs.randint(2, min(nb, 9) + 1), replace=False)
        for b in reads_bins:
            x[i, b] = min(c[i, b], rs.randint(1, 2000))
    got = e.mvhyperg(x, c)
    ref = np.array([orc.mvhyperg(x[i], c[i]) for i in range(n)])
    ok = np.isfinite(ref) & (ref > -1e300) & (ref != 0)
    rel = np.abs(got[ok] - ref[ok]) / np.abs(ref[ok])
    print("device-vs-oracle lnL: max rel %.3e, 99.9%% %.3e, exact %.1f%%" % (rel.max(), np.quantile(rel, 0.999),
                                                                          100 * (rel == 0).mean()))
    assert rel.max() <= 1e-9
    assert (got[~ok] == ref[~ok]).all()


def test_device_lnl_degenerate():
    e = engine()
    x = np.zeros((3, 8), dtype=np.int64)
    c = np.zeros((3, 8), dtype=np.int64)
    c[1, 3], x[1, 3] = 10, 11      # reads > pool: Fortran arithmetic gives exactly 0.0
    c[2, 2], x[2, 2] = -1, 0       # negative count
    got = e.mvhyperg(x, c)
    assert got[0] == -1.7976931348623157e308
    assert got[1] == orc.mvhyperg(x[1], c[1]) == 0.0
    assert got[2] == -1.7976931348623157e308
