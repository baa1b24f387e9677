"""Pins the oracle's restatement of numpy's legacy RandomState (MT19937, binomial,
poisson, uniform, randint/choice) bit for bit against golden vectors produced by
numpy.random itself (tests/golden/make_golden.py::gen_rng).  These are the
third-party primitives behind sonics.pyx:37-43,321,324,354,368,420,428,450-462,478."""
import numpy as np

from oracle import oracle as orc
from tests.util import load

G = load("numpy_legacy_rng.json")


def test_mt19937_stream():
    L = orc.lib()
    for case in G["stream"]:
        s = L.oracle_rng_new(case["seed"])
        assert [L.oracle_rng_uint32(s) for _ in range(8)] == case["uint32"]
        L.oracle_rng_seed(s, case["seed"])
        assert [L.oracle_rng_double(s).hex() for _ in range(4)] == case["doubles"]
        L.oracle_rng_free(s)


def test_binomial_matches_numpy_legacy():
    L = orc.lib()
    s = L.oracle_rng_new(0)
    for case in G["binomial"]:
        L.oracle_rng_seed(s, case["seed"])
        p = float.fromhex(case["p"])
        draws = [L.oracle_binomial(s, case["n"], p) for _ in range(6)]
        assert draws == case["draws"], case
        assert L.oracle_rng_double(s).hex() == case["next"], case
    L.oracle_rng_free(s)


def test_poisson_matches_numpy_legacy():
    L = orc.lib()
    s = L.oracle_rng_new(0)
    for case in G["poisson"]:
        L.oracle_rng_seed(s, case["seed"])
        lam = float.fromhex(case["lam"])
        draws = [L.oracle_poisson(s, lam) for _ in range(6)]
        assert draws == case["draws"], case
        assert L.oracle_rng_double(s).hex() == case["next"], case
    L.oracle_rng_free(s)


def test_uniform_randint_choice():
    L = orc.lib()
    s = L.oracle_rng_new(0)
    for case in G["uniform"]:
        L.oracle_rng_seed(s, case["seed"])
        lo, hi = float.fromhex(case["lo"]), float.fromhex(case["hi"])
        assert [L.oracle_uniform(s, lo, hi).hex() for _ in range(5)] == case["draws"]
    for case in G["randint"]:
        L.oracle_rng_seed(s, case["seed"])
        assert [L.oracle_randint(s, case["lo"], case["hi"]) for _ in range(8)] == case["draws"]
        # np.random.choice(arange(lo, hi)) == lo + randint(0, hi - lo)
        assert [case["lo"] + L.oracle_randint(s, 0, case["hi"] - case["lo"]) for _ in range(4)] == case["choice"]
    L.oracle_rng_free(s)


def test_live_numpy_agrees_when_available():
    """Same check against the numpy of this interpreter (legacy stream is frozen by NEP 19)."""
    L = orc.lib()
    s = L.oracle_rng_new(12345)
    np.random.seed(12345)
    exp = [np.random.binomial(400000, .05), np.random.binomial(20000, .7), np.random.binomial(14000, .4),
           np.random.binomial(20, .3), np.random.poisson(412345.6)]
    got = [L.oracle_binomial(s, 400000, .05), L.oracle_binomial(s, 20000, .7), L.oracle_binomial(s, 14000, .4),
           L.oracle_binomial(s, 20, .3), L.oracle_poisson(s, 412345.6)]
    assert exp == got == [20012, 13977, 5464, 7, 412631]  # SURVEY.md 8(c) legacy-RNG KAT
    L.oracle_rng_free(s)
