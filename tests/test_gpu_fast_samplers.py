"""The opt-in approximate sampler (sonics_set_samplers(SONICS_SAMPLERS_FAST, min_npq); include/sonics_b200.h).

Binomial draws with n p q >= min_npq come from one normal deviate through the first Cornish-Fisher term instead of
BTRS.  It is NOT the reference's sampler and not the default; what it has to pass are the north star's own gates --
lnL distributions indistinguishable from the oracle's by KS (p > 0.01 at >= 1e5 simulations per group) and
genotype calls identical on >= 99.5 % of the rows that PASS on both sides -- plus a direct bound on the sampled cdf."""
import os

import numpy as np
import pytest
import scipy.stats as st

from oracle import oracle as orc
from sonics_b200 import engine as eng
from tests.gpu_util import engine
from tests.test_gpu_simulate import _compare
from tests.util import CFG1, CFG2

pytestmark = pytest.mark.gpu
# the tested setting (the header's recommendation); SONICS_FAST_MIN_NPQ overrides it for one-off runs of the KS and
# concordance gates at another threshold (profiles/r02_fast_samplers_npq20.txt)
MIN_NPQ = float(os.environ.get("SONICS_FAST_MIN_NPQ", "100"))


@pytest.fixture()
def fast_engine():
    e = engine()
    e.set_samplers("fast", MIN_NPQ)
    try:
        yield e
    finally:
        e.set_samplers("exact")


@pytest.mark.parametrize("n,p", [(400, 0.5), (1200, 0.1), (150000, 0.05), (3000000, 0.1), (1200000, 0.0123),
                                 (2000000, 6e-5), (40000, 0.97), (11000, 0.0093)])
def test_sampled_cdf_within_the_stated_bound(n, p):
    """2e7 draws of the approximate sampler against the exact binomial cdf: |difference| <= 1.2e-2 / (n p q) plus
    the sampling noise of the empirical cdf (2 / sqrt(N): the 1e-7 quantile of the Kolmogorov statistic)."""
    e = engine()
    N = 20_000_000
    x = e.debug_sample("binomial_fast", n, p, N, seed=99 + n)
    p32 = float(np.float32(p))
    npq = n * p32 * (1 - p32)
    assert npq >= 100 and x.min() >= 0 and x.max() <= n
    lo, hi = int(x.min()), int(x.max())
    emp = np.cumsum(np.bincount(x - lo, minlength=hi - lo + 1)) / N
    d = np.abs(emp - st.binom.cdf(np.arange(lo, hi + 1), n, p32)).max()
    print("n=%d p=%g npq=%.0f max|dCDF|=%.2e (bound %.2e + noise %.1e)" % (n, p, npq, d, 1.2e-2 / npq, 2 / np.sqrt(N)))
    assert d <= 1.2e-2 / npq + 2 / np.sqrt(N), (n, p, d)
    assert abs(x.mean() - n * p32) < 5 * np.sqrt(npq / N) + 1e-3 * np.sqrt(npq)
    assert abs(x.var() / npq - 1) < 2e-3


def test_exact_is_the_default_and_fast_changes_the_draws(fast_engine):
    e = fast_engine
    params = eng.make_params()
    table = e.build_table_from_arrays(params, [orc.parse_reads(CFG2)], [0.0])
    fast = e.simulate(params, table, 4096, seed=7)["lnL"].cpu().numpy()
    again = e.simulate(params, table, 4096, seed=7)["lnL"].cpu().numpy()
    e.set_samplers("exact")
    exact = e.simulate(params, table, 4096, seed=7)["lnL"].cpu().numpy()
    assert (fast == again).all() and not (fast == exact).all()
    with pytest.raises(Exception, match="n p q >= 20"):
        e.set_samplers("fast", 5.0)


def test_ks_cfg1_fast(fast_engine):
    _compare(CFG1, 0.0, 600_000, 1, {}, {})


def test_ks_cfg2_fast(fast_engine):
    _compare(CFG2, 0.0, 1_600_000, 2, {}, {})


def test_ks_noise_case_fast(fast_engine):
    gt = "9|1;10|54;11|914;12|15"
    _compare(gt, orc.auto_noise_coef(orc.parse_reads(gt)), 480_000, 3, {}, {})


def test_concordance_fast(fast_engine):
    """1200 genotypes of the cfg4 generator through the whole loop (rounds, statistics, K4 replay) in fast mode:
    genotype identical on >= 99.5 % of the rows that PASS on both sides, (genotype, FILTER) agreement within
    sampling error of the oracle's agreement with itself."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
    import synth_vcf
    from sonics_b200 import driver
    from sonics_b200.sonics import filter_string
    from tests.test_gpu_cli import _oracle_pool
    e = fast_engine
    loci = synth_vcf.make_genotypes(150, 8, seed=20261019, cov=(50, 1000))
    pre = {"one_allele_threshold": 45, "noise_threshold": 0.05}
    work = [p for k, p in (driver.prefilter(t, pre, {"monte_carlo": False}) for t in synth_vcf.genotype_strings(loci))
            if k == "sim"]
    off, al, rd = eng.to_csr([w[0] for w in work])
    v = e.genotype_batch(eng.make_params(), off, al, rd, [w[1] for w in work], 1000, seed=778)
    gpu = [("%d/%d" % (x["allele_lo"], x["allele_hi"]) if x["allele_lo"] >= 0 else "./.", filter_string(int(x["filter"])))
           for x in v]
    oa = _oracle_pool([(w[0], w[1], 100 + i) for i, w in enumerate(work)])
    ob = _oracle_pool([(w[0], w[1], 900000 + i) for i, w in enumerate(work)])

    def agree(x, y):
        both = float(np.mean([a == b for a, b in zip(x, y)]))
        pp = [(a, b) for a, b in zip(x, y) if a[1] == "PASS" and b[1] == "PASS"]
        return both, float(np.mean([a[0] == b[0] for a, b in pp])) if pp else 1.0, len(pp)
    g_o, o_o = agree(gpu, oa), agree(oa, ob)
    n = len(work)
    print("fast n=%d device-vs-oracle: geno+filter %.4f geno|PASS-both %.4f (%d rows); oracle-vs-oracle %.4f %.4f (%d)"
          % ((n,) + g_o + o_o))
    assert g_o[1] >= 0.995 and g_o[2] >= 0.5 * n
    se = np.sqrt(2 * max(o_o[0] * (1 - o_o[0]), 1e-4) / n)
    assert g_o[0] >= o_o[0] - 4 * se - 0.005, (g_o, o_o)


def test_cli_flag(tmp_path):
    """`sonics --fast_samplers [MIN_NPQ]`: same row layout, the README example still calls 9/10 PASS; the flag changes
    the draws (another lnL than the default mode under the same seed) and is rejected below n p q = 20."""
    import subprocess
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

    def run(*extra):
        return subprocess.run([sys.executable, os.path.join(root, "bin", "sonics"), CFG1, "-o", str(tmp_path),
                               "--seed", "5"] + list(extra), capture_output=True, text=True, timeout=600)
    rows = {}
    for name, extra in (("exact", []), ("fast", ["--fast_samplers"]), ("fast50", ["--fast_samplers", "50"])):
        out = run("-n", name + ".txt", *extra)
        assert out.returncode == 0, out.stderr[-2000:]
        rows[name] = open(os.path.join(str(tmp_path), name + ".txt")).read().split("\n")[1].split("\t")
        assert rows[name][3] == "9/10" and rows[name][7] == "PASS" and len(rows[name]) == 13
    assert rows["fast"][6] != rows["exact"][6] and rows["fast50"][6] != rows["fast"][6]
    bad = run("-n", "bad.txt", "--fast_samplers", "5")
    assert bad.returncode != 0 and "n p q >= 20" in bad.stderr
