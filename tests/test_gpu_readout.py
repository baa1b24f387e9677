"""GPU parity of row a6: the simulated sequencing readout and its descriptors (identity, r_squared).

Reference: sonics.pyx:341-357 (stage 1: n_i ~ Binomial(D, pool_i / T) per bin), :365-372
(readout = bincount(choice(mid, D))), :373-384 (identity), rsq() :273-291, C-float store :300.

Two halves, like the lnL gates:
  * deterministic -- the readout vector K1 drew is exported (sonics_debug_readout) and fed to the oracle's
    restatement of identity / rsq(), which is itself pinned to the reference's own rsq() on 400 vectors
    (tests/golden/rsq.json): the device descriptors must be identical bit for bit, incl. the ss_tot == 0 -> 1e-16
    guard, the empty-readout branch (identity 0, r_squared -999999) and the float32 store;
  * distributional -- the device draws the readout as Binomial(D, pool_i/T) per bin followed by a chain of
    conditional binomials instead of choice() + bincount: per start-allele group, identity and r_squared of
    >= 1e5 device simulations against >= 1e5 oracle one_repeat(with_readout) records by two-sample KS (p > 0.01),
    the mean readout profile, the empty-readout frequency on a tiny-D genotype, and the --monte_carlo mode's
    75th percentiles against the oracle's monte_carlo() restatement."""
import os

import numpy as np
import pytest
import scipy.stats as st

from oracle import oracle as orc
from oracle import stats
from sonics_b200 import engine as eng
from tests.gpu_util import engine
from tests.test_gpu_trace import window
from tests.util import CFG1, CFG2, CFG3

pytestmark = pytest.mark.gpu

N_ORACLE = int(os.environ.get("SONICS_KS_ORACLE", "100000"))


def _dump_case(gt, noise, n_sims, seed, **kw):
    e = engine()
    params = eng.make_params(**kw)
    reads = orc.parse_reads(gt)
    table = e.build_table_from_arrays(params, [reads], [noise])
    out, ro = e.simulate_with_readout_dump(params, table, n_sims, seed)
    ident = out["identity"][0].cpu().numpy()
    r2 = out["rsq"][0].cpu().numpy()
    lo, W = window(reads, kw)
    assert W == table.max_window
    return reads, lo, W, ro[0], ident, r2


@pytest.mark.parametrize("gt,noise,kw", [
    (CFG1, 0.0, {}),
    (CFG2, 0.0, {}),
    (CFG3, "auto", dict(n_cycles=30, capture_cycle=16)),
    ("9|1;10|1", 0.0, {}),          # tiny D: empty readouts (identity 0, r_squared -999999)
    ("9|3;10|3", 0.0, {}),          # flat reads: ss_tot == 0 whenever the readout stays on the two alleles
    ("9|100;10|100", 0.0, {}),
    ("8|50;9|3;14|20", 0.004, dict(floor=0)),
])
def test_descriptors_bit_exact_on_exported_readouts(gt, noise, kw):
    if noise == "auto":
        noise = float(np.float32(orc.auto_noise_coef(orc.parse_reads(gt))))
    n_sims = 3000
    reads, lo, W, ro, ident, r2 = _dump_case(gt, noise, n_sims, 20261020, **kw)
    D = int(reads.sum())
    n_empty = n_guard = 0
    for j in range(n_sims):
        readout = np.zeros(500, dtype=np.int64)
        readout[lo:lo + W] = ro[j]
        tot = int(readout.sum())
        assert tot in (0, D), (j, tot, D)  # choice(mid, D) draws D reads, or nothing was sequenced
        oi, orsq = orc.descriptors(reads, readout)
        assert ident[j] == oi, (j, ident[j], oi)
        assert np.float32(r2[j]) == np.float32(orsq), (j, r2[j], orsq)
        n_empty += tot == 0
        n_guard += abs(orsq) > 1e12
    if gt == "9|1;10|1":
        assert n_empty > 20, n_empty           # the branch is exercised
    if gt == "9|3;10|3":
        assert n_guard > 20, n_guard           # so is the 1e-16 guard


def _oracle_by_group(gt, noise, a, n, okw):
    cfg = orc.make_cfg(noise_coef=noise, rng_mode=1, with_readout=True, force_second=int(a), **okw)
    return orc.run_parallel(cfg, orc.parse_reads(gt), 5000 + int(a), n)


def _ks_case(gt, noise, n_gpu, seed, okw, pkw, groups=None, n_oracle=None):
    e = engine()
    params = eng.make_params(**pkw)
    reads = orc.parse_reads(gt)
    table = e.build_table_from_arrays(params, [reads], [noise])
    out = e.simulate(params, table, n_gpu, seed, readout=True)
    e.sync()
    grp = out["group"][0].cpu().numpy().view(np.uint16)
    ident = out["identity"][0].cpu().numpy()
    r2 = out["rsq"][0].cpu().numpy().astype(np.float64)
    nz = np.nonzero(reads)[0]
    A = len(nz)
    first = int(np.argmax(reads))
    fi = list(nz).index(first)
    rep = []
    for a in (groups or nz):
        ai = list(nz).index(a)
        gid = min(fi, ai) * A + max(fi, ai)
        m = grp == gid
        o = _oracle_by_group(gt, noise, a, n_oracle or N_ORACLE, okw)
        assert m.sum() >= 100000 and len(o) >= min(100000, n_oracle or N_ORACLE)
        p_i = st.ks_2samp(ident[m], o["identity"]).pvalue
        p_r = st.ks_2samp(r2[m], o["r_squared"]).pvalue
        rep.append((a, int(m.sum()), len(o), ident[m].mean(), o["identity"].mean(), p_i, np.median(r2[m]),
                    np.median(o["r_squared"]), p_r))
    for r in rep:
        print("second=%d n_gpu=%d n_oracle=%d identity mean %.5f / %.5f KS p=%.3f | r2 median %.5f / %.5f KS p=%.3f" % r)
    for r in rep:
        assert r[5] > 0.01, ("identity", r)
        assert r[8] > 0.01, ("r_squared", r)


def test_ks_identity_rsq_cfg1():
    _ks_case(CFG1, 0.0, 400_000, 21, {}, {})


def test_ks_identity_rsq_cfg2():
    _ks_case(CFG2, 0.0, 1_000_000, 22, {}, {})


def test_ks_identity_rsq_noise_case():
    gt = "9|1;10|54;11|914;12|15"
    nc = orc.auto_noise_coef(orc.parse_reads(gt))
    _ks_case(gt, nc, 500_000, 23, {}, {})


def test_ks_identity_rsq_cfg3():
    nc = orc.auto_noise_coef(orc.parse_reads(CFG3))
    kw = dict(n_cycles=30, capture_cycle=16)
    _ks_case(CFG3, nc, 4_400_000, 24, kw, kw, groups=[28, 36], n_oracle=int(os.environ.get("SONICS_KS_ORACLE_CFG3", "100000")))


def test_mean_readout_profile_matches_oracle_pool_profile():
    """E[readout_i] = D * E[pool_i / T | something sequenced]: the per-bin means of 2e5 device readouts against
    the oracle's final pools (the chain of conditional binomials is a multinomial over n_i / N)."""
    gt = CFG2
    reads, lo, W, ro, ident, r2 = _dump_case(gt, 0.0, 200_000, 31)
    D = int(reads.sum())
    full = ro[ro.sum(axis=1) == D]
    got = full.mean(axis=0) / D
    cfg = orc.make_cfg(rng_mode=1, with_readout=False)
    # pools are not part of the oracle's record: use its single-simulation entry for a sample of pools
    lib = orc.lib()
    import ctypes
    rng = lib.oracle_rng_new(12345)
    prof = np.zeros(500)
    n = 4000
    rec = np.zeros(1, dtype=orc.RECORD_DTYPE)
    pool = np.zeros(500, dtype=np.int64)
    al = np.ascontiguousarray(reads, dtype=np.int64)
    for _ in range(n):
        lib.oracle_one_repeat(rng, ctypes.byref(cfg), al.ctypes.data, 500, rec.ctypes.data, pool.ctypes.data)
        prof += pool / pool.sum()
    lib.oracle_rng_free(rng)
    prof /= n
    exp = prof[lo:lo + W]
    # standard error of the oracle profile dominates (4000 pools, between-simulation spread ~ 0.1 per bin)
    assert np.abs(got - exp).max() < 0.012, (got, exp)


def test_empty_readout_frequency_on_tiny_genotype():
    gt = "9|1;10|1"
    e = engine()
    params = eng.make_params()
    reads = orc.parse_reads(gt)
    table = e.build_table_from_arrays(params, [reads], [0.0])
    n = 400_000
    out = e.simulate(params, table, n, 41, readout=True)
    e.sync()
    r2 = out["rsq"][0].cpu().numpy()
    ident = out["identity"][0].cpu().numpy()
    empty = r2 == np.float32(-999999.0)
    assert (ident[empty] == 0).all()
    o = orc.run_parallel(orc.make_cfg(rng_mode=1, with_readout=True), reads, 42, 200_000)
    oe = o["r_squared"] == float(np.float32(-999999.0))
    assert (o["identity"][oe] == 0).all()
    pg, po = empty.mean(), oe.mean()
    pooled = (empty.sum() + oe.sum()) / (len(empty) + len(oe))
    z = (pg - po) / np.sqrt(pooled * (1 - pooled) * (1 / len(empty) + 1 / len(oe)))
    print("empty-readout fraction gpu %.5f oracle %.5f z=%+.2f" % (pg, po, z))
    assert po > 0.01 and abs(z) < 4.0, (pg, po, z)
    # identity takes the values k/2 only; compare the whole pmf
    tg = np.array([(ident == v).sum() for v in (0.0, 0.5, 1.0)])
    to = np.array([(o["identity"] == v).sum() for v in (0.0, 0.5, 1.0)])
    assert tg.sum() == n and to.sum() == len(o)
    assert st.chi2_contingency(np.vstack([tg, to]))[1] > 1e-3, (tg, to)


def test_monte_carlo_mode_q75_columns_match_oracle():
    """--monte_carlo rows (sonics.pyx:115-134): identity, identity_q75, r_squared, r_squared_q75, lnL, lnL_q75 of R
    independent runs (-r 1000) on the device against R runs of the oracle's one_repeat(with_readout) through the
    oracle's monte_carlo() restatement."""
    e = engine()
    gt = "9|60;10|40;11|30;12|8"
    reads = orc.parse_reads(gt)
    R, reps = 300, 1000
    params = eng.make_params(monte_carlo=True)
    off, al, rd = eng.to_csr([reads] * R)
    v = e.genotype_batch(params, off, al, rd, [0.0] * R, reps, seed=77, uid=np.arange(R))
    assert (v["run_reps"] == reps).all()
    cfg = orc.make_cfg(rng_mode=1, with_readout=True)
    rows = []
    nz = np.nonzero(reads)[0]
    for r in range(R):
        recs = orc.run(cfg, reads, 9000 + r, reps)
        lab = np.array([orc.label(x) for x in recs])
        it = iter([[dict(identity=x["identity"], r_squared=x["r_squared"], lnL=x["lnL"], label=l)
                    for x, l in zip(recs, lab)]])
        row, info = stats.monte_carlo(reps, lambda n: next(it), monte_carlo_mode=True)
        f = row.split("\t")
        rows.append((f[0], float(f[1]), float(f[2]), float(f[3]), float(f[4]), float(f[5]), float(f[6])))
    o_gt = np.array([x[0] for x in rows])
    g_gt = np.array(["%d/%d" % (a, b) for a, b in zip(v["allele_lo"], v["allele_hi"])])
    # same modal genotype, similar frequency
    vals, cnt = np.unique(o_gt, return_counts=True)
    top = vals[np.argmax(cnt)]
    assert abs((g_gt == top).mean() - (o_gt == top).mean()) < 0.12, (top, (g_gt == top).mean(), (o_gt == top).mean())
    cols = {"identity": (v["identity"], 1), "identity_q75": (v["identity_q75"], 2), "r_squared": (v["r_squared"], 3),
            "r_squared_q75": (v["r_squared_q75"], 4), "lnL": (v["lnL"], 5), "lnL_q75": (v["lnL_q75"], 6)}
    for name, (g, k) in cols.items():
        o = np.array([x[k] for x in rows])
        p = st.ks_2samp(g, o).pvalue
        print("%-14s gpu median %.5f oracle median %.5f KS p=%.3f" % (name, np.median(g), np.median(o), p))
        assert p > 0.01, (name, p)
    assert len(nz) == 4
