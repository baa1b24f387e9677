"""Pins oracle_one_repeat()/oracle_simulate() against records produced by the
reference's own compiled one_repeat() (sonics.pyx:293-394, simulate :397-480, rsq
:273-291) under the same np.random.seed -- every field bit for bit."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import CFG1, CFG2, cfg_from_case, load

CASES = load("one_repeat.json")


@pytest.mark.parametrize("k", range(len(CASES)))
def test_one_repeat_bit_exact(k):
    case = CASES[k]
    cfg = cfg_from_case(case["noise_coef"], case["options"], rng_mode=0, with_readout=True)
    recs = orc.run(cfg, orc.parse_reads(case["genotype"]), case["seed"], case["n"])
    for got, exp in zip(recs, case["records"]):
        assert orc.label(got) == exp["label"]
        for f in ("identity", "r_squared", "lnL", "down", "up", "capture", "efficiency"):
            assert float(got[f]).hex() == exp[f], (f, float(got[f]), float.fromhex(exp[f]))
        assert float(np.float32(case["noise_coef"])).hex() == exp["noise_coef"]


def test_less_than_two_alleles_raises():
    cfg = orc.make_cfg()
    with pytest.raises(Exception, match="Less then two alleles"):
        orc.run(cfg, orc.parse_reads("9|100"), 1, 1)


def test_fast_mode_same_distribution():
    """rng_mode=1 drops the per-step self-reseed (sonics.pyx:428,450).  It must be
    statistically indistinguishable from the faithful mode."""
    from scipy.stats import ks_2samp
    gt = "5|1;6|1;7|5;8|11;9|20;10|24;11|2;12|1"
    a = orc.run_parallel(orc.make_cfg(rng_mode=0, with_readout=False, force_second=9), orc.parse_reads(gt), 1, 16000)
    b = orc.run_parallel(orc.make_cfg(rng_mode=1, with_readout=False, force_second=9), orc.parse_reads(gt), 2, 16000)
    assert ks_2samp(a["lnL"], b["lnL"]).pvalue > 0.001
    fa, fb = (a["lnL"] == -999999).mean(), (b["lnL"] == -999999).mean()
    assert abs(fa - fb) < 4 * np.sqrt(0.25 * 0.75 * 2 / 16000)


@pytest.mark.parametrize("kw", [{}, dict(rng_mode=0), dict(floor_opt=-3), dict(floor_opt=-1), dict(random_start=True),
                                dict(noise_coef=0.004), dict(n_cycles=30, capture_cycle=16)])
def test_scripted_draws_reproduce_a_sampled_run(kw):
    """The scripted-draw mode (used by tests/test_gpu_trace.py to replay the kernel's draws): a sampled
    one_repeat() records its draws; replaying the record reproduces the pool and lnL bit for bit, and a
    corrupted record is detected."""
    cfg = orc.make_cfg(**kw)
    for spec in (CFG1, CFG2):
        al = orc.parse_reads(spec)
        nz = np.nonzero(al)[0]
        for seed in range(4):
            rec, pool, script = orc.record_script(cfg, al, seed)
            idx = [int(np.searchsorted(nz, rec["second"]))]
            if cfg.random_start:
                idx = [int(np.searchsorted(nz, rec["first"]))] + idx
            par = (rec["down"], rec["up"], rec["capture"], rec["efficiency"])
            rc, rec2, pool2, used, dp, dlam, _ = orc.replay_script(cfg, al, par, idx, script)
            assert rc == 0 and used == len(script) and (pool == pool2).all() and rec["lnL"] == rec2["lnL"]
            assert dp < 1e-6 and dlam == 0.0  # p stored as float32 in the record
            if len(script) > 10:
                bad = script.copy()
                bad["n"][len(bad) // 2] += 1
                assert orc.replay_script(cfg, al, par, idx, bad)[0] == 3
                assert orc.replay_script(cfg, al, par, idx, script[:-1])[0] == 1
