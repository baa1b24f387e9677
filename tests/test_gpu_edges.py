"""GPU edge cases of the simulate path: ragged sizes, chunked scratch, wide windows, ABI errors."""
import numpy as np
import pytest
import scipy.stats as st

from oracle import oracle as orc
from sonics_b200 import engine as eng
from sonics_b200._abi import SonicsError
from tests.gpu_util import engine
from tests.util import CFG1, CFG2

pytestmark = pytest.mark.gpu


def test_prefix_property_and_ragged_counts():
    """Simulation j depends only on (seed, uid, j): any count / split / genotype mix gives the same values."""
    e = engine()
    params = eng.make_params()
    gts = [orc.parse_reads(CFG1), orc.parse_reads(CFG2), orc.parse_reads("3|50;4|40")]
    table = e.build_table_from_arrays(params, gts, [0.0, 0.0, 0.0], uid=[7, 8, 9])
    big = e.simulate(params, table, 1000, seed=3)
    e.sync()
    for n in (1, 31, 32, 33, 257):
        small = e.simulate(params, table, n, seed=3)
        e.sync()
        assert (small["lnL"] == big["lnL"][:, :n]).all() and (small["group"] == big["group"][:, :n]).all()
    # a window in the middle, written at offset 0
    mid = e.simulate(params, table, 100, seed=3, sim_begin=450, out_offset=0)
    e.sync()
    assert (mid["lnL"] == big["lnL"][:, 450:550]).all()
    # an active list picks genotypes; the others stay untouched
    t = e.torch
    out = {"lnL": t.full((3, 64), 7.0, dtype=t.float64, device=e.device),
           "group": t.zeros((3, 64), dtype=t.int16, device=e.device)}
    e.simulate(params, table, 64, seed=3, out=out, active=t.tensor([2, 0], dtype=t.int32, device=e.device))
    e.sync()
    assert (out["lnL"][1] == 7.0).all()
    assert (out["lnL"][0] == big["lnL"][0, :64]).all() and (out["lnL"][2] == big["lnL"][2, :64]).all()


def test_chunked_scratch_gives_identical_results():
    e = engine()
    params = eng.make_params()
    table = e.build_table_from_arrays(params, [orc.parse_reads(CFG2)] * 5, [0.0] * 5, uid=np.arange(5))
    a = e.simulate(params, table, 3000, seed=12, readout=True)
    e.sync()
    ref = {k: v.clone() for k, v in a.items()}
    cap, buf = e.SCRATCH_CAP, e._scratch_buf
    try:
        e.SCRATCH_CAP = 1 << 20  # 1 MiB: 15000 units no longer fit one chunk
        e._scratch_buf = None
        b = e.simulate(params, table, 3000, seed=12, readout=True)
        e.sync()
        assert e._scratch_buf.numel() <= (1 << 20)
        for k in ref:
            assert (b[k] == ref[k]).all(), k
    finally:
        e.SCRATCH_CAP, e._scratch_buf = cap, buf


def test_wide_window_and_high_alleles():
    """Alleles near the top of the 500-bin histogram with many cycles: window of > 100 bins."""
    e = engine()
    gt = "380|40;381|300;382|260;383|30"
    okw = dict(n_cycles=50, capture_cycle=20, floor_opt=-1, efficiency=(0.001, 0.05))
    params = eng.make_params(n_cycles=50, capture_cycle=20, floor=-1, efficiency=(0.001, 0.05))
    reads = orc.parse_reads(gt)
    table = e.build_table_from_arrays(params, [reads], [0.0])
    assert table.max_window >= 100
    out = e.simulate(params, table, 20000, seed=4)
    e.sync()
    lnL = out["lnL"][0].cpu().numpy()
    grp = out["group"][0].cpu().numpy().view(np.uint16)
    sel = lnL[grp == 1 * 4 + 2]  # 381/382
    o = orc.run_parallel(orc.make_cfg(rng_mode=1, with_readout=False, force_second=382, **okw), reads, 9, 3000)["lnL"]
    assert st.ks_2samp(sel, o).pvalue > 1e-3
    assert abs((sel == -999999).mean() - (o == -999999).mean()) < 0.05


def test_alleles_near_the_end_of_the_histogram_are_simulated():
    """max allele + n_cycles >= 500: the window is clamped at bin 499 (the reference only fails if a molecule
    actually stutters out of bin 499, which 18 consecutive up-stutters from 481 never do); the genotype is
    simulated normally and does not fail the batch it is part of."""
    e = engine()
    params = eng.make_params()
    gt = "480|10;481|20"
    reads = orc.parse_reads(gt)
    table = e.build_table_from_arrays(params, [orc.parse_reads(CFG1), reads], [0.0, 0.0])
    out = e.simulate(params, table, 30000, seed=14)
    e.sync()
    lnL = out["lnL"][1].cpu().numpy()
    grp = out["group"][1].cpu().numpy().view(np.uint16)
    sel = lnL[grp == 0 * 2 + 1]  # 480/481
    o = orc.run_parallel(orc.make_cfg(rng_mode=1, with_readout=False, force_second=480), reads, 15, 8000)["lnL"]
    assert st.ks_2samp(sel, o).pvalue > 1e-3
    assert abs((sel == -999999).mean() - (o == -999999).mean()) < 0.03


def test_abi_error_paths():
    e = engine()
    params = eng.make_params()
    with pytest.raises(SonicsError, match="Less then two alleles"):
        e.build_table_from_arrays(params, [orc.parse_reads("9|10")], [0.0])
    with pytest.raises(SonicsError, match="overflows int32"):
        e.build_table_from_arrays(eng.make_params(start_copies=2_000_000_000), [orc.parse_reads(CFG1)], [0.0])
    with pytest.raises(SonicsError, match="strictly ascending"):
        e.build_table(params, [0, 2], [10, 9], [5, 5])
    table = e.build_table_from_arrays(params, [orc.parse_reads(CFG1)], [0.0])
    t = e.torch
    small = {"lnL": t.empty((1, 10), dtype=t.float64, device=e.device),
             "group": t.empty((1, 10), dtype=t.int16, device=e.device)}
    with pytest.raises(SonicsError, match="out_stride"):
        e.simulate(params, table, 11, seed=1, out=small)
    other = eng.Table(t.empty(4096, dtype=t.uint8, device=e.device), 1, table.allele_off, table.allele, table.reads, 40)
    with pytest.raises(SonicsError, match="unknown genotype table"):
        e.simulate(params, other, 8, seed=1)
    with pytest.raises(SonicsError, match="at most 8"):
        bad = eng.make_params()
        bad.n_add_ratios = 9
        out = e.simulate(params, table, 100, seed=1)
        e.evaluate(bad, table, out, 100)
