"""The two forms of K1 run the same simulations.

k_pcr_ls (sort by priors + 32 similar simulations per warp in lockstep, lnL fused; sim_lockstep.cuh) and k_pcr +
k_lnl (persistent lanes, sim_kernel.cuh) must request the same draws in the same order from the same Philox
streams: every output -- lnL, group, identity, r_squared, the four prior parameters -- has to be identical bit for
bit.  The persistent form is the one the draw-trace replay (test_gpu_trace.py) pins against the reference's loop,
so this equality carries that proof over to the production (lockstep) form for every configuration below."""
import numpy as np
import pytest

from oracle import oracle as orc
from sonics_b200 import engine as eng
from tests.gpu_util import engine
from tests.util import CFG1, CFG2, CFG3

pytestmark = pytest.mark.gpu


def _both(params, genotypes, noise, n_sims, seed, readout=True, uid=None, active=None, sim_begin=0):
    e = engine()
    off, al, rd = eng.to_csr([orc.parse_reads(g) for g in genotypes])
    table = e.build_table(params, off, al, rd, noise, uid)
    res = {}
    try:
        for k in ("persistent", "lockstep"):
            e.set_kernel(k)
            act = None if active is None else e.torch.tensor(active, dtype=e.torch.int32, device=e.device)
            out = e.simulate(params, table, n_sims, seed, readout=readout, simpar=True, active=act,
                             sim_begin=sim_begin, out_offset=0)
            e.sync()
            res[k] = {n: v.cpu().numpy() for n, v in out.items()}
    finally:
        e.set_kernel("auto")
    return res["persistent"], res["lockstep"], table


def _assert_same(a, b, rows=None):
    for n in a:
        x, y = a[n], b[n]
        if rows is not None:
            x, y = x[rows], y[rows]
        if x.dtype.kind == "f":
            assert (x.view(np.uint8) == y.view(np.uint8)).all(), n   # bitwise, NaN-safe
        else:
            assert (x == y).all(), n


def test_default_configs_bit_identical():
    p = eng.make_params()
    a, b, _ = _both(p, [CFG1, CFG2, "3|40;4|25;7|2", "9|1;10|54;11|914;12|15"], [0.0, 0.0, 0.0, 0.0011], 6000, seed=5)
    _assert_same(a, b)
    assert (a["lnL"] > -999999).mean() > 0.3


def test_cfg3_wide_window_noise_bit_identical():
    nc = float(np.float32(orc.auto_noise_coef(orc.parse_reads(CFG3))))
    p = eng.make_params(n_cycles=30, capture_cycle=16)
    a, b, _ = _both(p, [CFG3], [nc], 5000, seed=6)
    _assert_same(a, b)


@pytest.mark.parametrize("kw", [dict(floor=-1), dict(floor=0), dict(floor=-3), dict(floor=2),
                                dict(random=True), dict(up_preference=True),
                                dict(capture_cycle=1), dict(capture_cycle=24), dict(capture_cycle=40),
                                dict(n_cycles=3, capture_cycle=2), dict(n_cycles=0, capture_cycle=1),
                                dict(start_copies=1000), dict(start_copies=7),
                                dict(efficiency=(0.9, 1.0), n_cycles=8, capture_cycle=4, start_copies=1000),
                                dict(down=(0, 0.5), up=(0, 0.5)), dict(capture=(-1.9, -1.5))])
def test_option_matrix_bit_identical(kw):
    p = eng.make_params(**kw)
    a, b, _ = _both(p, [CFG1, "8|50;9|3;14|20", "2|9;3|1"], [0.0, 0.004, 0.0], 1500, seed=11)
    _assert_same(a, b)


def test_ragged_counts_active_subset_and_offsets():
    p = eng.make_params()
    gts = [CFG1, CFG2, "9|60;10|40;11|30;12|8", "14|700;15|2100;16|1900;17|2500;18|900;19|300", "9|100;10|100"]
    # 77 simulations per genotype (not a multiple of 32), genotypes 4, 0, 2 only, simulation indices from 1000
    a, b, _ = _both(p, gts, [0.0] * 5, 77, seed=9, uid=[7, 8, 9, 10, 11], active=[4, 0, 2], sim_begin=1000)
    _assert_same(a, b, rows=[4, 0, 2])


def test_chunked_lockstep_scratch():
    """A scratch that holds only part of the work units: the lockstep form sorts and runs chunk by chunk."""
    e = engine()
    p = eng.make_params()
    table = e.build_table_from_arrays(p, [orc.parse_reads(CFG2)], [0.0])
    n = 20000
    full = e.simulate(p, table, n, seed=3)
    e.sync()
    keep, cap = e._scratch_buf, e.SCRATCH_CAP
    small = int(e.lib.sonics_scratch_bytes(3000, table.max_window))
    assert small < int(e.lib.sonics_scratch_bytes(n, table.max_window))
    try:
        e._scratch_buf = e.torch.empty(small, dtype=e.torch.uint8, device=e.device)
        e.SCRATCH_CAP = small
        part = e.simulate(p, table, n, seed=3)
        e.sync()
    finally:
        e.SCRATCH_CAP = cap
        e._scratch_buf = keep
    assert (full["lnL"] == part["lnL"]).all() and (full["group"] == part["group"]).all()


def test_sort_key_bits_do_not_change_results():
    """The sort only decides which lane runs which simulation."""
    from sonics_b200._abi import check
    e = engine()
    p = eng.make_params()
    table = e.build_table_from_arrays(p, [orc.parse_reads(CFG2), orc.parse_reads(CFG1)], [0.0, 0.0])
    ref = None
    try:
        # the default key (start pair | noise reach | Morton-interleaved priors), then the genotype-major key with
        # different bit fields
        for mode, bits in ((0, (-1, -1, -1)), (1, (-1, -1, -1)), (1, (0, 0, 0)), (1, (3, 2, 1)), (1, (7, 0, 0)),
                           (1, (2, 4, 4)), (1, (15, 15, 15))):
            check(e.lib.sonics_set_sort_mode(e.ctx, mode))
            check(e.lib.sonics_set_sort_bits(e.ctx, *bits))
            out = e.simulate(p, table, 30000, seed=12, readout=True)
            e.sync()
            got = {k: v.cpu().numpy() for k, v in out.items()}
            if ref is None:
                ref = got
            _assert_same(ref, got)
    finally:
        check(e.lib.sonics_set_sort_bits(e.ctx, -1, -1, -1))
        check(e.lib.sonics_set_sort_mode(e.ctx, 0))


def test_pair_major_sort_pools_start_pairs_across_genotypes():
    """Launches over many genotypes sort by (noise flag, the two start alleles as absolute repeat counts, priors):
    a warp then holds simulations of different genotypes -- different window origins, reads, allele lists, some
    with noise molecules.  Same results as the genotype-major order and as the persistent form."""
    import sys
    import os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
    import synth_vcf
    from sonics_b200 import driver
    from sonics_b200._abi import check
    loci = synth_vcf.make_genotypes(40, 4, seed=77, cov=(60, 900))
    pre = {"one_allele_threshold": 45, "noise_threshold": 0.05}
    work = [pl for k, pl in (driver.prefilter(t, pre, {"monte_carlo": False}) for t in synth_vcf.genotype_strings(loci))
            if k == "sim"]
    reads = [w[0] for w in work] + [orc.parse_reads("9|1;10|54;11|914;12|15"), orc.parse_reads("30|5;31|400;32|9")]
    noise = [float(np.float32(w[1])) for w in work] + [0.0011, 0.004]
    assert len(reads) > 100 and any(n > 0 for n in noise)
    e = engine()
    p = eng.make_params()
    table = e.build_table_from_arrays(p, reads, noise)
    res = {}
    try:
        for name, kernel, mode in (("persistent", "persistent", 0), ("pair", "lockstep", 0), ("slot", "lockstep", 1)):
            e.set_kernel(kernel)
            check(e.lib.sonics_set_sort_mode(e.ctx, mode))
            out = e.simulate(p, table, 100, seed=21, readout=True, simpar=True)
            e.sync()
            res[name] = {k: v.cpu().numpy() for k, v in out.items()}
    finally:
        e.set_kernel("auto")
        check(e.lib.sonics_set_sort_mode(e.ctx, 0))
    _assert_same(res["persistent"], res["pair"])
    _assert_same(res["persistent"], res["slot"])
