"""Helpers shared by the -m gpu tests."""
import numpy as np

from oracle import oracle as orc
from sonics_b200 import engine as eng

_ENGINE = None


def engine():
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = eng.Engine(0)
    return _ENGINE


def philox4x32_10(ctr, key):
    """Pure-Python Philox4x32-10 (Salmon et al. SC'11) used to check the device stream."""
    c = list(ctr)
    k = list(key)
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF,
             p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def philox2x32_10(ctr, key):
    """Pure-Python Philox2x32-10 (same paper): the stream of the PCR cycles' draws."""
    c0, c1 = ctr
    k = key
    for _ in range(10):
        p = 0xD256D193 * c0
        c0, c1 = ((p >> 32) ^ k ^ c1) & 0xFFFFFFFF, p & 0xFFFFFFFF
        k = (k + 0x9E3779B9) & 0xFFFFFFFF
    return [c0, c1]


def gpu_lnl_by_group(e, genotype, n_sims, seed, noise_coef=0.0, **pkw):
    """Run n_sims simulations of one genotype; returns {label: lnL array}, table."""
    params = eng.make_params(**pkw)
    reads = orc.parse_reads(genotype)
    table = e.build_table_from_arrays(params, [reads], [noise_coef])
    out = e.simulate(params, table, n_sims, seed)
    e.sync()
    lnL = out["lnL"][0].cpu().numpy()
    grp = out["group"][0].cpu().numpy().view(np.uint16)
    res = {}
    for gid in np.unique(grp):
        res[table.group_label(0, int(gid))] = lnL[grp == gid]
    return res, table
