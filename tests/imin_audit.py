"""Instruction budget of ONE simulation of K1 (the I_min behind roofline.achieved_algorithmic).

The oracle replays the reference's loop and records every draw a simulation requests (kind, n, p).  Each draw class
is priced with the number of SASS instructions one lane executes for it in k_pcr_ls when it runs alone -- its own
rejection trials only, no idle slots while neighbours retry -- read off the per-instruction counts of
profiles/r02_k1_lockstep_ncu.txt / tools/ncu_sass.py (the instruction stream as compiled, not a lower bound of the
algorithm):

  BTRS draw (n min(p,q) >= 20)   77 set-up + trials x (33 Philox2x32-10 + 4 uniforms + 14 candidate / quick test +
                                 8 accept) + 0.21 x trials x 32 squeeze;   trials = (2.83 + 5.1 / b) / sqrt(2 pi),
                                 b = 1.15 + 2.53 sqrt(n p q)  (Hoermann 1993: hat area / pmf mass)
  inversion draw                 42 set-up + 38 Philox + uniform + 9 x (steps + 1);  steps = E[X] = n min(p,q)
  Poisson draw (capture)         120
  (bin, cycle) visit             60   load, carry, floor test, slip probabilities, three stage selections, stores
  per simulation                 1500 priors (3 Philox4x32-10 blocks), start pool, batch bookkeeping
                                 + 150 per input allele of lnL (25 fp64 Lanczos terms each)

tests/test_instruction_budget.py pins profiles/k1_imin.json to what this returns."""
import math

import numpy as np

from oracle import oracle as orc
from tests.util import CFG1, CFG2, CFG3

CONFIGS = {
    "cfg1": (CFG1, {}, 300),
    "cfg2": (CFG2, {}, 300),
    "cfg3": (CFG3, dict(n_cycles=30, capture_cycle=16), 60),
}
COST = {"btrs_setup": 77, "btrs_trial": 59, "btrs_slow": 32, "btrs_slow_frac": 0.21, "inv_setup": 42, "inv_rng": 38,
        "inv_step": 9, "poisson": 120, "visit": 60, "fixed": 1500, "lnl_per_allele": 150}


def audit(name, n_sims=None, seed=1000):
    spec, kw, n_default = CONFIGS[name]
    reads = orc.parse_reads(spec)
    nc = orc.auto_noise_coef(reads)
    cfg = orc.make_cfg(noise_coef=nc, rng_mode=1, with_readout=False, **kw)
    n_sims = n_sims or n_default
    tot = dict(visits=0.0, btrs=0.0, btrs_trials=0.0, inv=0.0, inv_steps=0.0, poisson=0.0)
    for s in range(n_sims):
        rec, _, sc = orc.record_script(cfg, reads, seed + s, cap=1 << 18)
        tot["visits"] += int(rec["steps_stutter"]) + int(rec["steps_plain"])
        b = sc["kind"] == 1
        n = sc["n"][b].astype(np.float64)
        p = sc["p"][b].astype(np.float64)
        pp = np.minimum(p, 1.0 - p)
        big = n * pp >= 20.0   # SONICS_INV_BELOW (samplers.cuh)
        small = (~big) & (n > 0) & (p > 0) & (p < 1)
        spq = np.sqrt(n[big] * pp[big] * (1.0 - pp[big]))
        tot["btrs"] += big.sum()
        tot["btrs_trials"] += ((2.83 + 5.1 / (1.15 + 2.53 * spq)) / math.sqrt(2.0 * math.pi)).sum()
        tot["inv"] += small.sum()
        tot["inv_steps"] += (n[small] * pp[small]).sum()
        tot["poisson"] += (sc["kind"] == 2).sum()
    per = {k: v / n_sims for k, v in tot.items()}
    c = COST
    n_alleles = int(np.count_nonzero(reads))
    i_min = (per["btrs"] * c["btrs_setup"] + per["btrs_trials"] * (c["btrs_trial"] + c["btrs_slow_frac"] * c["btrs_slow"])
             + per["inv"] * (c["inv_setup"] + c["inv_rng"] + c["inv_step"]) + per["inv_steps"] * c["inv_step"]
             + per["poisson"] * c["poisson"] + per["visits"] * c["visit"] + c["fixed"] + n_alleles * c["lnl_per_allele"])
    per.update(trials_per_btrs_draw=per["btrs_trials"] / max(per["btrs"], 1e-9), input_alleles=n_alleles,
               simulations_sampled=n_sims, I_min_thread_inst_per_reaction=i_min)
    return per
