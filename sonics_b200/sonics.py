"""Drop-in mirror of the reference's extension module `sonics` (sonics.pyx).

The reference CLI does `import sonics` and calls sonics.get_alleles() and
sonics.monte_carlo() (sonics:150,350,427-432).  This module exposes the same names
with the same signatures and return formats; the simulation, the lnL and the
statistics run on the GPU behind include/sonics_b200.h.  There is no CPU fallback.

Added on top of the reference interface: monte_carlo_batch(), which is what the
batched driver calls instead of mapping monte_carlo() over genotypes
(sonics:216-228), and seed(): the reference is irreproducible by design (it
reseeds numpy's global RNG from itself, sonics.pyx:428,450); here a run is
reproducible when a seed is set and draws fresh entropy otherwise.
"""
import os

import numpy as np

DTYPE = np.int64

_ENGINES = {}
_SEED = None
_CALLS = 0


def get_alleles(genot_input):
    """'allele1|#;allele2|#' -> (reads per repeat count int64[500], 500, n_alleles).

    Same contract as get_alleles() at sonics.pyx:17-31: empty fields are skipped,
    alleles <= 0 are ignored, an allele >= 500 raises IndexError."""
    max_allele = 500
    alleles = np.zeros(max_allele, dtype=DTYPE)
    for field in genot_input.split(";"):
        if field != "":
            pair = [int(x) for x in field.split("|")]
            if pair[0] > 0:
                alleles[pair[0]] = pair[1]
    n_alleles = len(alleles.nonzero()[0])
    return alleles, max_allele, n_alleles


def seed(value=None):
    """Fix the Philox seed of subsequent calls (None: fresh entropy, the reference's behaviour)."""
    global _SEED, _CALLS
    _SEED = None if value is None else int(value) & 0xFFFFFFFFFFFFFFFF
    _CALLS = 0


def _next_seed():
    global _CALLS
    if _SEED is None:
        env = os.environ.get("SONICS_SEED")
        base = int(env) if env else int.from_bytes(os.urandom(8), "little")
    else:
        base = _SEED
    s = (base + 0x9E3779B97F4A7C15 * _CALLS) & 0xFFFFFFFFFFFFFFFF
    _CALLS += 1
    return s


def engine(device=0):
    """The process-wide Engine of a device (created on first use; raises without a GPU)."""
    from . import engine as eng
    if device not in _ENGINES:
        _ENGINES[device] = eng.Engine(device)
    return _ENGINES[device]


def _fmt(v):
    """str() as the reference applies it when joining the row (sonics.pyx:132,267)."""
    if isinstance(v, (float, np.floating)):
        return repr(float(v))
    return str(v)


def filter_string(bits):
    from ._abi import (FILTER_BEST_RATIO, FILTER_MWU_TEST, FILTER_NO_SUCCESS, FILTER_PASS, FILTER_PERCENTILE_RATIO)
    if bits & FILTER_PASS:
        return "PASS"
    if bits & FILTER_NO_SUCCESS:
        return "no_success"
    conds = []
    if bits & FILTER_MWU_TEST:
        conds.append("MWU_test")
    if bits & FILTER_BEST_RATIO:
        conds.append("best_ratio")
    if bits & FILTER_PERCENTILE_RATIO:
        conds.append("percentile_ratio")
    return ",".join(conds)


def format_row(v, options):
    """One SonicsVerdict record -> the tab-joined row monte_carlo() returns (sonics.pyx:122-134, 208-269)."""
    from ._abi import FILTER_NO_SUCCESS
    if options.get("monte_carlo"):
        ret = ["%d/%d" % (v["allele_lo"], v["allele_hi"]), float(v["identity"]), float(v["identity_q75"]),
               float(v["r_squared"]), float(v["r_squared_q75"]), float(v["lnL"]), float(v["lnL_q75"]),
               int(v["run_reps"])]
        return "\t".join(_fmt(e) for e in ret)
    if v["filter"] & FILTER_NO_SUCCESS:
        return "\t".join(["./.", ".", ".", ".", "no_success", ".", ".", ".", str(int(v["run_reps"])), "."])
    add_data = "."
    add = options.get("add_ratios", "")
    if add != "":
        parts = ["{}|{}".format(perc, _fmt(float(v["add_ratio"][i]))) for i, perc in enumerate(add.split(";"))]
        add_data = ";".join(parts)
    high = 1 if v["p_clamped"] else float(v["high_pval"])
    ret = ["%d/%d" % (v["allele_lo"], v["allele_hi"]), float(v["identity"]), float(v["r_squared"]), float(v["lnL"]),
           filter_string(int(v["filter"])), high, float(v["best_lnL_ratio"]), float(v["percentile_lnL_ratio"]),
           int(v["run_reps"]), add_data]
    return "\t".join(_fmt(e) for e in ret)


def monte_carlo_batch(genotypes, noise_coefs, max_n_reps, constants, ranges, options, seed_value=None, device=0,
                      uids=None):
    """monte_carlo() for many genotypes at once: list of reads arrays (get_alleles()[0]) + noise_coefs.

    Returns (rows, verdicts): the row strings in input order and the structured verdict array."""
    from . import engine as eng
    e = engine(device)
    params = eng.params_from_dicts(constants, ranges, options)
    off, al, rd = eng.to_csr(genotypes)
    s = _next_seed() if seed_value is None else int(seed_value)
    verdicts = e.genotype_batch(params, off, al, rd, np.asarray(noise_coefs, dtype=np.float32), int(max_n_reps), s,
                                uid=uids)
    return [format_row(v, options) for v in verdicts], verdicts


def monte_carlo(max_n_reps, constants, ranges, options):
    """Drop-in for monte_carlo() (sonics.pyx:47-269): one genotype, returns the tab-joined row.

    Reads constants['alleles'] / ['noise_coef'] exactly as the reference does (set per genotype by
    process_one_genotype, sonics:418-419)."""
    if options.get("save_report"):
        return _monte_carlo_with_report(max_n_reps, constants, ranges, options)
    rows, _ = monte_carlo_batch([constants["alleles"]], [constants["noise_coef"]], max_n_reps, constants, ranges,
                                options)
    return rows[0]


def _simulate_records(e, params, table, n, seed_value, sim_begin=0):
    out = e.simulate(params, table, n, seed_value, sim_begin=sim_begin, readout=True, simpar=True, out_offset=0)
    e.sync()
    return out


def one_repeat(constants, ranges, how_many_reps=100):
    """Drop-in for one_repeat() (sonics.pyx:293-394): one simulation record
    [identity, r_squared, lnL, 'a/b', noise_coef, down, up, capture, efficiency]."""
    from . import engine as eng
    e = engine(0)
    params = eng.params_from_dicts(constants, ranges, {})
    table = e.build_table_from_arrays(params, [constants["alleles"]], [constants["noise_coef"]])
    out = _simulate_records(e, params, table, 1, _next_seed())
    return _records(out, table, constants["noise_coef"], 1)[0]


def _records(out, table, noise_coef, n):
    ident = out["identity"][0, :n].cpu().numpy()
    rsq = out["rsq"][0, :n].cpu().numpy()
    lnL = out["lnL"][0, :n].cpu().numpy()
    grp = out["group"][0, :n].cpu().numpy().view(np.uint16)
    sp = out["simpar"][0, :n].cpu().numpy()
    nc = float(np.float32(noise_coef))  # cdef float noise_coef (sonics.pyx:300)
    return [[float(ident[i]), float(rsq[i]), float(lnL[i]), table.group_label(0, int(grp[i])), nc, float(sp[i, 0]),
             float(sp[i, 1]), float(sp[i, 2]), float(sp[i, 3])] for i in range(n)]


REPORT_COLUMNS = ["ident", "r_squared", "lnL", "genotype", "noise_coef", "down", "up", "capture", "efficiency"]


def _monte_carlo_with_report(max_n_reps, constants, ranges, options):
    """--save_report path (sonics.pyx:110-113, 201-206): every simulation's record is written to
    <out_path>/<block>_<name>.txt, so the rounds are driven from here with the readout switched on."""
    from . import engine as eng
    e = engine(0)
    t = e.torch
    params = eng.params_from_dicts(constants, ranges, options)
    table = e.build_table_from_arrays(params, [constants["alleles"]], [constants["noise_coef"]])
    s = _next_seed()
    mc = bool(options.get("monte_carlo"))
    out = {"lnL": t.empty((1, max_n_reps), dtype=t.float64, device=e.device),
           "group": t.empty((1, max_n_reps), dtype=t.int16, device=e.device),
           "identity": t.empty((1, max_n_reps), dtype=t.float64, device=e.device),
           "rsq": t.empty((1, max_n_reps), dtype=t.float32, device=e.device),
           "simpar": t.empty((1, max_n_reps, 4), dtype=t.float64, device=e.device)}
    run, rnd = 0, 1
    reps = max_n_reps if mc else (100 if max_n_reps > 100 else max_n_reps)
    verdict = None
    while run < max_n_reps:
        e.simulate(params, table, reps, s, sim_begin=run, out=out)
        run += reps
        verdict = e.evaluate(params, table, out, run, final_round=run >= max_n_reps or mc)
        v = e.verdicts_to_host(verdict)[0]
        if mc or (v["filter"] & (1 | 16)):  # PASS, or passed with min_sim < 25 -> no_success (sonics.pyx:208)
            break
        reps = 4 * run if rnd % 2 == 1 else run
        if reps + run > max_n_reps:
            reps = max_n_reps - run
        rnd += 1
    recs = _records(out, table, constants["noise_coef"], run)
    path = os.path.join(options["out_path"], "{}_{}.txt".format(options["block"], options["name"]))
    with open(path, "w") as f:
        f.write("\t".join(REPORT_COLUMNS) + "\n")
        for r in recs:
            f.write("\t".join(_fmt(x) for x in r) + "\n")
    v = e.verdicts_to_host(verdict)[0].copy()
    if not mc and v["best_sim"] >= 0:
        v["identity"] = recs[int(v["best_sim"])][0]
        v["r_squared"] = recs[int(v["best_sim"])][1]
    return format_row(v, options)
