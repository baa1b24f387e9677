#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the REFERENCE ITSELF.

Runs in the build container only (needs oracle/_ref, built by oracle/build_ref.py
from /root/reference/sonics.pyx, plus the numpy/scipy/pandas of this image).
The outputs are committed; nothing at test time reads /root/reference.

  numpy_legacy_rng.json   numpy.random (legacy RandomState) primitive outputs
                          -> pins the MT19937 / binomial / poisson restatement
  one_repeat.json         sonics.one_repeat() records under np.random.seed(s)
                          -> pins oracle_one_repeat()/oracle_simulate() bit for bit
  monte_carlo_rows.json   sonics.monte_carlo() rows under np.random.seed(s)
                          -> pins oracle/stats.py (controller, MWU, quantiles, row format)
  rsq.json                sonics.rsq() / identity on random (reads, readout) pairs
                          -> pins oracle_descriptors() (and through it K1's readout descriptors)
  lnl_kat.json            lnL known-answer vectors: SURVEY.md 8(c) KAT-1..3 and
                          gammln spot values, plus mpmath 50-digit truths

Usage: python tests/golden/make_golden.py
"""
import json
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import sonics as ref  # noqa: E402  (the compiled reference)

DEFAULT_RANGES = ((0, 0.1), (0, 0.1), (-0.25, 0.25), (0.001, 0.1))
CFG2 = "5|1;6|1;7|5;8|11;9|20;10|24;11|2;12|1"
CFG3 = ";".join(
    "%d|%d" % (a, int(5 + 400 * math.exp(-0.5 * ((a - 28) / 6) ** 2) + 300 * math.exp(-0.5 * ((a - 36) / 5) ** 2)))
    for a in range(10, 50))


def auto_noise(gt, thr=0.05):
    """sonics:407-416 (CLI pre-filter), evaluated with the reference's get_alleles."""
    alleles = ref.get_alleles(gt)[0]
    frac = np.amin(alleles[alleles.nonzero()]) / sum(alleles) * np.count_nonzero(alleles)
    noise_frac = frac / (frac + 1)
    return float(np.amin(alleles[alleles.nonzero()]) / sum(alleles)) if noise_frac < thr else 0.0


def constants_for(gt, kw):
    alleles, max_allele, _ = ref.get_alleles(gt)
    noise = kw.get("noise_coef", 0)
    if noise == "auto":
        noise = auto_noise(gt)
    return dict(random=kw.get("random", False), n_cycles=kw.get("n_cycles", 24),
                capture_cycle=kw.get("capture_cycle", 13), up_preference=kw.get("up_preference", False),
                floor=kw.get("floor", 5), lnL_threshold=kw.get("lnL_threshold", 2.3),
                start_copies=kw.get("start_copies", 300000), genotype=gt, padjust=kw.get("padjust", 0.001),
                noise_threshold=0.05, one_allele_threshold=45, max_allele=max_allele,
                genotype_total=sum(alleles), alleles=alleles, noise_coef=noise), noise


def gen_rng():
    out = {"binomial": [], "poisson": [], "uniform": [], "randint": [], "stream": []}
    for seed in (0, 1, 12345, 4294967295):
        np.random.seed(seed)
        st = np.random.get_state()
        out["stream"].append({"seed": seed, "uint32": [int(x) for x in np.random.randint(0, 2**32, size=8, dtype=np.uint32)]})
        np.random.seed(seed)
        out["stream"][-1]["doubles"] = [np.random.random_sample().hex() for _ in range(4)]
    rs = np.random.RandomState(99)
    grid = []
    for n in (0, 1, 2, 7, 30, 100, 301, 1000, 4999, 20000, 150000, 300000, 1234567, 5000000):
        for p in (0.0, 1e-16, 0.001, 0.0123, 0.05, 0.1, 0.3, 0.5, 0.5000001, 0.7, 0.97, 0.999, 1.0):
            grid.append((n, p))
    for i in range(300):
        n = int(10 ** rs.uniform(0, 6.7))
        p = float(rs.choice([rs.uniform(0, 1), 10 ** rs.uniform(-5, 0)]))
        grid.append((n, p))
    for k, (n, p) in enumerate(grid):
        seed = 1000 + k
        np.random.seed(seed)
        draws = [int(np.random.binomial(n, p)) for _ in range(6)]
        tail = np.random.random_sample().hex()  # pins how much of the stream was consumed
        out["binomial"].append({"seed": seed, "n": n, "p": float(p).hex(), "draws": draws, "next": tail})
    lams = [0.0, 0.001, 0.5, 1.0, 5.0, 9.999, 10.0, 10.5, 33.3, 100.0, 1234.5, 99999.9, 412345.6, 1.5e6, 4.2e6]
    lams += [float(10 ** rs.uniform(-2, 6.5)) for _ in range(100)]
    for k, lam in enumerate(lams):
        seed = 5000 + k
        np.random.seed(seed)
        draws = [int(np.random.poisson(lam)) for _ in range(6)]
        tail = np.random.random_sample().hex()
        out["poisson"].append({"seed": seed, "lam": float(lam).hex(), "draws": draws, "next": tail})
    for k, (lo, hi) in enumerate([(1e-16, 0.1), (-0.25 + 1e-16, 0.25), (0.001 + 1e-16, 0.1), (1e-16, 0.0371)]):
        np.random.seed(7000 + k)
        out["uniform"].append({"seed": 7000 + k, "lo": float(lo).hex(), "hi": float(hi).hex(),
                               "draws": [float(np.random.uniform(lo, hi)).hex() for _ in range(5)]})
    for k, (lo, hi) in enumerate([(1, 4294967295), (0, 1), (0, 2), (0, 3), (0, 8), (0, 40), (0, 65), (0, 9937), (0, 2**31)]):
        np.random.seed(8000 + k)
        out["randint"].append({"seed": 8000 + k, "lo": lo, "hi": hi,
                               "draws": [int(np.random.randint(lo, hi)) for _ in range(8)],
                               "choice": [int(np.random.choice(np.arange(lo, hi))) for _ in range(4)]})
    return out


ONE_REPEAT_CASES = [
    ("8|5;9|113;10|89", {}),
    (CFG2, {}),
    ("9|1;10|54;11|914;12|15", {"noise_coef": "auto"}),
    ("7|2;8|22;9|150;10|1215;11|11", {"noise_coef": "auto"}),
    ("8|5;9|113;10|89", {"random": True}),
    ("8|5;9|113;10|89", {"up_preference": True}),
    ("8|5;9|113;10|89", {"floor": -1}),
    ("8|5;9|113;10|89", {"floor": 0}),
    ("8|5;9|113;10|89", {"floor": 1, "noise_coef": 0.01}),
    ("8|5;9|113;10|89", {"start_copies": 30001}),
    ("3|50;4|40", {}),
    ("9|100;10|100", {}),
    ("5|30;9|113;10|89;20|30", {}),
    ("14|700;15|2100;16|1900;17|2500;18|900;19|300", {}),
    (CFG3, {"n_cycles": 30, "capture_cycle": 16, "noise_coef": "auto"}),
    (CFG2, {"n_cycles": 10, "capture_cycle": 4,
            "ranges": ((0.01, 0.3), (0.0, 0.2), (-0.5, 0.5), (0.2, 0.9))}),
    ("1|40;2|60;3|7", {"floor": 5}),
]


def gen_one_repeat():
    out = []
    for k, (gt, kw) in enumerate(ONE_REPEAT_CASES):
        constants, noise = constants_for(gt, kw)
        ranges = kw.get("ranges", DEFAULT_RANGES)
        seed = 20261019 + k
        n = 12 if len(gt) > 100 else 25
        np.random.seed(seed)
        recs = [ref.one_repeat(constants, ranges, 100) for _ in range(n)]
        out.append({
            "genotype": gt, "seed": seed, "n": n, "noise_coef": float(noise),
            "options": {k2: v for k2, v in kw.items() if k2 not in ("noise_coef",)},
            "records": [{"identity": float(r[0]).hex(), "r_squared": float(r[1]).hex(), "lnL": float(r[2]).hex(),
                         "label": r[3], "noise_coef": float(r[4]).hex(), "down": float(r[5]).hex(),
                         "up": float(r[6]).hex(), "capture": float(r[7]).hex(), "efficiency": float(r[8]).hex()}
                        for r in recs],
        })
    return out


MC_CASES = [
    # (genotype, constants overrides, options overrides, max reps)
    ("8|5;9|113;10|89", {}, {}, 1000),
    ("9|1;10|54;11|914;12|15", {"noise_coef": "auto"}, {}, 1000),
    ("7|2;8|22;9|150;10|1215;11|11", {"noise_coef": "auto"}, {}, 1000),
    ("5|30;9|113;10|89;20|30", {}, {}, 1000),
    ("14|700;15|2100;16|1900;17|2500;18|900;19|300", {}, {}, 1000),
    ("9|60;10|40;11|30;12|8", {}, {}, 1000),
    ("9|100;10|100", {}, {}, 1000),
    (CFG2, {}, {}, 1000),
    ("8|5;9|113;10|89", {}, {"add_ratios": "0.5;0.6;0.8;0.9"}, 1000),
    ("9|60;10|40;11|30;12|8", {}, {"add_ratios": "0.5"}, 300),
    ("8|5;9|113;10|89", {}, {"monte_carlo": True}, 300),
    (CFG2, {}, {"monte_carlo": True}, 400),
    ("8|5;9|113;10|89", {"random": True}, {}, 1000),
    ("8|5;9|113;10|89", {}, {}, 60),
    ("12|30;13|200;14|180;15|25;16|3", {"noise_coef": "auto", "padjust": 0.05, "lnL_threshold": 1.0}, {}, 1000),
]


def gen_monte_carlo():
    out = []
    for k, (gt, ckw, okw, max_reps) in enumerate(MC_CASES):
        constants, noise = constants_for(gt, ckw)
        options = dict(repetitions=max_reps, out_path=".", strict=1, file_name="x", vcf_mode=False, name="sample",
                       block="Block", processes=0, verbose=False, save_report=False,
                       monte_carlo=okw.get("monte_carlo", False), add_ratios=okw.get("add_ratios", ""),
                       min_sim=okw.get("min_sim", 25))
        seed = 777000 + k
        np.random.seed(seed)
        row = ref.monte_carlo(max_reps, constants, DEFAULT_RANGES, options)
        out.append({"genotype": gt, "seed": seed, "max_reps": max_reps, "noise_coef": float(noise),
                    "constants": {k2: v for k2, v in ckw.items() if k2 != "noise_coef"},
                    "options": okw, "row": row})
        print(k, row)
    return out


def gen_lnl():
    import mpmath as mp
    mp.mp.dps = 50
    kats = [
        {"name": "KAT-1", "reads": {8: 5, 9: 113, 10: 89},
         "pool": {6: 300, 7: 4000, 8: 60000, 9: 900000, 10: 700000, 11: 30000, 12: 900},
         "survey_lnL": -9.431851759552956},
        {"name": "KAT-2", "reads": {5: 1, 6: 1, 7: 5, 8: 11, 9: 20, 10: 24, 11: 2, 12: 1},
         "pool": {3: 1, 4: 4, 5: 76, 6: 1436, 7: 17556, 8: 127899, 9: 441503, 10: 346980, 11: 33938, 12: 1962,
                  13: 85, 14: 1},
         "survey_lnL": -22.496943600177474},
        {"name": "KAT-3", "reads": {9: 1, 10: 54, 11: 914, 12: 15},
         "pool": {7: 12, 8: 950, 9: 41000, 10: 310000, 11: 1250000, 12: 52000, 13: 1100, 14: 9},
         "survey_lnL": -114.98292688390939},
    ]
    rs = np.random.RandomState(5)
    for j in range(40):
        lo = int(rs.randint(3, 30))
        nb = int(rs.randint(3, 12))
        pool = {lo + i: int(10 ** rs.uniform(0, 6.3)) for i in range(nb)}
        reads = {a: int(min(c, rs.randint(0, 400))) for a, c in pool.items() if rs.rand() < 0.7}
        reads = {a: c for a, c in reads.items() if c > 0}
        if not reads:
            reads = {lo: 1}
        kats.append({"name": "RAND-%d" % j, "reads": reads, "pool": pool})
    for kat in kats:
        T = sum(kat["pool"].values())
        D = sum(kat["reads"].values())
        v = mp.mpf(0)
        for a, x in kat["reads"].items():
            v += mp.log(mp.binomial(kat["pool"][a], x))
        v -= mp.log(mp.binomial(T, D))
        kat["truth"] = mp.nstr(v, 30)
        kat["reads"] = {str(a): c for a, c in kat["reads"].items()}
        kat["pool"] = {str(a): c for a, c in kat["pool"].items()}
    spot = {"gammln(1)": 0.0, "gammln(2)": -4.440892098500626e-16, "factln(10)": 15.104412573076393,
            "factln(100)": 363.73937555566835, "factln(150000)": 1637765.4640961345,
            "factln(1500000)": 19831471.52853508}
    return {"kats": kats, "spot": spot}


VCF_TEXT = """##fileformat=VCFv4.1
#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tsample1\tsample2\tsample3
genotype1\t100\t.\tN\t.\t.\t.\tEND=140;MOTIF=TCTA;REF=10;\tGT:ALLREADS\t0/0:-4|1;0|54;4|914;8|15\t0/0:-12|2;-8|22;-4|150;0|1215;4|11\t0/0:0|77
locus2\t200\t.\tN\t.\t.\t.\tEND=260;MOTIF=AC;REF=14.5;\tGT:ALLREADS\t0/0:-2|30;0|200;2|180;3|7\t0/0:0|40;2|45\t0/0:-6|3;-4|9;0|88;1|2
locus3\t300\t.\tN\t.\t.\t.\tEND=360;MOTIF=GAT;REF=9;\tGT:DP:ALLREADS\t0/0:9:-3|12;0|130\t0/0:3:0|3\t0/0:9:-9|1;3|5;6|60
"""


def gen_vcf():
    import importlib.machinery
    import tempfile
    cli = importlib.machinery.SourceFileLoader("sonics_cli_ref", "/root/reference/sonics").load_module()
    with tempfile.NamedTemporaryFile("w", suffix=".vcf", delete=False) as f:
        f.write(VCF_TEXT)
        path = f.name
    clean = lambda rows: [list(map(lambda x: int(x) if isinstance(x, (int, np.integer)) else x, t)) for t in rows]
    out = {"vcf": VCF_TEXT, "strict1": clean(cli.parse_vcf(path, 1))}
    os.unlink(path)
    # odd-but-valid random files (tests/util.random_vcf, regenerated from the seed at test time)
    import hashlib
    import logging
    from tests.util import random_vcf
    logging.disable(logging.WARNING)
    out["random"] = []
    for seed in range(8):
        rs = np.random.RandomState(1000 + seed)
        text = random_vcf(rs, 40, int(rs.randint(1, 6)))
        with tempfile.NamedTemporaryFile("w", suffix=".vcf", delete=False) as f:
            f.write(text)
        out["random"].append({"seed": 1000 + seed, "sha256": hashlib.sha256(text.encode()).hexdigest(),
                              "strict1": clean(cli.parse_vcf(f.name, 1))})
        os.unlink(f.name)
    logging.disable(logging.NOTSET)
    return out


def gen_prefilter():
    """process_one_genotype() of the reference CLI (sonics:345-439) with monte_carlo() stubbed out: what it
    writes for 0- / 1-allele genotypes and which alleles / noise_coef it hands to the simulation."""
    import importlib.machinery
    import logging
    import tempfile
    from tests.util import random_vcf
    cli = importlib.machinery.SourceFileLoader("sonics_cli_ref2", "/root/reference/sonics").load_module()

    def stub(reps, constants, ranges, options):
        al = constants["alleles"]
        return "STUB\t%r\t%s" % (float(constants["noise_coef"]),
                                  ";".join("%d|%d" % (a, al[a]) for a in np.nonzero(al)[0]))
    cli.sonics.monte_carlo = stub
    logging.disable(logging.WARNING)
    tuples = [("s", "b", 10, g) for g in
              ["9|46", "9|45", "9|44", "9|0", "", "0|5;9|100", "8|5;9|113;10|89", "5|1;6|1;7|5;8|11;9|20;10|24;11|2;12|1",
               "8|1;9|1000", "8|2;9|1000;10|30", "3|1;4|1;5|1;6|1;7|1;8|1;9|1;10|1;11|1;12|1;13|1;14|1;15|1;16|1;17|1;18|1;19|1;20|1;21|1;22|1;23|1;24|900",
               "9|7;9|300;10|5", "-3|5;9|50;10|40"]]
    for seed in (1000, 1001):
        rs = np.random.RandomState(seed)
        text = random_vcf(rs, 40, int(rs.randint(1, 6)))
        with tempfile.NamedTemporaryFile("w", suffix=".vcf", delete=False) as f:
            f.write(text)
        tuples += [tuple(int(x) if isinstance(x, (int, np.integer)) else x for x in t) for t in cli.parse_vcf(f.name, 1)]
        os.unlink(f.name)
    cases = []
    for mc in (False, True):
        for t in tuples:
            with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
                out_path = f.name
            constants = {"one_allele_threshold": 45, "noise_threshold": 0.05}
            options = {"out_file_path": out_path, "monte_carlo": mc, "repetitions": 1000}
            cli.process_one_genotype(t, constants, None, options)
            cases.append({"tuple": list(t), "mc": mc, "written": open(out_path).read()})
            os.unlink(out_path)
    # whole files through run_sonics(): header, row order, name / block / ref columns, no_simulations rows
    files = []
    texts = [VCF_TEXT]
    rs = np.random.RandomState(1002)
    texts.append(random_vcf(rs, 40, int(rs.randint(2, 6))))
    for text in texts:
        for mc in (False, True):
            d = tempfile.mkdtemp()
            vcf = os.path.join(d, "in.vcf")
            open(vcf, "w").write(text)
            constants = {"one_allele_threshold": 45, "noise_threshold": 0.05}
            options = {"out_path": d, "file_name": "out.txt", "monte_carlo": mc, "vcf_mode": True, "strict": 1,
                       "processes": 0, "repetitions": 1000, "name": "sample", "block": "Block"}
            cli.run_sonics(vcf, constants, None, options)
            files.append({"vcf": text, "mc": mc, "out": open(os.path.join(d, "out.txt")).read()})
    logging.disable(logging.NOTSET)
    return {"cases": cases, "files": files}


def gen_rsq():
    """The reference's own rsq() (sonics.pyx:273-291) and the identity expression (:382-383) on random
    (reads, readout) pairs, incl. ss_tot == 0 (flat reads -> the 1e-16 guard), readouts outside the span of
    the reads, tiny totals.  r_squared is stored through a C float in one_repeat (:300): both forms are kept."""
    rs = np.random.RandomState(20261020)
    cases = []
    for k in range(400):
        reads = np.zeros(500, dtype=np.int64)
        readout = np.zeros(500, dtype=np.int64)
        a0 = int(rs.randint(1, 440))
        n_al = int(rs.randint(1, 12))
        idx = np.sort(rs.choice(np.arange(a0, a0 + 40), n_al, replace=False))
        kind = k % 5
        if kind == 0:      # flat reads on contiguous alleles: ss_tot == 0 when the readout stays inside
            idx = np.arange(a0, a0 + n_al)
            reads[idx] = int(rs.randint(1, 500))
        elif kind == 1:    # tiny totals
            reads[idx] = rs.randint(1, 4, n_al)
        else:
            reads[idx] = np.maximum(1, (10 ** rs.uniform(0, 3.7, n_al)).astype(np.int64))
        D = int(reads.sum())
        if kind == 0 and k % 2 == 0:
            ridx = idx
        else:
            span = np.arange(max(1, idx[0] - int(rs.randint(0, 4))), min(499, idx[-1] + int(rs.randint(0, 4))) + 1)
            ridx = span[rs.rand(len(span)) < 0.7]
            if len(ridx) == 0:
                ridx = span[:1]
        w = rs.dirichlet(np.ones(len(ridx)) * 0.5)
        readout[ridx] = rs.multinomial(D, w)
        if readout.sum() == 0:
            readout[ridx[0]] = D
        r = float(ref.rsq(reads, readout))
        nz = reads.nonzero()[0]
        ident = sum([min(reads[i], readout[i]) for i in nz]) / sum(reads)
        cases.append({"reads": {str(int(i)): int(reads[i]) for i in nz},
                      "readout": {str(int(i)): int(readout[i]) for i in readout.nonzero()[0]},
                      "rsq": r.hex(), "rsq_f32": float(np.float32(r)).hex(), "identity": float(ident).hex()})
    return cases


def dump(name, obj):
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(obj, f, indent=0, separators=(",", ":"))
        f.write("\n")
    print("wrote", name, os.path.getsize(os.path.join(HERE, name)), "bytes")


if __name__ == "__main__":
    which = sys.argv[1:] or ["rng", "one_repeat", "monte_carlo", "lnl", "vcf", "prefilter", "rsq"]
    if "rsq" in which:
        dump("rsq.json", gen_rsq())
    if "prefilter" in which:
        dump("prefilter.json", gen_prefilter())
    if "vcf" in which:
        dump("vcf_parse.json", gen_vcf())
    if "rng" in which:
        dump("numpy_legacy_rng.json", gen_rng())
    if "one_repeat" in which:
        dump("one_repeat.json", gen_one_repeat())
    if "lnl" in which:
        dump("lnl_kat.json", gen_lnl())
    if "monte_carlo" in which:
        dump("monte_carlo_rows.json", gen_monte_carlo())
