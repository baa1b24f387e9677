"""Shared helpers for the test-suite (oracle side)."""
import json
import math
import os

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEFAULT_RANGES = ((0, 0.1), (0, 0.1), (-0.25, 0.25), (0.001, 0.1))
CFG1 = "8|5;9|113;10|89"
CFG2 = "5|1;6|1;7|5;8|11;9|20;10|24;11|2;12|1"
CFG3 = ";".join(
    "%d|%d" % (a, int(5 + 400 * math.exp(-0.5 * ((a - 28) / 6) ** 2) + 300 * math.exp(-0.5 * ((a - 36) / 5) ** 2)))
    for a in range(10, 50))


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def cfg_from_case(noise_coef, kw, **extra):
    rg = kw.get("ranges", DEFAULT_RANGES)
    return orc.make_cfg(noise_coef=noise_coef, random_start=kw.get("random", False),
                        up_preference=kw.get("up_preference", False), floor_opt=kw.get("floor", 5),
                        n_cycles=kw.get("n_cycles", 24), capture_cycle=kw.get("capture_cycle", 13),
                        start_copies=kw.get("start_copies", 300000), down=rg[0], up=rg[1], capture=rg[2],
                        efficiency=rg[3], **extra)


def random_vcf(rs, n_lines, n_samples):
    """Odd-but-valid lobSTR VCF text (numpy RandomState rs): partial / negative / duplicate offsets, zero support,
    extra FORMAT fields, decoy REF= keys, samples without ALLREADS.  Used by the differential ingest tests and by
    tests/golden/make_golden.py (reference parse_vcf on the same text)."""
    lines = ["##fileformat=VCFv4.1",
             "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join("s%d" % i for i in range(n_samples))]
    for li in range(n_lines):
        L = int(rs.randint(1, 7))
        ref = int(rs.randint(3, 60))
        motif = "".join("ACGT"[i] for i in rs.randint(0, 4, L))
        info = ["END=%d" % (100 + li), "MOTIF=%s" % motif, "REF=%d%s" % (ref, rs.choice(["", ".0", ".5"]))]
        if rs.rand() < 0.3:
            info.insert(0, "XREF=99")        # a key that merely ends in REF=
        if rs.rand() < 0.3:
            info.insert(0, "REF=7")          # an earlier REF=: the last one wins
        fmt = ["GT", "ALLREADS"]
        if rs.rand() < 0.4:
            fmt.insert(int(rs.randint(0, 3)), "DP")
        cols = []
        for si in range(n_samples):
            if rs.rand() < 0.05:
                cols.append(".")             # no ALLREADS for this sample
                continue
            ents = []
            for _ in range(int(rs.randint(0, 7))):
                whole = rs.rand() < 0.8
                off = int(rs.randint(-6, 7)) * L + (0 if whole or L == 1 else int(rs.randint(1, L)))
                ents.append("%d|%d" % (off, int(rs.choice([0, 1, 3, 20, 46, 200, 5000]))))
            if rs.rand() < 0.2 and ents:
                ents.append(ents[0].split("|")[0] + "|%d" % int(rs.randint(1, 99)))  # duplicate allele: last wins
            sub = {"GT": "0/1", "DP": "12", "ALLREADS": ";".join(ents) if ents else "0|0"}
            cols.append(":".join(sub[k] for k in fmt))
        lines.append("b%d\t%d\t.\tN\t.\t.\t.\t%s;\t%s\t%s" % (li, 100 + li, ";".join(info), ":".join(fmt), "\t".join(cols)))
    return "\n".join(lines) + "\n"
