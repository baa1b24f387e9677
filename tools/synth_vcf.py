#!/usr/bin/env python
"""Synthetic lobSTR-style VCF generator for the genotyping benchmarks (BASELINE.json configs 4-5).

numpy only, deterministic in the seed.  Per locus: motif length L in {2,3,4}, REF repeat count R in
[8,30]; per sample a true genotype (R+d1, R+d2), d in {-3..3}, coverage C ~ U{cov_lo..cov_hi}; the
read support is a multinomial draw from a two-peak stutter profile (down-stutter 4-14 % per step,
up-stutter a quarter of that, geometric decay).  `noisy` adds, for a fraction of the samples, either a
far contaminating allele (-> no_success) or a flat 3-4 peak profile (-> MWU_test) as in config 5.
ALLREADS is encoded as in README.md:75-78 of the reference: ((allele - R) * L)|reads;...
"""
import argparse

import numpy as np

BASES = "ACGT"


def profile(a1, a2, down, up, span=6):
    w = {}
    for a in (a1, a2):
        for k in range(-span, span + 1):
            r = down ** (-k) if k < 0 else up ** k
            w[a + k] = w.get(a + k, 0.0) + 0.5 * r
    return w


def make_genotypes(n_loci, n_samples, seed=20261019, cov=(50, 1000), noisy=0.0):
    rs = np.random.RandomState(seed)
    loci = []
    for li in range(n_loci):
        L = int(rs.randint(2, 5))
        R = int(rs.randint(8, 31))
        motif = "".join(BASES[i] for i in rs.randint(0, 4, L))
        samples = []
        for si in range(n_samples):
            d1, d2 = rs.randint(-3, 4, 2)
            a1, a2 = R + int(d1), R + int(d2)
            C = int(rs.randint(cov[0], cov[1] + 1))
            down = rs.uniform(0.04, 0.14)
            w = profile(a1, a2, down, down * rs.uniform(0.15, 0.35))
            mode = rs.rand()
            if mode < noisy * 0.5:
                far = max(a1, a2) + int(rs.randint(6, 10))
                w[far] = w.get(far, 0.0) + rs.uniform(0.12, 0.3)
            elif mode < noisy:
                for k, a in enumerate(range(min(a1, a2) - 1, min(a1, a2) + 3)):
                    w[a] = w.get(a, 0.0) + rs.uniform(0.5, 0.9)
            keys = sorted(k for k in w if k > 0)
            p = np.array([w[k] for k in keys])
            counts = rs.multinomial(C, p / p.sum())
            samples.append([(k, int(c)) for k, c in zip(keys, counts) if c > 0])
        loci.append((L, R, motif, samples))
    return loci


def sample_columns(loci):
    """The sample columns of every locus as one string each (what write_vcf renders; reusable across copies)."""
    out = []
    for L, R, motif, samples in loci:
        out.append("\t".join("0/0:" + ";".join("%d|%d" % ((a - R) * L, c) for a, c in reads) for reads in samples))
    return out


def write_vcf(path, loci, n_samples, start=0, header=True, columns=None, mode="w"):
    """start: number of loci before these (names / positions continue); header=False: data lines only;
    columns: sample_columns(loci) rendered earlier; mode "a" appends."""
    columns = columns or sample_columns(loci)
    with open(path, mode) as f:
        if header:
            f.write("##fileformat=VCFv4.1\n")
            f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" +
                    "\t".join("sample%d" % (i + 1) for i in range(n_samples)) + "\n")
        f.write("".join("locus%d\t%d\t.\tN\t.\t.\t.\tEND=%d;MOTIF=%s;REF=%d;\tGT:ALLREADS\t%s\n" %
                        (li + 1, 1000 + li, 1000 + li + L * R, motif, R, cols)
                        for (li, (L, R, motif, samples)), cols in zip(enumerate(loci, start), columns)))


def genotype_strings(loci):
    """The same genotypes as 'allele|reads;...' strings, in VCF order (what parse_vcf returns)."""
    out = []
    for li, (L, R, motif, samples) in enumerate(loci):
        for si, reads in enumerate(samples):
            out.append(("sample%d" % (si + 1), "locus%d" % (li + 1), R, ";".join("%d|%d" % (a, c) for a, c in reads)))
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--loci", type=int, default=100)
    ap.add_argument("--samples", type=int, default=8)
    ap.add_argument("--seed", type=int, default=20261019)
    ap.add_argument("--cov", type=int, nargs=2, default=(50, 1000))
    ap.add_argument("--noisy", type=float, default=0.0)
    a = ap.parse_args()
    write_vcf(a.out, make_genotypes(a.loci, a.samples, a.seed, tuple(a.cov), a.noisy), a.samples)
