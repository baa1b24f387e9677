"""CPU tests of the native host I/O (csrc/host_io.inl) against its Python mirrors: VCF ingest +
pre-filters -> CSR, Python-compatible float formatting, and the output-file writer."""
import os
import random
import struct
import sys
import tempfile

import numpy as np
import pytest

from sonics_b200 import driver, engine as eng, hostio, sonics as drop_in
from tests.util import load, random_vcf

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _python_path(vcf_path, strict, monte_carlo=False):
    constants = {"one_allele_threshold": 45, "noise_threshold": 0.05}
    parsed = driver.parse_vcf(vcf_path, strict)
    kinds = [driver.prefilter(t, constants, {"monte_carlo": monte_carlo}) for t in parsed]
    work = [k[1] for k in kinds if k[0] == "sim"]
    off, al, rd = eng.to_csr([w[0] for w in work])
    return parsed, kinds, off, al, rd, np.array([w[1] for w in work], dtype=np.float32)


@pytest.mark.parametrize("strict", [1, 2])
def test_native_vcf_ingest_matches_python_mirror(strict):
    import synth_vcf
    texts = [load("vcf_parse.json")["vcf"]]
    with tempfile.TemporaryDirectory() as d:
        p0 = os.path.join(d, "golden.vcf")
        open(p0, "w").write(texts[0])
        p1 = os.path.join(d, "synth.vcf")
        loci = synth_vcf.make_genotypes(40, 5, seed=9, cov=(20, 400), noisy=0.3)
        synth_vcf.write_vcf(p1, loci, 5)
        # sprinkle edge cases: single-allele samples above / below the threshold, empty support, duplicates
        p2 = os.path.join(d, "edge.vcf")
        open(p2, "w").write(
            "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ta\tb\tc\td\n"
            "L1\t1\t.\tN\t.\t.\t.\tXREF=3;END=9;MOTIF=AC;REF=12;\tGT:ALLREADS\t0/0:0|46\t0/0:0|45\t0/0:0|0;2|0\t0/0:-2|5;0|9;-2|7\n"
            "L2\t2\t.\tN\t.\t.\t.\tMOTIF=GATA;REF=7.25;X=1\tGT:ALLREADS:DP\t0/0:-4|10;0|200;4|8:9\t0/0:-40|3;0|5\t.\t0/0:1|4;0|50;4|60\n")
        for path in (p0, p1, p2):
            parsed, kinds, off, al, rd, nz = _python_path(path, strict)
            v = hostio.Vcf(path, strict)
            if strict == 1:  # with strict 2 the mirror drops excluded samples before they become rows
                assert v.n_rows == len(parsed)
                assert list(v.row_kind) == [{"skip": 0, "row": 1, "sim": 2}[k[0]] for k in kinds]
            assert (v.allele_off == off).all() and (v.allele == al).all() and (v.reads == rd).all()
            assert (v.noise_coef == nz).all()
            v.close()


def test_native_vcf_errors():
    with tempfile.NamedTemporaryFile("w", suffix=".vcf", delete=False) as f:
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts1\nb\t1\t.\tN\t.\t.\t.\tMOTIF=AC;\tGT:ALLREADS\t0/0:0|5\n")
    try:
        with pytest.raises(Exception, match="expecting REF"):
            hostio.Vcf(f.name)
    finally:
        os.unlink(f.name)
    with pytest.raises(Exception, match="cannot open"):
        hostio.Vcf("/nonexistent/file.vcf")


def _random_doubles(n, seed=1):
    rnd = random.Random(seed)
    out = [0.0, -0.0, 1.0, -1.0, 0.1, 1e-4, 9.999e-5, 1e-5, 1e15, 1e16, 9.999999999999998e15, 123456789012345680.0,
           0.9033816425120773, -7.919938269071281, 6.471387605701286e-10, 65.68490269267932, -2.000000054512845e+16,
           5e-324, 1.7976931348623157e308, 0.9586113691329956, 100.0, 2.5e-3]
    while len(out) < n:
        bits = rnd.getrandbits(64)
        x = struct.unpack("<d", struct.pack("<Q", bits))[0]
        if x == x and abs(x) != float("inf"):
            out.append(x)
        out.append(rnd.uniform(-500, 10) * 10 ** rnd.randint(-12, 3))
    return out


def test_rows_writer_matches_python_format(tmp_path):
    vals = _random_doubles(4000)
    n = len(vals) // 8
    v = np.zeros(n, dtype=eng.VERDICT_DTYPE)
    rs = np.random.RandomState(0)
    v["allele_lo"] = rs.randint(1, 40, n)
    v["allele_hi"] = v["allele_lo"] + rs.randint(0, 4, n)
    v["filter"] = rs.choice([1, 2, 4, 8, 6, 10, 12, 14, 16, 0], n)
    v["run_reps"] = rs.choice([100, 500, 1000], n)
    v["p_clamped"] = rs.rand(n) < 0.1
    for k, fld in enumerate(["identity", "r_squared", "lnL", "high_pval", "best_lnL_ratio", "percentile_lnL_ratio",
                             "identity_q75", "r_squared_q75"]):
        v[fld] = vals[k * n:(k + 1) * n]
    v["lnL_q75"] = v["lnL"] - 1.5
    v["add_ratio"][:, 0] = v["best_lnL_ratio"] * 0.5
    v["add_ratio"][:, 1] = v["percentile_lnL_ratio"] * 3
    for mc, add in ((False, ""), (False, "0.5;0.9"), (True, "")):
        path = tmp_path / "out.txt"
        for i in range(0, n, 97):  # string mode writes one row: check a sample of verdicts
            hostio.write_rows(path, v[i:i + 1], None, mc, add, "S", "B")
            lines = open(path).read().split("\n")
            assert lines[0] == "\t".join(driver.HEADER_MC if mc else driver.HEADER).rstrip("\n")
            assert lines[1] == "S\tB\t.\t" + drop_in.format_row(v[i], {"monte_carlo": mc, "add_ratios": add})
    # skipped genotype in string mode: header only
    hostio.write_rows(tmp_path / "e.txt", v[:0], None, False, "")
    assert open(tmp_path / "e.txt").read().count("\n") == 1


def test_rows_writer_vcf_mode_order_and_no_simulations(tmp_path):
    p = tmp_path / "in.vcf"
    p.write_text("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ta\tb\tc\n"
                 "L1\t1\t.\tN\t.\t.\t.\tMOTIF=AC;REF=12;\tGT:ALLREADS\t0/0:0|46\t0/0:0|45\t0/0:-2|5;0|9\n"
                 "L2\t2\t.\tN\t.\t.\t.\tMOTIF=AC;REF=20;\tGT:ALLREADS\t0/0:-2|50;0|60\t0/0:0|0\t0/0:2|7;0|90\n")
    vcf = hostio.Vcf(p)
    assert list(vcf.row_kind) == [1, 0, 2, 2, 0, 2] and vcf.n_sim == 3
    v = np.zeros(3, dtype=eng.VERDICT_DTYPE)
    v["allele_lo"], v["allele_hi"], v["filter"], v["run_reps"] = [11, 19, 20], [12, 20, 21], [1, 16, 2], [100, 1000, 1000]
    v["lnL"] = [-3.5, 0, -8.25]
    out = tmp_path / "o.txt"
    hostio.write_rows(out, v, vcf)
    lines = open(out).read().strip().split("\n")[1:]
    assert lines[0] == "a\tL1\t12\t12/12\t.\t.\t.\tno_simulations\t.\t.\t.\t0\t."
    assert lines[1].startswith("c\tL1\t12\t11/12\t") and lines[2].startswith("a\tL2\t20\t./.\t.\t.\t.\tno_success")
    assert lines[3].startswith("c\tL2\t20\t20/21\t") and len(lines) == 4
    with pytest.raises(Exception, match="verdicts for more genotypes"):
        hostio.write_rows(out, v[:2], vcf)


def _csr(v):
    return (v.n_rows, v.n_sim, list(v.allele_off), list(v.allele), list(v.reads), list(v.noise_coef), list(v.row_kind))


def _join(parts):
    """Concatenate the CSR pieces of consecutive ranges the way the ranks' global numbering sees them."""
    n_rows = sum(p[0] for p in parts)
    n_sim = sum(p[1] for p in parts)
    off, al, rd, nz, kind = [0], [], [], [], []
    for p in parts:
        base = off[-1]
        off += [base + o for o in p[2][1:]]
        al += p[3]
        rd += p[4]
        nz += p[5]
        kind += p[6]
    return (n_rows, n_sim, off, al, rd, nz, kind)


@pytest.mark.parametrize("strict", [0, 1, 2])
def test_byte_range_tilings_parse_every_line_once(tmp_path, strict):
    """Multi-GPU ingest: any tiling of the file into byte ranges yields the whole-file parse, in order, including
    cuts inside the header, at line starts, inside lines, CRLF endings and a last line without newline; the
    random choice of --strict 0 does not depend on the tiling."""
    import synth_vcf
    p = str(tmp_path / "t.vcf")
    loci = synth_vcf.make_genotypes(7, 3, seed=5, cov=(20, 300), noisy=0.3)
    synth_vcf.write_vcf(p, loci, 3)
    text = open(p).read()
    # partial-repeat offsets (exercise --strict), a comment line inside the data, CRLF, no final newline
    lines = text.rstrip("\n").split("\n")
    lines[3] = lines[3].replace("0/0:", "0/0:1|7;", 1)
    lines.insert(5, "# a comment between data lines")
    lines[6] += "\r"
    open(p, "w", newline="").write("\n".join(lines))
    size = os.path.getsize(p)
    whole = hostio.Vcf(p, strict, seed=77)
    ref = _csr(whole)
    assert ref[0] > 0 and ref[1] > 0
    cuts = sorted(set(list(range(0, size + 1, 13)) + [text.index("locus1\t"), text.index("locus2\t"),
                                                      text.index("locus2\t") - 1, text.index("locus2\t") + 1, size]))
    for c in cuts:  # two ranges [0, c) [c, end)
        a = hostio.Vcf(p, strict, seed=77, byte_range=(0, c))
        b = hostio.Vcf(p, strict, seed=77, byte_range=(c, -1))
        assert _join([_csr(a), _csr(b)]) == ref, c
        # the newline scan counts exactly the data lines the parser assigns to [0, c) ...
        assert hostio.lines_before(p, c) == a.n_blocks, c
        # ... and so does the per-range count the ranks of a multi-GPU run exchange (memchr scan of the range only)
        assert hostio.count_lines(p, 0, c) == a.n_blocks and hostio.count_lines(p, c, -1) == b.n_blocks, c
        # ... so the RNG keys (cells of the line x sample grid) of a range do not depend on the cut
        assert list(a.uids(0)) + list(b.uids(a.n_blocks)) == list(whole.uids(0)), c
    for world in (3, 4, 8, 64):
        parts = [_csr(hostio.Vcf(p, strict, seed=77, byte_range=r)) for r in hostio.byte_ranges(p, world)]
        assert _join(parts) == ref, world
    for piece in (17, 64, 301, size, size + 5):  # the round-robin pieces of _run_vcf_ranks
        pcs = hostio.pieces_of(p, piece)
        vs = [hostio.Vcf(p, strict, seed=77, byte_range=r) for r in pcs]
        assert _join([_csr(v) for v in vs]) == ref, piece
        assert [hostio.count_lines(p, *r) for r in pcs] == [v.n_blocks for v in vs], piece


def test_part_files_concatenate_to_the_whole_output(tmp_path):
    import synth_vcf
    p = str(tmp_path / "t.vcf")
    synth_vcf.write_vcf(p, synth_vcf.make_genotypes(9, 4, seed=6, cov=(20, 300), noisy=0.3), 4)
    whole = hostio.Vcf(p)
    rs = np.random.RandomState(3)
    verdicts = np.zeros(whole.n_sim, dtype=hostio.VERDICT_DTYPE)
    verdicts["allele_lo"] = rs.randint(5, 20, whole.n_sim)
    verdicts["allele_hi"] = verdicts["allele_lo"] + rs.randint(0, 3, whole.n_sim)
    verdicts["filter"] = 1
    verdicts["lnL"] = -rs.rand(whole.n_sim) * 50
    verdicts["identity"] = rs.rand(whole.n_sim)
    verdicts["run_reps"] = 100
    hostio.write_rows(str(tmp_path / "whole.txt"), verdicts, whole)
    parts, lo = [], 0
    for r, br in enumerate(hostio.byte_ranges(p, 3)):
        v = hostio.Vcf(p, byte_range=br)
        part = str(tmp_path / ("part%d" % r))
        hostio.write_rows(part, verdicts[lo:lo + v.n_sim], v, header=(r == 0))
        lo += v.n_sim
        parts.append(part)
    assert lo == whole.n_sim
    hostio.concat_parts(str(tmp_path / "joined.txt"), parts)
    assert open(tmp_path / "joined.txt").read() == open(tmp_path / "whole.txt").read()
    assert not any(os.path.exists(q) for q in parts)


@pytest.mark.parametrize("seed", range(8))
def test_native_vcf_ingest_random_differential(tmp_path, seed):
    """Randomised differential test against the Python mirror of parse_vcf() + the pre-filters (--strict 1)."""
    rs = np.random.RandomState(1000 + seed)
    p = str(tmp_path / "r.vcf")
    open(p, "w").write(random_vcf(rs, 40, int(rs.randint(1, 6))))
    parsed, kinds, off, al, rd, nz = _python_path(p, 1)
    v = hostio.Vcf(p, 1)
    assert v.n_rows == len(parsed)
    assert list(v.row_kind) == [{"skip": 0, "row": 1, "sim": 2}[k[0]] for k in kinds]
    assert (v.allele_off == off).all() and (v.allele == al).all() and (v.reads == rd).all()
    assert (v.noise_coef == nz).all()
    assert {0, 2} <= set(v.row_kind.tolist())


def test_native_prefilters_match_reference_golden(tmp_path):
    """The native ingest's row kinds and noise coefficients on the random files of the golden fixture equal what the
    reference's process_one_genotype() does with the same genotypes (tests/golden/prefilter.json)."""
    import logging
    cases = load("prefilter.json")["cases"]
    logging.disable(logging.WARNING)
    try:
        for mc in (False, True):
            want = {tuple(c["tuple"]): c["written"] for c in cases if c["mc"] == mc}
            for seed in (1000, 1001):
                rs = np.random.RandomState(seed)
                p = str(tmp_path / ("r%d.vcf" % seed))
                open(p, "w").write(random_vcf(rs, 40, int(rs.randint(1, 6))))
                parsed = driver.parse_vcf(p, 1)
                v = hostio.Vcf(p, 1, monte_carlo=mc)
                assert v.n_rows == len(parsed)
                j = 0
                for t, kind in zip(parsed, v.row_kind):
                    w = want[tuple(t)]
                    assert kind == (0 if w == "" else 2 if "STUB" in w else 1), (t, w)
                    if kind == 2:
                        ref_nc = float(w.split("\t")[4])
                        assert v.noise_coef[j] == np.float32(ref_nc), (t, w)
                        j += 1
                assert j == v.n_sim
    finally:
        logging.disable(logging.NOTSET)


def test_native_output_file_structure_matches_reference_run_sonics(tmp_path):
    """Golden: whole output files of the reference's run_sonics() with monte_carlo() stubbed.  The native path
    (ingest -> verdicts -> row writer) must produce the same header, the same rows in the same order with the same
    sample / block / ref columns and the same no_simulations rows; simulated rows are compared up to their payload."""
    import logging
    logging.disable(logging.WARNING)
    try:
        for k, f in enumerate(load("prefilter.json")["files"]):
            p = str(tmp_path / ("in%d.vcf" % k))
            open(p, "w").write(f["vcf"])
            v = hostio.Vcf(p, 1, monte_carlo=f["mc"])
            verdicts = np.zeros(v.n_sim, dtype=hostio.VERDICT_DTYPE)
            verdicts["allele_lo"], verdicts["allele_hi"], verdicts["filter"], verdicts["run_reps"] = 9, 10, 1, 100
            out = str(tmp_path / ("out%d.txt" % k))
            hostio.write_rows(out, verdicts, v, monte_carlo=f["mc"])

            def mask(text):
                lines = text.split("\n")
                return [lines[0]] + [ln if "no_simulations" in ln else "\t".join(ln.split("\t")[:3]) for ln in lines[1:]]
            assert mask(open(out).read()) == mask(f["out"]), (k, f["mc"])
            assert sum("no_simulations" in ln for ln in f["out"].split("\n")) == int((v.row_kind == 1).sum())
    finally:
        logging.disable(logging.NOTSET)
