"""CPU tests of the host driver (CLI flags, VCF ingest, pre-filters, sharding, 2-rank gather)."""
import os
import subprocess
import sys
import tempfile
import textwrap

import numpy as np
import pytest

from oracle import oracle as orc
from sonics_b200 import driver, sonics as drop_in
from tests.util import load

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_get_alleles_contract():
    a, mx, n = drop_in.get_alleles("8|5;9|113;10|89")
    assert mx == 500 and n == 3 and a.dtype == np.int64 and a.shape == (500,)
    assert (a == orc.parse_reads("8|5;9|113;10|89")).all()
    a, _, n = drop_in.get_alleles("0|7;-3|9;;12|4;")  # alleles <= 0 ignored, empty fields skipped (sonics.pyx:24-27)
    assert n == 1 and a[12] == 4
    with pytest.raises(IndexError):
        drop_in.get_alleles("500|1")


def test_cli_defaults_match_reference():
    """Flags and defaults of the reference's argparse (sonics:445-693)."""
    a = driver.build_parser().parse_args(["8|5;9|113;10|89"])
    exp = dict(out_path=".", file_name="sonics_out.txt", processes=0, strict=1, pcr_cycles=12, after_capture=12,
               block="Block", name="sample", repetitions=1000, padjust=0.001, lnL_threshold=2.3, start_copies=300000,
               noise_threshold=0.05, floor=5, min_sim=25, one_allele=45, efficiency=(0.001, 0.1), down=(0, 0.1),
               up=(0, 0.1), capture=(-0.25, 0.25), add_ratios="", vcf_mode=False, monte_carlo=False, random=False,
               save_report=False, up_preference=False, verbose=False)
    for k, v in exp.items():
        assert getattr(a, k) == v, k
    short = {"-o": "out_path", "-n": "file_name", "-p": "processes", "-t": "strict", "-y": "pcr_cycles",
             "-c": "after_capture", "-b": "block", "-a": "name", "-r": "repetitions", "-g": "padjust",
             "-i": "lnL_threshold", "-s": "start_copies", "-l": "noise_threshold", "-f": "floor", "-j": "one_allele"}
    for flag, dest in short.items():
        val = "7"
        b = driver.build_parser().parse_args([flag, val, "x"])
        assert str(getattr(b, dest)) in ("7", "7.0")
    constants, ranges, options = driver.dicts_from_args(driver.build_parser().parse_args(["-y", "15", "-c", "15", "x"]))
    assert constants["n_cycles"] == 30 and constants["capture_cycle"] == 16  # sonics:711-712
    assert set(constants) == {"random", "n_cycles", "capture_cycle", "up_preference", "floor", "lnL_threshold",
                              "start_copies", "genotype", "padjust", "noise_threshold", "one_allele_threshold"}
    assert set(options) == {"repetitions", "out_path", "strict", "file_name", "vcf_mode", "name", "block", "processes",
                            "verbose", "save_report", "monte_carlo", "add_ratios", "min_sim"}
    assert len(ranges) == 4


def test_parse_vcf_matches_reference():
    g = load("vcf_parse.json")
    with tempfile.NamedTemporaryFile("w", suffix=".vcf", delete=False) as f:
        f.write(g["vcf"])
    try:
        got = driver.parse_vcf(f.name, 1)
        assert [list(t) for t in got] == g["strict1"]
        # strict 2 drops the sample with a partial repeat, later samples keep their own column
        got2 = driver.parse_vcf(f.name, 2)
        assert ("sample2", "locus2", 14, "14|40;15|45") in got2
        assert not any(t[0] in ("sample1", "sample3") and t[1] == "locus2" for t in got2)
        # strict 0 moves the partial support (15.5 -> 15 or 16; 14.5 -> 14 or 15)
        got0 = driver.parse_vcf(f.name, 0, rng=np.random.default_rng(1))
        s1 = dict(x.split("|") for x in [t for t in got0 if t[:2] == ("sample1", "locus2")][0][3].split(";"))
        assert sum(int(v) for v in s1.values()) == 30 + 200 + 180 + 7
    finally:
        os.unlink(f.name)
    with tempfile.NamedTemporaryFile("w", suffix=".vcf", delete=False) as f:
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\ts1\nb\t1\t.\tN\t.\t.\t.\tMOTIF=AC;\tGT:ALLREADS\t0/0:0|5\n")
    try:
        with pytest.raises(Exception, match="expecting REF"):
            driver.parse_vcf(f.name, 1)
    finally:
        os.unlink(f.name)


def test_prefilter_rules():
    constants = {"one_allele_threshold": 45, "noise_threshold": 0.05}
    opts = {"monte_carlo": False}
    assert driver.prefilter(("s", "b", 9, ""), constants, opts) == ("skip", None)
    assert driver.prefilter(("s", "b", 9, "10|45"), constants, opts) == ("skip", None)  # not above the threshold
    kind, row = driver.prefilter(("s", "b", 9, "10|46"), constants, opts)
    assert kind == "row" and row == "10/10\t.\t.\t.\tno_simulations\t.\t.\t.\t0\t."
    assert driver.prefilter(("s", "b", 9, "10|46"), constants, {"monte_carlo": True}) == ("skip", None)
    for gt in ("9|1;10|54;11|914;12|15", "8|5;9|113;10|89", "9|100;10|100"):
        kind, (alleles, nc) = driver.prefilter(("s", "b", 9, gt), constants, opts)
        assert kind == "sim" and nc == orc.auto_noise_coef(orc.parse_reads(gt))


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 1000, 80001):
        for world in (1, 2, 3, 8):
            parts = [driver.shard_bounds(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def test_parse_vcf_matches_reference_on_random_files(tmp_path):
    """The reference's own parse_vcf() on odd-but-valid random files (golden, generated in the build container)."""
    import hashlib
    import logging
    from tests.util import random_vcf
    logging.disable(logging.WARNING)
    try:
        for case in load("vcf_parse.json")["random"]:
            rs = np.random.RandomState(case["seed"])
            text = random_vcf(rs, 40, int(rs.randint(1, 6)))
            assert hashlib.sha256(text.encode()).hexdigest() == case["sha256"], "generator drifted from the fixture"
            p = tmp_path / ("r%d.vcf" % case["seed"])
            p.write_text(text)
            assert [list(t) for t in driver.parse_vcf(str(p), 1)] == case["strict1"], case["seed"]
    finally:
        logging.disable(logging.NOTSET)


def test_synthetic_vcf_roundtrip():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import synth_vcf
    loci = synth_vcf.make_genotypes(5, 3, seed=3, noisy=0.4)
    with tempfile.NamedTemporaryFile("w", suffix=".vcf", delete=False) as f:
        pass
    try:
        synth_vcf.write_vcf(f.name, loci, 3)
        assert driver.parse_vcf(f.name, 1) == synth_vcf.genotype_strings(loci)
    finally:
        os.unlink(f.name)


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np
    import torch.distributed as dist
    from sonics_b200 import driver, hostio
    vcf_path, out_path = sys.argv[1], sys.argv[2]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    def compute(off, al, rd, nz, uids):   # stands in for the GPU: a verdict that encodes (uid, first allele, reads)
        v = np.zeros(len(nz), dtype=hostio.VERDICT_DTYPE)
        if len(nz):
            v["allele_lo"] = al[off[:-1]]
            v["allele_hi"] = al[off[1:] - 1]
            v["filter"] = 1
            v["run_reps"] = uids
            v["lnL"] = -np.add.reduceat(rd, off[:-1]).astype(np.float64)
        return v
    constants = {"one_allele_threshold": 45, "noise_threshold": 0.05}
    options = {"strict": 1, "monte_carlo": False, "add_ratios": "", "name": "sample", "block": "Block"}
    driver._run_vcf_ranks(vcf_path, constants, None, options, 5, 65536, out_path, rank, world, compute=compute,
                          piece_bytes=700)   # several pipeline pieces per rank
    dist.barrier()
    if rank == 0:
        assert driver.LAST_TIMINGS["pieces"] > 2
        whole = hostio.Vcf(vcf_path)
        hostio.write_rows(out_path + ".single",
                          compute(whole.allele_off, whole.allele, whole.reads, whole.noise_coef, whole.uids(0)), whole)
        assert open(out_path).read() == open(out_path + ".single").read()
        assert not os.path.exists(out_path + ".part0") and not os.path.exists(out_path + ".part1")
        # a longer file of an earlier run under the same name: the join must cut it to the new length
        open(out_path + ".auto", "w").write("x" * (os.path.getsize(out_path) + 5000))
    dist.barrier()
    # the automatic piece plan (a short first piece per rank, then equal pieces, the same number for every rank)
    driver.PIECE_BYTES = 900
    driver._run_vcf_ranks(vcf_path, constants, None, options, 5, 65536, out_path + ".auto", rank, world,
                          compute=compute)
    dist.barrier()
    if rank == 0:
        assert driver.LAST_TIMINGS["pieces_total"] %% world == 0 and driver.LAST_TIMINGS["pieces_total"] >= 2 * world
        assert open(out_path + ".auto").read() == open(out_path + ".single").read()
        print("RANKS_OK", whole.n_sim)
    dist.destroy_process_group()
""")


def _stub_compute(off, al, rd, nz, uids):
    from sonics_b200 import hostio
    v = np.zeros(len(nz), dtype=hostio.VERDICT_DTYPE)
    if len(nz):
        v["allele_lo"] = al[off[:-1]]
        v["allele_hi"] = al[off[1:] - 1]
        v["filter"] = 1
        v["run_reps"] = uids
        v["lnL"] = -np.add.reduceat(rd, off[:-1]).astype(np.float64)
    return v


@pytest.mark.parametrize("piece_bytes", [64, 333, 5000, 1 << 20])
def test_pipelined_vcf_pieces_reproduce_the_whole_file_run(tmp_path, piece_bytes):
    """Single-process VCF mode is a pipeline over byte pieces of the file (parse k+1 / simulate k / write k-1):
    the output and the RNG keys (cell in the line x sample grid) do not depend on the piece size."""
    from sonics_b200 import hostio
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import synth_vcf
    vcf = str(tmp_path / "in.vcf")
    synth_vcf.write_vcf(vcf, synth_vcf.make_genotypes(31, 4, seed=8, cov=(20, 300), noisy=0.3), 4)
    constants = {"one_allele_threshold": 45, "noise_threshold": 0.05}
    options = {"strict": 1, "monte_carlo": False, "add_ratios": "", "name": "sample", "block": "Block"}
    whole = hostio.Vcf(vcf)
    uids = whole.uids(0)
    assert len(set(uids.tolist())) == whole.n_sim and (np.diff(uids.astype(np.int64)) > 0).all()
    assert hostio.lines_before(vcf, -1) == whole.n_blocks == 31
    hostio.write_rows(str(tmp_path / "whole.txt"),
                      _stub_compute(whole.allele_off, whole.allele, whole.reads, whole.noise_coef, uids), whole)
    out = str(tmp_path / "piped.txt")
    n_rows, n_sim, tm = driver.vcf_pipeline(vcf, (0, -1), 0, out, True, constants, options, 5, _stub_compute,
                                            piece_bytes=piece_bytes)
    assert (n_rows, n_sim) == (whole.n_rows, whole.n_sim)
    assert open(out).read() == open(tmp_path / "whole.txt").read()
    # a range in the middle of the file with its line_base from the newline scan
    size = os.path.getsize(vcf)
    lo = size // 3
    base = hostio.lines_before(vcf, lo)
    mid = hostio.Vcf(vcf, byte_range=(lo, -1))
    assert list(mid.uids(base)) == list(uids[whole.n_sim - mid.n_sim:])


def test_two_rank_gloo_vcf_ingest_simulate_write(tmp_path):
    """N > 1 host path on CPU (gloo, 2 ranks): byte-range ingest, uid = position in the whole file, per-rank part
    files joined in rank order == the single-process output (the GPU call is replaced by a stub)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import synth_vcf
    vcf = str(tmp_path / "in.vcf")
    synth_vcf.write_vcf(vcf, synth_vcf.make_genotypes(23, 3, seed=4, cov=(20, 300), noisy=0.3), 3)
    worker = str(tmp_path / "worker.py")
    open(worker, "w").write(WORKER % ROOT)
    import socket
    with socket.socket() as sk:  # a free rendezvous port (a fixed one may be taken on a shared box)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), worker, vcf,
                          str(tmp_path / "out.txt")], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "RANKS_OK" in out.stdout


def test_piece_plan_tiles_the_file_evenly(tmp_path):
    """even_piece_bytes / pieces_of(heads): the pieces tile the file without gap or overlap, every rank gets the same
    number of them, and (files large enough for it) each rank starts on a short piece."""
    from sonics_b200 import hostio
    path = str(tmp_path / "f.bin")
    for size in (1, 99, 3000, 70_001, 3_000_000, 50_000_000):
        with open(path, "wb") as f:
            f.truncate(size)
        for world in (1, 2, 3, 4, 8):
            head = driver.even_piece_bytes(size, world) // 4
            body = driver.even_piece_bytes(max(0, size - world * head), world)
            pcs = hostio.pieces_of(path, body, world, head)
            assert pcs[0][0] == 0 and pcs[-1][1] == -1
            assert all(pcs[i][1] == pcs[i + 1][0] for i in range(len(pcs) - 1))
            assert all(b > a for a, b in pcs[:-1]) and pcs[-1][0] < size
            counts = [len(range(r, len(pcs), world)) for r in range(world)]
            if size >= 64 * world:
                assert max(counts) - min(counts) <= 1, (size, world, counts)
            if size >= 3_000_000:
                assert len(set(counts)) == 1 and counts[0] >= 2, (size, world, counts)
                assert all(pcs[r][1] - pcs[r][0] == head for r in range(world))
                assert max(b - a for a, b in pcs[:-1]) <= driver.PIECE_BYTES


def test_shard_jobs_cover_the_batch_with_their_own_keys():
    """--gpus N in one process: contiguous CSR shards whose concatenation is the batch, RNG keys sliced alike."""
    off = np.array([0, 2, 5, 7, 9, 12, 14, 16, 19, 21, 23, 25], dtype=np.int32)   # 11 ragged genotypes
    al = np.arange(25, dtype=np.int32)
    rd = al + 1
    nz = np.linspace(0, 1, 11).astype(np.float32)
    uids = (np.arange(11) * 7 + 3).astype(np.uint32)
    for gpus in (2, 3, 4, 11, 16):
        jobs = driver.shard_jobs(off, al, rd, nz, uids, gpus, ("c", "r", "o", 5), 65536)
        assert len(jobs) == gpus and [j[8] for j in jobs] == list(range(gpus))
        got_al = np.concatenate([j[1] for j in jobs])
        got_uid = np.concatenate([j[10] for j in jobs])
        got_first = [int(j[1][j[0][k]]) for j in jobs for k in range(len(j[3]))]
        assert (got_al == al).all() and (got_uid == uids).all()
        assert got_first == [int(al[off[g]]) for g in range(11)]
        assert all(j[0][0] == 0 and j[0][-1] == len(j[1]) == len(j[2]) for j in jobs)
        assert [j[4:8] for j in jobs] == [("c", "r", "o", 5)] * gpus and all(j[9] == 65536 for j in jobs)


def _expected_written(t, kind, payload):
    if kind == "skip":
        return ""
    if kind == "row":
        return "{}\t{}\t{}\t{}\n".format(t[0], t[1], t[2], payload)
    alleles, noise_coef = payload
    return "{}\t{}\t{}\tSTUB\t{!r}\t{}\n".format(
        t[0], t[1], t[2], float(noise_coef), ";".join("%d|%d" % (a, alleles[a]) for a in np.nonzero(alleles)[0]))


def test_prefilter_matches_reference_process_one_genotype():
    """Golden: the reference's process_one_genotype() with monte_carlo() stubbed (tests/golden/make_golden.py) --
    what is written for 0-/1-allele genotypes (incl. the MC-mode row that is built but never written) and the
    alleles / auto-noise coefficient handed to the simulation, on hand-made and random genotypes."""
    import logging
    logging.disable(logging.WARNING)
    try:
        cases = load("prefilter.json")["cases"]
        assert len(cases) > 400
        for c in cases:
            t = tuple(c["tuple"])
            kind, payload = driver.prefilter(t, {"one_allele_threshold": 45, "noise_threshold": 0.05},
                                             {"monte_carlo": c["mc"]})
            assert _expected_written(t, kind, payload) == c["written"], c
    finally:
        logging.disable(logging.NOTSET)
