"""ctypes front-end of the native host I/O in libsonics_b200.so (csrc/host_io.inl):
VCF ingest + pre-filters straight to CSR, and the output-file writer."""
import ctypes
import logging

import numpy as np

from . import _abi
from ._abi import check
from .engine import VERDICT_DTYPE


class Vcf:
    """Parsed lobSTR VCF: CSR of the genotypes to simulate + per-row kinds (0 skip, 1 no_simulations, 2 simulate)."""

    def __init__(self, path, strict=1, one_allele_threshold=45, noise_threshold=0.05, monte_carlo=False, seed=0,
                 byte_range=None):
        """byte_range=(begin, end): only the data lines whose first byte lies in [begin, end) (end < 0: to the end
        of the file); ranges that tile the file parse every line exactly once (one range per rank)."""
        self.lib = _abi.load()
        h = ctypes.c_void_p()
        b0, b1 = (0, -1) if byte_range is None else (int(byte_range[0]), int(byte_range[1]))
        check(self.lib.sonics_vcf_parse_range(str(path).encode(), b0, b1, int(strict), int(one_allele_threshold),
                                              float(noise_threshold), int(bool(monte_carlo)),
                                              int(seed) & (2**64 - 1), ctypes.byref(h)))
        self.handle = h
        c = lambda w: int(self.lib.sonics_vcf_count(h, w))
        self.n_rows, self.n_sim, n_entries, self.n_samples, self.n_blocks, n_warn = (c(i) for i in range(6))
        self.allele_off = np.zeros(self.n_sim + 1, dtype=np.int32)
        self.allele = np.zeros(n_entries, dtype=np.int32)
        self.reads = np.zeros(n_entries, dtype=np.int32)
        self.noise_coef = np.zeros(self.n_sim, dtype=np.float32)
        self.row_kind = np.zeros(self.n_rows, dtype=np.int32)
        check(self.lib.sonics_vcf_fill(h, self.allele_off.ctypes.data, self.allele.ctypes.data, self.reads.ctypes.data,
                                       self.noise_coef.ctypes.data, self.row_kind.ctypes.data))
        self.warnings = [self.lib.sonics_vcf_warning(h, i).decode("utf-8", "replace") for i in range(min(n_warn, 1000))]
        for w in self.warnings[:20]:
            logging.warning(w)

    def uids(self, line_base=0):
        """RNG keys of the genotypes to simulate: (line_base + data-line ordinal) * samples + sample index, with
        line_base = the number of data lines before this range (lines_before)."""
        out = np.zeros(self.n_sim, dtype=np.uint32)
        check(self.lib.sonics_vcf_uids(self.handle, int(line_base), out.ctypes.data))
        return out

    def close(self):
        if self.handle:
            self.lib.sonics_vcf_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def lines_before(path, byte_end):
    """Data lines of the VCF that start before byte_end (the line_base of a range starting there)."""
    n = ctypes.c_int64(0)
    check(_abi.load().sonics_vcf_lines_before(str(path).encode(), int(byte_end), ctypes.byref(n)))
    return n.value


def count_lines(path, byte_begin, byte_end):
    """Data lines whose first byte lies in [byte_begin, byte_end) (byte_end < 0: end of file): the blocks a
    Vcf(byte_range=(byte_begin, byte_end)) parse yields."""
    n = ctypes.c_int64(0)
    check(_abi.load().sonics_vcf_count_lines(str(path).encode(), int(byte_begin), int(byte_end), ctypes.byref(n)))
    return n.value


def pieces_of(path, piece_bytes, heads=0, head_bytes=0):
    """[begin, end) byte pieces that tile the file (the last one ends at -1 = end of file).  heads > 0: the first
    `heads` pieces are head_bytes long (one short first piece per rank of a round-robin deal: the first parse of a
    pipeline is the one nothing overlaps)."""
    import os
    size = os.path.getsize(path)
    piece_bytes = max(1, int(piece_bytes))
    cuts = []
    if heads > 0 and 0 < head_bytes < piece_bytes and heads * head_bytes < size:
        cuts = [i * int(head_bytes) for i in range(heads)]
        start = heads * int(head_bytes)
    else:
        start = 0
    cuts += list(range(start, size, piece_bytes)) + [size]
    out = [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)]
    if out:
        out[-1] = (out[-1][0], -1)
    return out


def write_rows(path, verdicts, vcf=None, monte_carlo=False, add_ratios="", name="sample", block="Block", header=True,
               append=False):
    """Header + rows of run_sonics() (sonics:174-207, :437); verdicts: structured array (VERDICT_DTYPE).
    header=False writes the rows only, append=True adds them to an existing file (pieces of a pipelined or
    multi-rank run, see concat_parts)."""
    lib = _abi.load()
    v = np.ascontiguousarray(verdicts, dtype=VERDICT_DTYPE)
    check(lib.sonics_rows_write_part(str(path).encode(), vcf.handle if vcf is not None else None, v.ctypes.data,
                                     len(v), int(bool(monte_carlo)), add_ratios.encode(), name.encode(),
                                     block.encode(), int(bool(header)) | (2 if append else 0)))


def byte_ranges(path, world):
    """[begin, end) per rank: equal byte shares of the file (the parser assigns each line to the range that
    holds its first byte)."""
    import os
    size = os.path.getsize(path)
    return [(r * size // world, (r + 1) * size // world if r + 1 < world else -1) for r in range(world)]


def concat_parts(path, parts):
    """Join the per-rank part files, in rank order, into the output file; the parts are removed."""
    lib = _abi.load()
    arr = (ctypes.c_char_p * len(parts))(*[str(p).encode() for p in parts])
    check(lib.sonics_rows_concat(str(path).encode(), arr, len(parts)))
