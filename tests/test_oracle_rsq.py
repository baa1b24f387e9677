"""oracle_descriptors() (identity, r_squared of a readout) against the reference's own rsq() and identity
expression (sonics.pyx:273-291, :382-383): tests/golden/rsq.json was produced by calling the compiled
reference (oracle/_ref) on 400 random (reads, readout) pairs, 40 of them with ss_tot == 0 (the 1e-16 guard)."""
import numpy as np

from oracle import oracle as orc
from tests.util import load


def _arrays(c):
    reads = np.zeros(500, dtype=np.int64)
    readout = np.zeros(500, dtype=np.int64)
    for k, v in c["reads"].items():
        reads[int(k)] = v
    for k, v in c["readout"].items():
        readout[int(k)] = v
    return reads, readout


def test_descriptors_bit_exact_against_reference_rsq():
    cases = load("rsq.json")
    guard = 0
    for c in cases:
        reads, readout = _arrays(c)
        ident, r2 = orc.descriptors(reads, readout)
        assert ident == float.fromhex(c["identity"])
        assert r2 == float.fromhex(c["rsq_f32"])  # through the C float of sonics.pyx:300
        guard += abs(float.fromhex(c["rsq"])) > 1e10
    assert guard >= 20


def test_descriptors_empty_readout_branch():
    # sonics.pyx:376-380: nothing sequenced -> identity 0, r_squared -999999
    reads = orc.parse_reads("9|1;10|1")
    ident, r2 = orc.descriptors(reads, np.zeros(500, dtype=np.int64))
    assert ident == 0.0 and r2 == -999999.0
