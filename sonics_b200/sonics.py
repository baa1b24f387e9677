"""Drop-in mirror of the reference's extension module `sonics` (sonics.pyx).

The reference CLI does `import sonics` and calls sonics.get_alleles() and
sonics.monte_carlo() (sonics:150,350,427-432).  This module exposes the same names
with the same signatures and return formats; the simulation, the lnL and the
statistics run on the GPU behind include/sonics_b200.h.  There is no CPU fallback.
"""
import numpy as np

DTYPE = np.int64


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
