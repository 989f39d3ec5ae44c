"""Writes pbrt-v3-rs_b200/data/sobol_matrices_32.bin: the 1024 x 52 u32 generator matrices of the Sobol' sequence that
the reference's SobolSampler indexes (core/src/sobol_matrices.rs: SOBOL_MATRICES_32; the Joe-Kuo direction numbers as
tabulated by L. Gruenschloss' sobol generator, the same table pbrt-v3 ships).  They are data, not code, and cannot be
regenerated offline (the direction-number file is not in this image), so this script lifts the numbers out of the mounted
reference once and the binary is committed.  The van der Corput / inverse matrices sobol_interval_to_index needs are
NOT copied: oracle and product derive them from dimensions 0 and 1 (see oracle/oracle_sobol.h)."""
import os
import re
import sys

import numpy as np

SRC = "/root/reference/core/src/sobol_matrices.rs"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pbrt-v3-rs_b200", "data", "sobol_matrices_32.bin")


def main():
    text = open(SRC).read()
    start = text.index("pub const SOBOL_MATRICES_32")
    body = text[text.index("= [", start) + 3:text.index("];", start)]
    vals = [int(t, 16) if t.lower().startswith("0x") else int(t) for t in re.findall(r"0x[0-9a-fA-F_]+|\d[\d_]*", body.replace("_", ""))]
    a = np.array(vals, dtype=np.uint64)
    assert a.size == 1024 * 52, a.size
    assert a.max() < 2 ** 32
    a.astype("<u4").tofile(OUT)
    print("wrote", OUT, a.size, "values")


if __name__ == "__main__":
    sys.exit(main())
