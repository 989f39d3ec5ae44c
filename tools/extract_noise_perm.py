"""Writes pbrt-v3-rs_b200/data/noise_perm.bin: the 256-entry permutation of Ken Perlin's "Improved Noise" reference
implementation (SIGGRAPH 2002), which the reference tabulates twice over as NOISE_PERM (core/src/texture/common.rs:8-36)
and which its "dots" / fbm / wrinkled / marble textures index.  It is a published constant, data rather than code; this
script lifts the numbers out of the mounted reference once, checks that the table is the permutation repeated and that it
is a permutation of 0..255, and the 256 bytes are committed."""
import os
import re
import sys

import numpy as np

SRC = "/root/reference/core/src/texture/common.rs"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pbrt-v3-rs_b200", "data", "noise_perm.bin")


def main():
    text = open(SRC).read()
    start = text.index("const NOISE_PERM:")
    body = text[text.index("= [", start) + 3:text.index("];", start)]
    vals = np.array([int(t) for t in re.findall(r"\d+", body)], dtype=np.int64)
    assert vals.size == 512, vals.size
    assert np.array_equal(vals[:256], vals[256:])
    assert sorted(vals[:256].tolist()) == list(range(256))
    vals[:256].astype(np.uint8).tofile(OUT)
    print("wrote", OUT)


if __name__ == "__main__":
    sys.exit(main())
