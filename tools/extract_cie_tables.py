"""Writes pbrt-v3-rs_b200/data/cie_xyz.bin: the CIE 1931 2-degree colour-matching functions x, y, z at 360..830 nm in 1 nm
steps (3 x 471 little-endian f32, x first), as the reference tabulates them (core/src/spectrum/cie.rs: CIE_X, CIE_Y, CIE_Z;
the same published tables pbrt-v3 ships).  They are data, needed to turn a "blackbody" parameter into RGB the way
ParamSet::add_blackbody_spectrum -> RGBSpectrum::from(samples) does (paramset/mod.rs:236-249, spectrum/rgb_spectrum.rs:76-103)."""
import os
import re
import sys

import numpy as np

SRC = "/root/reference/core/src/spectrum/cie.rs"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pbrt-v3-rs_b200", "data", "cie_xyz.bin")


def table(text, name):
    start = text.index("pub const %s:" % name)
    body = text[text.index("= [", start) + 3:text.index("];", start)]
    v = np.array([float(t) for t in re.findall(r"[-+]?\d*\.\d+(?:[eE][-+]?\d+)?|[-+]?\d+(?:[eE][-+]?\d+)", body)], dtype=np.float64)
    assert v.size == 471, (name, v.size)
    return v.astype("<f4")


def main():
    text = open(SRC).read()
    x, y, z = table(text, "CIE_X"), table(text, "CIE_Y"), table(text, "CIE_Z")
    assert abs(float(y.astype(np.float64).sum()) - 106.856895) < 1e-3  # CIE_Y_INTEGRAL
    np.concatenate([x, y, z]).tofile(OUT)
    print("wrote", OUT)


if __name__ == "__main__":
    sys.exit(main())
