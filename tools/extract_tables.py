#!/usr/bin/env python
"""Extract the two numeric DATA tables the hot path needs into compact binary files.

 * Sobol direction matrices (Joe & Kuo 2008 "new-joe-kuo-6.21201" numbers, tabulated by
   L. Gruenschloss, MIT licence) -- reference table at src/sobol.inc:32-35 (1024 dims x 52 bits).
   We keep the first NUM_DIMS dimensions (a path uses 2 + 7*max_bounces main-sampler dimensions and
   2 + 4 + 7*max_bounces edge-sampler dimensions per sample, see src/pathtracer.cpp:260-340,:505-641,:788-882).
 * Linearly-transformed-cosine matrices fitted to the Blinn-Phong microfacet BRDF
   (Heitz et al. fitting code; reference table at src/ltc.inc:14, 128x128x9 float) used by the
   secondary-edge importance sampler (src/edge.cpp:803-814).

These are numeric tables (data), not code; they are stored as raw little-endian arrays in
redner_b200/data/.  Run once in the build container (where /root/reference exists):

    python tools/extract_tables.py
"""
import os
import re
import sys

import numpy as np

REF = os.environ.get("REDNER_REF", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "redner_b200", "data")
NUM_DIMS = 1024  # all 1024 dims x 52 x 8 B = 426 KB


def main():
    os.makedirs(OUT, exist_ok=True)
    txt = open(os.path.join(REF, "src", "sobol.inc")).read()
    body = txt[txt.index("matrices_["):]
    vals = re.findall(r"0x([0-9a-fA-F]+)ULL", body)
    mat = np.array([int(v, 16) for v in vals], dtype=np.uint64)
    assert mat.size == 1024 * 52, mat.size
    mat = mat.reshape(1024, 52)[:NUM_DIMS]
    mat.tofile(os.path.join(OUT, "sobol_joe_kuo_%dx52_u64.bin" % NUM_DIMS))

    txt = open(os.path.join(REF, "src", "ltc.inc")).read()
    start = txt.index("tabM_[")
    # the table has an #if/#else with two variants on some versions; take the first brace block
    body = txt[txt.index("{", start):]
    nums = re.findall(r"-?\d+\.\d+(?:[eE][-+]?\d+)?", body)
    tab = np.array([float(v) for v in nums[:128 * 128 * 9]], dtype=np.float32)
    assert tab.size == 128 * 128 * 9, tab.size
    tab.tofile(os.path.join(OUT, "ltc_blinn_phong_128x128x9_f32.bin"))
    print("wrote", os.listdir(OUT))


if __name__ == "__main__":
    sys.exit(main())
