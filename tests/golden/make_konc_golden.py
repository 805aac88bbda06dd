"""Regenerates tests/golden/konc_layers.npz: inputs and CPU-oracle outputs of SUBROUTINE konc
(kpp.f90:3370-3585) for 12 synthetic layers.
Run from the repo root:  python tests/golden/make_konc_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mistra_b200 import konc             # noqa: E402  (synthetic inputs only)
from oracle import konc_oracle as kco    # noqa: E402

d = konc.synthetic_sums(12, seed=77)
sl1, sion1, warn = kco.konc(d["ka"], d["sums"], d["vol2"], d["pntot"], d["sl1"], d["sion1"])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "konc_layers.npz"), ka=d["ka"], vol2=d["vol2"],
                    pntot=d["pntot"], sl1=d["sl1"], sion1=d["sion1"], sl1_out=sl1, sion1_out=sion1, warn=warn,
                    **d["sums"])
print("wrote konc_layers.npz", warn.sum(0))
