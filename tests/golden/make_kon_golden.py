"""Regenerates tests/golden/kon_layers.npz: inputs and CPU-oracle outputs of subkon/advec
(str.f90:4987-5204, 5321-5516) for 8 synthetic humid layers.
Run from the repo root:  python tests/golden/make_kon_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mistra_b200 import kon              # noqa: E402  (grid + synthetic inputs only)
from oracle import kon_oracle as ko      # noqa: E402

g = kon.kon_grid()
d = kon.synthetic_layers(g, 8, seed=31)
ffk, to, xm1o, st = ko.subkon(g, 10.0, d["ffk"], d["totr"], d["dfdt"], d["feualt"], d["pp"], d["to"], d["tn"],
                              d["xm1o"], d["xm1n"], d["kr"])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "kon_layers.npz"), dt=10.0, ffk_out=ffk, to_out=to,
                    xm1o_out=xm1o, status=st, **d)
print("wrote kon_layers.npz", st)
