"""Regenerates tests/golden/bins_layers.npz: inputs and CPU-oracle outputs of the 2-D bin
redistribution (str.f90:5916-6134) for 12 synthetic layers on the default particle grid.
Run from the repo root:  python tests/golden/make_bins_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mistra_b200 import bins            # noqa: E402  (grid + synthetic inputs only)
from oracle import bins_oracle as bo    # noqa: E402

g = bins.particle_grid()
d = bins.synthetic_layers(g, 12, seed=77, growth=0.08)
d["cm"][3] = 0.0                                   # a dry layer: nothing happens
d["sion1_new"][5] = d["sion1"][5]                  # no mass change: identity
sap, smp, so = bo.snapshot(g, d["ff"], d["cm"], d["sion1"])
ff2, si2, sl2, nw = bo.redistribute(g, d["ff"], d["cm"], d["cw"], sap, smp, so, d["sion1_new"], d["sl1"])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bins_layers.npz"),
                    ff=d["ff"], cm=d["cm"], cw=d["cw"], sion1=d["sion1"], sion1_new=d["sion1_new"], sl1=d["sl1"],
                    sap=sap, smp=smp, sion1o=so, ff_out=ff2, sion1_out=si2, sl1_out=sl2, nwarn=nw,
                    grid_args=np.array([0.005, 15.0, 0.005, 150.0]))
print("wrote bins_layers.npz", ff2.shape)
