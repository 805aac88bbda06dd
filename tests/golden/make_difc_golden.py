"""Mints tests/golden/difc_columns.npz from the CPU oracle (oracle/difc_oracle.c): 3 synthetic columns
of 150 levels, reduced species counts (12, 5, 8 x 4 bins, 3 x 4 bins; 2 bins diffused).  Run from the repo root:
python tests/golden/make_difc_golden.py.  The reference has no fixtures for this path; the oracle itself
is pinned by tests/test_difc_oracle.py."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import difc_oracle as dfo                 # noqa: E402
from tests.test_difc_oracle import inputs, run        # noqa: E402

c, fields = inputs(3, 11, nkc_l=2, sizes=(12, 5, 8 * 4, 3 * 4))
outs = run(dfo.difc, 60.0, c, fields)
d = {k: c[k] for k in ("atkh", "w", "am3", "detw", "deta")}
for i, ((a, _), o) in enumerate(zip(fields, outs)):
    d["f%d" % i], d["o%d" % i] = a, o
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "difc_columns.npz"), dt=60.0,
                    nproc=np.array([p for _, p in fields]), **d)
print("wrote difc_columns.npz")
