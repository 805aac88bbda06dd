"""Regenerates tests/golden/cwrc_layers.npz: inputs and CPU-oracle outputs of SUBROUTINE cw_rc
(kpp.f90:2152-2414) for 10 synthetic layers.
Run from the repo root:  python tests/golden/make_cwrc_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.test_cwrc_oracle import inputs   # noqa: E402  (synthetic inputs)
from oracle import cwrc_oracle as cwo       # noqa: E402

g, ff, feu, cloud = inputs(10, 41)
rc, cw, cm, conv2 = cwo.cw_rc(g, ff, feu, cloud)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cwrc_layers.npz"), ff=ff, feu=feu, cloud=cloud,
                    rc=rc, cw=cw, cm=cm, conv2=conv2)
print("wrote cwrc_layers.npz", (cm > 0).sum(0))
