"""Mints tests/golden/fastkmt_layers.npz from the CPU oracle (oracle/fastkmt_oracle.c): 24 synthetic aer
layers.  Run from the repo root: python tests/golden/make_fastkmt_golden.py.  The reference has no
fixtures for this path; the oracle itself is pinned by tests/test_fastkmt_oracle.py."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import fastkmt_oracle as fko          # noqa: E402
from tests.test_fastkmt_oracle import call, inputs   # noqa: E402

x = inputs(24, 11, "aer")
xk, vt = call(fko.fast_k_mt, x)
out = {k: x[k] for k in ("lex", "ff", "freep", "t", "p", "cw", "cm", "alpha", "vmean", "xkmt", "vt")}
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fastkmt_layers.npz"),
                    xkmt_out=xk, vt_out=vt, **out)
print("wrote fastkmt_layers.npz", xk.shape, vt.shape)
