import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure; builds oracle/libkpp_oracle.so on demand)."""
    from oracle import kpp_oracle as ko
    ko.lib()
    return ko


@pytest.fixture(scope="session")
def kpp():
    """Product library binding; the .so must have been built (no CPU fallback)."""
    from mistra_b200 import kpp as k
    if not os.path.exists(os.path.join(ROOT, "mistra_b200", "libmistra_kpp.so")):
        from mistra_b200 import build
        build.build(strict=True)
    k.library()
    # The default picks the kernel variant by batch size (on-chip for small batches).  The parity suites pin the
    # cell-per-thread kernels so that every module tests one thing; the on-chip variant and the choice by batch
    # size have their own module (tests/test_gpu_zz_onchip.py, which switches and restores the setting).
    for strict in (False, True):
        for mech in (0, 1, 2):
            k.set_kernel(mech, 0, strict=strict)
    return k


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)
