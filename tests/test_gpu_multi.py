"""mistra_kpp_integrate_multi: the batch over several GPUs of the box from one process (one host thread per slice, one
lock per device), and the page-locked host memory helpers of the C ABI."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def test_two_slices_on_one_device_equal_the_single_call(kpp, cuda_device):
    """Runs on a 1-GPU box: the same device listed twice = two host threads that take the device lock in turn."""
    var, fix, rc = util.random_cells("aer", 700, 31)
    ref, ierr_r, stats_r, hexit_r, texit_r = kpp.integrate(1, rc, fix, var)
    out, ierr, stats, hexit, texit = kpp.integrate_multi(1, rc, fix, var, devices=[0, 0])
    assert np.array_equal(out, ref) and np.array_equal(ierr, ierr_r) and np.array_equal(stats, stats_r)
    assert np.array_equal(hexit, hexit_r) and np.array_equal(texit, texit_r)
    out3, ierr3, _, _, _ = kpp.integrate_multi(1, rc, fix, var, devices=[0, 0, 0])    # ragged slices: 233 + 233 + 234
    assert np.array_equal(out3, ref)
    out0, ierr0, _, _, _ = kpp.integrate_multi(1, rc[:0], fix[:0], var[:0], devices=[0, 0])
    assert out0.shape == (0, var.shape[1])
    with pytest.raises(kpp.KppError):
        kpp.integrate_multi(1, rc, fix, var, devices=[0, 99])


def test_all_visible_devices(kpp, cuda_device):
    import torch
    n = kpp.device_count()
    assert n == torch.cuda.device_count()
    var, fix, rc = util.random_cells("gas", 4000, 32)
    ref, ierr_r, stats_r, _, _ = kpp.integrate(0, rc, fix, var)
    out, ierr, stats, _, _ = kpp.integrate_multi(0, rc, fix, var)            # ndev <= 0: every visible device
    assert np.array_equal(out, ref) and np.array_equal(ierr, ierr_r) and np.array_equal(stats, stats_r)
    if n >= 2:
        out2, _, _, _, _ = kpp.integrate_multi(0, rc, fix, var, devices=[1, 0])
        assert np.array_equal(out2, ref)
    assert torch.cuda.current_device() == 0                                   # the caller's device is restored


def test_pinned_helpers(kpp, cuda_device):
    var, fix, rc = util.random_cells("gas", 300, 33)
    ref = kpp.integrate(0, rc, fix, var)[0]
    pv = kpp.PinnedArray(var.shape)
    pv.array[:] = var
    out = kpp.integrate(0, rc, fix, pv.array, out=pv.array)[0]
    assert np.array_equal(out, ref)
    r2 = np.ascontiguousarray(rc.copy())
    kpp.host_register(r2)
    try:
        assert np.array_equal(kpp.integrate(0, r2, fix, var)[0], ref)
    finally:
        kpp.host_unregister(r2)
