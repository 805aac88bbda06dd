"""CPU oracle of the layer loop of SUBROUTINE kpp_driver (numpy restatement of /root/reference/src/kpp.f90:4305-4470).
TEST INFRASTRUCTURE ONLY - see oracle/kpp_oracle.h for who may import this.  Pinned by tests/golden/driver_reference.npz
(the reference's own statements executed by tests/golden/make_driver_reference.py)."""
import numpy as np

AIRMOLEC = 6.022e+20 / 18.0                      # kpp.f90:4278
CONV1 = 6.022140857e+23 * 1.e-6                  # constants.f90:45


def layers(cfg, u0, t, p, rho, cm3, am3, xm1, conv2, cm, cloud, photol_j, adv_row=None, xadv=None, s1=None, s3=None):
    """cfg: dict(nf, halo, iod, lpBuys13_0D, neula, box, n_bl, kinv, dt_ch).  Arrays as in include/mistra_driver.h.
    Returns a dict with cb1, scal, ph_rat, air, h2o, cvv, mech (rows L = col * n + k - 1; untouched rows zero,
    mech -1), layers (list of three ascending index arrays), cloud, s1, s3 (updated copies)."""
    ncol, n = t.shape
    nkc, nph = cm.shape[2], photol_j.shape[2]
    n_min, n_max = (cfg["n_bl"], cfg["n_bl"]) if cfg["box"] else (2, n - 1)
    k = np.arange(1, n + 1)[None, :] * np.ones((ncol, 1), dtype=np.int64)
    on = (k >= n_min) & (k <= n_max)
    s1 = None if s1 is None else np.where(s1 < 0.0, 0.0, s1)                          # 4305-4306
    s3 = None if s3 is None else np.where(s3 < 0.0, 0.0, s3)
    h2o = xm1 * rho / 1.8e-2                                                          # 4318-4320
    h2o_cc = xm1 * AIRMOLEC * rho
    h2oppm = h2o_cc * 1.e6 / cm3
    cb1 = np.stack([cm3, t, h2oppm, p], axis=-1)
    u0min = 1.75e-2 if cfg["lpBuys13_0D"] else 3.48e-2                                # 4344-4348
    ph = np.zeros((ncol, n, nph))
    ph[:, 1:] = (photol_j[:, :-1] + photol_j[:, 1:]) / 2.0                            # 4354
    ph *= (np.asarray(u0) >= u0min)[:, None, None]
    xhal = 1.0 if cfg["halo"] else 0.0                                                # 4365-4371
    xiod = 1.0 if (cfg["halo"] and cfg["iod"]) else 0.0
    xliq = np.zeros((ncol, n, 4))
    cvv = np.zeros((ncol, n, 4))
    xliq[..., :nkc] = ((k < cfg["nf"])[..., None] & (cm != 0.0)).astype(np.float64)   # 4374-4390
    cvv[..., :nkc] = conv2
    new_cloud = np.where(on[..., None], xliq[..., :nkc] == 1.0, cloud != 0).astype(np.int32)      # 4392-4412
    liq12 = (xliq[..., 0] == 1.0) | (xliq[..., 1] == 1.0)
    mech = np.where(liq12, np.where((xliq[..., 2] == 1.0) | (xliq[..., 3] == 1.0), 2, 1), 0)      # 4452-4468
    xhet = 1.0 - xliq[..., :2]                                                        # 4435-4438 (gas layers: both 1 anyway)
    if cfg["neula"] == 0 and s1 is not None and adv_row is not None:                  # 4441-4449
        sel = on & (k <= cfg["kinv"])
        for r, xa in zip(adv_row, xadv):
            if r >= 0:
                s1[..., r] = np.where(sel, s1[..., r] + xa * cfg["dt_ch"] * am3 / 86400., s1[..., r])
    scal = np.concatenate([np.full((ncol, n, 1), CONV1), np.full((ncol, n, 1), xhal), np.full((ncol, n, 1), xiod), xhet, xliq, cvv],
                          axis=-1)
    z = lambda a: np.where(on.reshape(on.shape + (1,) * (a.ndim - 2)), a, 0.0).reshape((ncol * n,) + a.shape[2:])
    mech = np.where(on, mech, -1).reshape(-1).astype(np.int32)
    return dict(cb1=z(cb1), scal=z(scal), ph_rat=z(ph), air=z(am3), h2o=z(h2o), cvv=z(cvv), mech=mech,
                layers=[np.nonzero(mech == m)[0].astype(np.int64) for m in range(3)], cloud=new_cloud, s1=s1, s3=s3)
