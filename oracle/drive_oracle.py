"""numpy restatement of the gather / scatter halves of aer_drive / tot_drive / gas_drive
(/root/reference/src/aer.f:146-178, 233-245; kpp.f90:4472-4477) for a batch of layers.
TEST INFRASTRUCTURE ONLY - see oracle/kpp_oracle.h for who may import this."""
import numpy as np


def _rows(s1, s3, sl1, sion1):
    n = s1.shape[0]
    return [s1, s3, sl1.reshape(n, -1), sion1.reshape(n, -1)]


def gather(mp, layer, s1, s3, sl1, sion1, air, h2o, cvv, var, fix, clamp_liquid=True, f32_literals=True):
    """Returns (var, fix, sl1, sion1) after the gather; inputs are not modified."""
    s1, s3, sl1, sion1, var, fix = (np.array(x, dtype=np.float64) for x in (s1, s3, sl1, sion1, var, fix))
    if clamp_liquid:
        sl1[layer] = np.maximum(0.0, sl1[layer])
        sion1[layer] = np.maximum(0.0, sion1[layer])
    rows = _rows(s1, s3, sl1, sion1)
    nvar = mp["nvar"]
    for kp, ar, of in zip(mp["kpp"], mp["arr"], mp["off"]):          # statement order of the reference
        v = rows[ar][layer, of]
        if kp <= nvar:
            var[:, kp - 1] = v
        else:
            fix[:, kp - nvar - 1] = v
    lit = (lambda x: float(np.float32(x))) if f32_literals else float
    if mp["indf_o2"]:
        fix[:, mp["indf_o2"] - 1] = lit(0.21) * air
    if mp["indf_h2o"]:
        fix[:, mp["indf_h2o"] - 1] = h2o
    if mp["indf_n2"]:
        fix[:, mp["indf_n2"] - 1] = lit(0.79) * air
    for b, p in enumerate(mp["indf_h2ol"]):
        if p:
            cv = cvv[:, b]
            fix[:, p - 1] = np.where(cv > 0, lit(55.55) / np.where(cv > 0, cv, 1.0), 0.0)
    return var, fix, sl1, sion1


def scatter(mp, layer, s1, s3, sl1, sion1, var, fix, clip_negative=True):
    """Returns (s1, s3, sl1, sion1) after the scatter; inputs are not modified."""
    s1, s3, sl1, sion1 = (np.array(x, dtype=np.float64) for x in (s1, s3, sl1, sion1))
    rows = _rows(s1, s3, sl1, sion1)                                  # views: writes go to the copies
    nvar = mp["nvar"]
    for kp, ar, of in zip(mp["kpp"], mp["arr"], mp["off"]):
        rows[ar][layer, of] = var[:, kp - 1] if kp <= nvar else fix[:, kp - nvar - 1]
    if clip_negative:
        for r in rows:
            r[layer] = np.where(r[layer] < 0.0, 0.0, r[layer])
    return s1, s3, sl1, sion1
