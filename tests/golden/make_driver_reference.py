"""Independent evaluation of the layer loop of SUBROUTINE kpp_driver (row a15) by executing the reference's own
Fortran statements (authoring container only, needs /root/reference):

    python tests/golden/make_driver_reference.py      # writes tests/golden/driver_reference.npz

The statements /root/reference/src/kpp.f90:4305-4306 (clip of s1, s3) and 4310-4470 (do k = n_min, n_max ... enddo)
are translated by the back end of make_sed_reference.py and executed as they stand; gas_drive / aer_drive / tot_drive
are stubs that record which mechanism the reference called for the layer, with which arguments, and what COMMON /cb_1/
held at that moment.  tests/test_driver_oracle.py holds oracle/driver_oracle.py to these records,
tests/test_gpu_driver.py the CUDA kernels.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_sed_reference import FA, namespace, translate  # noqa: E402
from mistra_b200.driver import synthetic_columns as columns  # noqa: E402

NPH = 47
ARR = ("t", "cm3", "am3", "xm1", "rho", "p", "conv2", "cm", "cloud", "photol_j", "ph_rat", "s1", "s3", "ind_gas_rev",
       "nindadv", "xadv")


def run(ncol=5, n=14, nf=10, nkc=4, j1=9, j5=4, seed=3, **cfg):
    flags = dict(halo=True, iod=True, lpbuys13_0d=False, neula=0, box=False, n_bl=2, kinv=8, dt_ch=10.0)
    flags.update(cfg)
    body = translate(4310, 4470, arrays=ARR, fname="kpp.f90")
    clip = "s1.a[s1.a < 0.0] = 0.0\ns3.a[s3.a < 0.0] = 0.0\n"          # where (s1 < 0.d0) s1=0._dp ; same for s3 (4305-4306)
    d = columns(ncol, n, nf, nkc, j1, j5, seed)
    nindadv = np.array([3, 0, 7, 3], dtype=np.int32)                      # reference species numbers, 0 = unused slot
    ind_gas_rev = np.zeros(j1 + 1, dtype=np.int32)
    ind_gas_rev[3], ind_gas_rev[7] = 5, 2                                # ... and their rows in s1
    xadv = np.array([1.0e-9, 5.0e-9, -2.0e-10, 3.0e-10])
    rec = []
    out = dict(cloud=[], s1=[], s3=[])
    for col in range(ncol):
        ns = namespace()
        T = lambda a: np.ascontiguousarray(a.T)                           # Fortran index order
        arrs = {k: T(d[k][col]) for k in ("t", "cm3", "am3", "xm1", "rho", "p", "conv2", "cm", "cloud", "photol_j", "s1", "s3")}
        for k, a in arrs.items():
            ns[k] = FA(*a.shape, data=a)
        ns.update(ph_rat=FA(NPH), nphrxn=NPH, nf=nf, n=n, chamber=False, halo=flags["halo"], iod=flags["iod"],
                  lpbuys13_0d=flags["lpbuys13_0d"], neula=flags["neula"], kinv=flags["kinv"], dd_ch=flags["dt_ch"],
                  u0=float(d["u0"][col]), airmolec=6.022e+20 / 18.0, nadvmax=len(xadv), xadv=FA(len(xadv), data=xadv.copy()),
                  nindadv=FA(len(nindadv), data=nindadv.copy()),
                  ind_gas_rev=FA((0, j1), data=ind_gas_rev.copy()),
                  n_min=flags["n_bl"] if flags["box"] else 2, n_max=flags["n_bl"] if flags["box"] else n - 1)

        def stub(mech, names):
            def f(*a):
                v = dict(zip(names, a))
                g = lambda key: float(v[key]) if key in v else float(ns[key])
                rec.append(dict(col=col, k=int(v["k"]), mech=mech, te=float(ns["te"]), air_cc=float(ns["air_cc"]),
                                h2oppm=float(ns["h2oppm"]), pk=float(ns["pk"]), air=float(v["air"]), h2o=float(v["h2o"]),
                                tkpp=float(v["tkpp"]), dt_ch=float(v["dt_ch"]), ph_rat=v["ph_rat"].a.copy(),
                                scal=[g(x) for x in ("xhal", "xiod", "xhet1", "xhet2", "xliq1", "xliq2", "xliq3", "xliq4",
                                                     "cvv1", "cvv2", "cvv3", "cvv4")],
                                passed=sorted(v.keys())))
            return f
        ns["tot_drive"] = stub(2, "tkpp dt_ch k cvv1 cvv2 cvv3 cvv4 xhal xiod xliq1 xliq2 xliq3 xliq4 xhet1 xhet2 air h2o ph_rat".split())
        ns["aer_drive"] = stub(1, "tkpp dt_ch k cvv1 cvv2 xhal xiod xliq1 xliq2 xhet1 xhet2 air h2o ph_rat".split())
        ns["gas_drive"] = stub(0, "tkpp dt_ch k xhal xiod xhet1 xhet2 air h2o ph_rat".split())
        exec(clip + body, ns)
        out["cloud"].append(arrs["cloud"].T.copy()); out["s1"].append(arrs["s1"].T.copy()); out["s3"].append(arrs["s3"].T.copy())
    fx = {("in_" + k): v for k, v in d.items()}
    fx.update(sizes=np.array([ncol, n, nf, nkc, j1, j5]), nindadv=nindadv, ind_gas_rev=ind_gas_rev, xadv=xadv,
              flags=np.array([flags["halo"], flags["iod"], flags["lpbuys13_0d"], flags["neula"], flags["box"], flags["n_bl"],
                              flags["kinv"]], dtype=np.int64), dt_ch=np.array(flags["dt_ch"]))
    fx.update({("out_" + k): np.array(v) for k, v in out.items()})
    fx.update(rec_col=np.array([r["col"] for r in rec]), rec_k=np.array([r["k"] for r in rec]),
              rec_mech=np.array([r["mech"] for r in rec]),
              rec_cb1=np.array([[r["air_cc"], r["te"], r["h2oppm"], r["pk"]] for r in rec]),
              rec_air=np.array([r["air"] for r in rec]), rec_h2o=np.array([r["h2o"] for r in rec]),
              rec_scal=np.array([r["scal"] for r in rec]), rec_ph=np.array([r["ph_rat"] for r in rec]),
              rec_t=np.array([[r["tkpp"], r["dt_ch"]] for r in rec]))
    return fx, body


if __name__ == "__main__":
    fx, body = run()
    if "--show" in sys.argv:
        print(body)
    allfx = {("a_" + k): v for k, v in fx.items()}
    fx2, _ = run(ncol=3, seed=9, halo=False, lpbuys13_0d=True, neula=1)
    allfx.update({("b_" + k): v for k, v in fx2.items()})
    fx3, _ = run(ncol=2, seed=11, iod=False, box=True, n_bl=4, kinv=2)
    allfx.update({("c_" + k): v for k, v in fx3.items()})
    path = os.path.join(HERE, "driver_reference.npz")
    np.savez_compressed(path, **allfx)
    print("wrote", path, os.path.getsize(path), "bytes;", len(fx["rec_k"]), "+", len(fx2["rec_k"]), "+", len(fx3["rec_k"]),
          "layers; mechanisms", np.bincount(fx["rec_mech"], minlength=3))
