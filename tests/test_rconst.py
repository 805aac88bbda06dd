"""Host-side Update_RCONST_x restatement (libmistra_rconst.so) against direct
numpy evaluations of the rate laws it cites (kpp.f90:7127-8373) - CPU."""
import numpy as np

from mistra_b200 import rconst as rc
from mistra_b200 import synthetic
from mistra_b200.mechgen import mech as mechmod


def _ens():
    return synthetic.GasEnsemble(2, f32_literals=0)


def test_arrhenius_photolysis_and_switches():
    e = _ens()
    R = e.rconst()
    te, aircc = e.cb1[:, 1], e.cb1[:, 0]
    conv1 = synthetic.CONV1
    assert np.allclose(R[:, 0], 3.2e-11 * np.exp(67 / te) * conv1, rtol=1e-14)       # gas.f:316 farr
    assert np.allclose(R[:, 1], 2.4e-10 * conv1, rtol=1e-15)                          # gas.f:317
    assert np.allclose(R[:, 6], 6.0e-34 * (te / 300) ** -2.6 * conv1 * conv1, rtol=1e-13)  # farr_sp
    assert np.array_equal(R[:, 17], e.ph_rat[:, 2])                                   # RCONST(18)=ph_rat(3)
    assert np.allclose(R[:, 89], e.ph_rat[:, 2] / 9.0, rtol=1e-15)                    # RCONST(90)
    # Troe fall-off atk_3 (kpp.f90:7171) for RCONST(15)
    a0 = 6.9e-31 * aircc * (te / 300.0) ** -0.8
    b0 = 2.6e-11 * (te / 300.0) ** 0.0
    k = (a0 / (1 + a0 / b0)) * 0.5 ** (1 / (1 + np.log10(a0 / b0) ** 2))
    assert np.allclose(R[:, 14], k * conv1, rtol=1e-13)
    # halogen switches zero the halogen block (gas.f:427-...)
    e2 = synthetic.GasEnsemble(2, halo=False, iod=False, f32_literals=0)
    R2 = e2.rconst()
    m = mechmod.load("gas")
    hal = [i for i, ex in enumerate(m.d["rconst"]) if "S_xhal" in ex or "S_xiod" in ex]
    assert len(hal) > 100 and (R2[:, hal] == 0).all() and (R[:, hal] != 0).any()


def test_dry_uptake_rates_follow_fdhetg():
    e = _ens()
    R = e.rconst()
    i = e.idx
    # RCONST(325) = xhet1*fdhetg(1,2) = xkmtd(N2O5,1)*cwd(1)   (gas.f:658, kpp.f90:8198)
    assert np.allclose(R[:, 324], e.yxkmtd[:, 0, i["N2O5"]] * e.ycwd[:, 0], rtol=1e-15)
    assert np.allclose(R[:, 330], e.yxkmtd[:, 1, i["H2SO4"]] * e.ycwd[:, 1], rtol=1e-15)
    assert (R[:, 323] >= 0).all()


def test_f32_literals_shift_results_at_1e8_level():
    a = synthetic.GasEnsemble(1, f32_literals=1)
    b = synthetic.GasEnsemble(1, f32_literals=0)
    Ra, Rb = a.rconst(), b.rconst()
    rel = np.abs(Ra - Rb) / np.maximum(np.abs(Rb), 1e-300)
    assert rel.max() < 1e-6 and rel.max() > 0


def test_species_index_lookup():
    assert rc.spc_index(0, "HNO3") == mechmod.load("gas").spc_names.index("HNO3")
    assert rc.spc_index(2, "Clml4") == mechmod.load("tot").spc_names.index("Clml4")
    assert rc.spc_index(0, "nope") == -1
