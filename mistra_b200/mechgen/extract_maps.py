"""Extracts the gather / scatter statements of the *_drive routines between the model arrays sl1 / sion1
and the KPP concentration vector C: /root/reference/src/aer_mk.dat and aer_km.dat (included by aer_drive
and tot_drive, aer.f:178 / 245, tot.f:252 / 631) and the inline l3 / l4 statements of tot_drive
(tot.f:256-597, 637-979), into mistra_b200/mech/drive_maps.json.  Build-time only, run in the authoring
container: python -m mistra_b200.mechgen.extract_maps /root/reference.  Entries are
[species name, array ('sl1' | 'sion1'), j (1-based, j2 / j3 resolved), kc (1-based)]."""
import json
import os
import re
import sys

J1_FAKE, J3 = 96, 25                       # global_params.f90:93-98
J2 = J1_FAKE + J3

MK = re.compile(r"^\s+C\(ind_(\w+)\)\s*=\s*(sl1|sion1)\(([^,]+),(\d),k\)")
KM = re.compile(r"^\s+(sl1|sion1)\(([^,]+),(\d),k\)\s*=\s*C\(ind_(\w+)\)")


def _j(expr):
    v = eval(expr.replace("j2", str(J2)).replace("j3", str(J3)), {"__builtins__": {}})
    assert isinstance(v, int) and v >= 1, expr
    return v


def _scan(lines):
    mk, km = [], []
    for ln in lines:
        if ln[:1] in "cC!*":
            continue
        m = MK.match(ln)
        if m:
            mk.append([m.group(1), m.group(2), _j(m.group(3)), int(m.group(4))])
        m = KM.match(ln)
        if m:
            km.append([m.group(4), m.group(1), _j(m.group(2)), int(m.group(3))])
    return mk, km


def main(ref):
    src = os.path.join(ref, "src")
    mk12, _ = _scan(open(os.path.join(src, "aer_mk.dat")).read().splitlines())
    _, km12 = _scan(open(os.path.join(src, "aer_km.dat")).read().splitlines())
    mk34, km34 = _scan(open(os.path.join(src, "tot.f")).read().splitlines())
    out = {"j2": J2, "j3": J3, "source": "aer_mk.dat, aer_km.dat, tot.f",
           "aer": {"gather": mk12, "scatter": km12}, "tot": {"gather": mk12 + mk34, "scatter": km12 + km34}}
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(here, "mech", "drive_maps.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    for k in ("aer", "tot"):
        print(k, len(out[k]["gather"]), "gather,", len(out[k]["scatter"]), "scatter statements")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
