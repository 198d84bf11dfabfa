#!/usr/bin/env python
"""Validate a user-supplied D3 reference table (tad-dftd3's reference.npz content) against the reference's own goldens.

    python tools/check_d3_table.py d3_reference.npz      # keys: cn (Z+1,7), c6 (Z+1,Z+1,7,7), r4r2 (Z+1,)
    python tools/check_d3_table.py --make-fixture         # (authoring container) writes tests/golden/d3_implied.json

The D3(BJ) reference data is third-party data that exists nowhere offline, so neither the oracle nor the CUDA path can be
pinned for dispersion here.  What the reference tree DOES hold are tblite totals that include D3:
  * total energies  test/test_singlepoint/samples.py ("egfn1") -> implied dispersion energy
        E_disp = E_total(tblite) - E_nodisp(oracle, tblite's eV, SCF converged to 1e-10)   [tests/golden/d3_implied.json]
  * total gradients test/test_singlepoint/refs/gfn1/grad.npz (float32)                     [tests/golden/reference.npz total_grad/*]
Given a real table this script asserts the oracle's D3 energies against the implied ones to 1e-8 Eh and the total gradients
to 2e-6 Eh/bohr (float32 goldens); tests/test_d3_table.py runs the same check (and the CUDA path on a GPU) when
DXTB_B200_D3_REFERENCE points to a table, and is skipped with a loud reason otherwise.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import gfn1_oracle as O  # noqa: E402

TBLITE_EV2AU = 1.0 / 27.21138505
TIGHT = dict(x_atol=1e-10, x_atol_max=1e-10)
FIXTURE = ROOT / "tests" / "golden" / "d3_implied.json"
E_TOL, G_TOL = 1e-8, 2e-6


def _mols():
    return json.load(open(ROOT / "tests" / "golden" / "molecules.json"))


def make_fixture():
    en = json.load(open(ROOT / "tests" / "golden" / "energies.json"))["total_gfn1_tblite"]
    mols = _mols()
    par = O.params()
    old, par.ev2au = par.ev2au, TBLITE_EV2AU
    out = {}
    try:
        for name, etot in en.items():
            m = mols[name]
            r = O.singlepoint(np.array(m["numbers"]), np.array(m["positions"]), float(m["charge"]), opts=dict(exclude=("disp",), **TIGHT))
            out[name] = {"e_total_tblite": etot, "e_nodisp_oracle": r.energy, "e_disp_implied": etot - r.energy, "nat": len(m["numbers"])}
            print(f"{name:12s} E_disp(implied) = {etot - r.energy:+.10f} Eh", flush=True)
    finally:
        par.ev2au = old
    FIXTURE.write_text(json.dumps({"source": "test/test_singlepoint/samples.py egfn1 minus oracle (exclude disp, tblite eV, x_atol 1e-10)",
                                   "molecules": out}, indent=1))


def check(table: dict, names=None, gradients=True) -> dict:
    """max |dE_disp|, max |dG_total| of the oracle with ``table`` against the goldens."""
    fx = json.load(open(FIXTURE))["molecules"]
    gold = np.load(ROOT / "tests" / "golden" / "reference.npz")
    mols = _mols()
    par = O.params()
    old, par.ev2au = par.ev2au, TBLITE_EV2AU
    rep = {}
    try:
        for name in names or fx:
            m = mols[name]
            z, p, c = np.array(m["numbers"]), np.array(m["positions"]), float(m["charge"])
            want_g = gradients and f"total_grad/{name}" in gold.files
            r = O.singlepoint(z, p, c, opts=dict(**TIGHT), grad=want_g, d3_table=table)
            rep[name] = {"dE_disp": abs(r.e_disp - fx[name]["e_disp_implied"]),
                         "dG": float(np.abs(r.gradient - gold[f"total_grad/{name}"]).max()) if want_g else None}
    finally:
        par.ev2au = old
    return rep


def main():
    if "--make-fixture" in sys.argv:
        make_fixture()
        return 0
    if len(sys.argv) < 2:
        print(__doc__)
        return 2
    with np.load(sys.argv[1]) as f:
        table = {k: f[k] for k in ("cn", "c6", "r4r2")}
    rep = check(table)
    ok = True
    for name, r in rep.items():
        good = r["dE_disp"] < E_TOL and (r["dG"] is None or r["dG"] < G_TOL)
        ok = ok and good
        print(f"{name:12s} |dE_disp| = {r['dE_disp']:.2e}  |dG_total| = {r['dG'] if r['dG'] is None else format(r['dG'], '.2e')}  {'ok' if good else 'MISMATCH'}")
    print("table", "ACCEPTED" if ok else "REJECTED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
