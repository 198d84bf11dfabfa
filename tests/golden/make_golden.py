#!/usr/bin/env python
"""Collect the reference's own golden vectors for the GFN1 hot path into small committed fixtures.

Run in the authoring container (needs /root/reference as a data source; no reference code is imported):

    python tests/golden/make_golden.py

Writes
  tests/golden/molecules.json   geometries (bohr) from test/test_singlepoint/mols/*/coord and
                                examples/molecules/*, plus the caffeine geometry used by bench.py
  tests/golden/reference.npz    float32 goldens of test_overlap/overlap.npz, test_hamiltonian/h0.npz,
                                test_scf/grad.npz (tblite), test_hamiltonian/grad_no_overlap.npz (*_dcn),
                                test_hamiltonian/grad.npz (H0 part of the nuclear gradient)
  tests/golden/energies.json    tblite fp64 literals: SCF energies (test_scf/samples.py), total
                                energies (test_singlepoint/samples.py), EEQ known answer (test_scf/test_guess.py)
"""
import json
import re
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
AA2AU = 1.0 / 0.529177210903

SYM = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr "
       "Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir "
       "Pt Au Hg Tl Pb Bi Po At Rn").split()
S2Z = {s.lower(): i for i, s in enumerate(SYM)}

CAFFEINE_XYZ = """
H -3.3804130 -1.1272367 0.5733036
N 0.9668296 -1.0737425 -0.8198227
C 0.0567293 0.8527195 0.3923156
N -1.3751742 -1.0212243 -0.0570552
C -1.2615018 0.2590713 0.5234135
C -0.3068337 -1.6836331 -0.7169344
C 1.1394235 0.1874122 -0.2700900
N 0.5602627 2.0839095 0.8251589
O -0.4926797 -2.8180554 -1.2094732
C -2.6328073 -1.7303959 -0.0060953
O -2.2301338 0.7988624 1.0899730
H 2.5496990 2.9734977 0.6229590
C 2.0527432 -1.7360887 -1.4931279
H -2.4807715 -2.7269528 0.4882631
H -3.0089039 -1.9025254 -1.0498023
H 2.9176101 -1.8481516 -0.7857866
H 2.3787863 -1.1211917 -2.3743655
H 1.7189877 -2.7489920 -1.8439205
C -0.1518450 3.0970046 1.5348347
C 1.8934096 2.1181245 0.4193193
N 2.2861252 0.9968439 -0.2440298
H -0.1687028 4.0436553 0.9301094
H 0.3535322 3.2979060 2.5177747
H -1.2074498 2.7537592 1.7203047
"""


def read_coord(p):
    nums, pos, on = [], [], False
    for line in open(p):
        if line.startswith("$coord"):
            on = True
            continue
        if line.startswith("$"):
            if on:
                break
            continue
        if on and line.strip():
            t = line.split()
            pos.append([float(x) for x in t[:3]])
            nums.append(S2Z[t[3].lower()])
    return nums, pos


def read_xyz(p, first=True):
    lines = open(p).read().splitlines()
    n = int(lines[0])
    nums, pos = [], []
    for line in lines[2 : 2 + n]:
        t = line.split()
        nums.append(S2Z[t[0].lower()])
        pos.append([float(x) * AA2AU for x in t[1:4]])
    return nums, pos


def atom_refs():
    txt = open(REF / "test/test_scf/test_elements_gfn1.py").read()
    a = txt.index("ref = torch.tensor(")
    blk = txt[a : txt.index("ref_cation", a)]
    return [float(x) for x in re.findall(r"-?\d+\.\d+(?:e[-+]?\d+)?", blk)]


def main():
    mols = {}
    base = REF / "test/test_singlepoint/mols"
    for d in sorted(base.iterdir()):
        nums, pos = read_coord(d / "coord")
        chrg = float((d / ".CHRG").read_text()) if (d / ".CHRG").exists() else 0.0
        mols[d.name] = {"numbers": nums, "positions": pos, "charge": chrg, "source": f"test/test_singlepoint/mols/{d.name}/coord"}
    ex = REF / "examples/molecules"
    for name in ("capsaicin", "nicotine", "lih"):
        nums, pos = read_xyz(ex / f"{name}.xyz")
        mols[name] = {"numbers": nums, "positions": pos, "charge": 0.0, "source": f"examples/molecules/{name}.xyz"}
    for name in ("h2o", "vancoh2", "sh3"):
        nums, pos = read_coord(ex / f"{name}.coord")
        mols[f"ex_{name}"] = {"numbers": nums, "positions": pos, "charge": 0.0, "source": f"examples/molecules/{name}.coord"}
    nums, pos = [], []
    for line in CAFFEINE_XYZ.strip().splitlines():
        t = line.split()
        nums.append(S2Z[t[0].lower()])
        pos.append([float(x) * AA2AU for x in t[1:4]])
    mols["caffeine"] = {"numbers": nums, "positions": pos, "charge": 0.0,
                        "source": "standard optimised caffeine geometry (not in the reference tree; SURVEY 8d config 2)"}
    # README example (README.md:93-123)
    mols["LiH_readme"] = {"numbers": [3, 1], "positions": [[0, 0, 0], [0, 0, 1.5]], "charge": 0.0, "source": "README.md:93-123"}
    # EEQ known answer geometry (test/test_scf/test_guess.py:30-38)
    mols["CH_guess"] = {"numbers": [6, 1], "positions": [[0, 0, 0], [0, 0, 1.0]], "charge": 0.0, "source": "test/test_scf/test_guess.py:30-38"}
    (OUT / "molecules.json").write_text(json.dumps(mols))

    arrays = {}
    names = {"h2": "H2", "lih": "LiH", "h2o": "H2O", "ch4": "CH4", "sih4": "SiH4", "lys_xao": "LYS_xao", "mb16_43_01": "MB16_43_01",
             "c60": "C60", "vancoh2": "vancoh2"}
    ov = np.load(REF / "test/test_overlap/overlap.npz")
    h0 = np.load(REF / "test/test_hamiltonian/h0.npz")
    gs = np.load(REF / "test/test_scf/grad.npz")
    gn = np.load(REF / "test/test_hamiltonian/grad_no_overlap.npz")
    gh = np.load(REF / "test/test_hamiltonian/grad.npz")
    for k, name in names.items():
        if k in ov.files:
            arrays[f"overlap/{name}"] = ov[k]
        if k in h0.files:
            arrays[f"h0/{name}"] = h0[k]
        if k in gs.files:
            arrays[f"scf_grad/{name}"] = gs[k]
        if k in gh.files:
            arrays[f"h0_grad/{name}"] = gh[k]
        if f"{k}_dcn" in gn.files:
            arrays[f"dcn/{name}"] = gn[f"{k}_dcn"]
            arrays[f"dedcn/{name}"] = gn[f"{k}_dedcn"]
    # total GFN1-xTB gradients INCLUDING D3(BJ) (tblite, float32): test/test_singlepoint/refs/gfn1/grad.npz
    gt = np.load(REF / "test/test_singlepoint/refs/gfn1/grad.npz")
    tnames = {"h2": "H2", "h2o": "H2O", "no2": "NO2", "ch4": "CH4", "sih4": "SiH4", "lys_xao": "LYS_xao", "c60": "C60",
              "vancoh2": "vancoh2", "ad7en": "AD7en+"}
    for k, name in tnames.items():
        if k in gt.files:
            arrays[f"total_grad/{name}"] = gt[k]
    np.savez_compressed(OUT / "reference.npz", **arrays)

    def literals(path, key):
        txt = open(path).read()
        out = {}
        for m in re.finditer(r'"([A-Za-z0-9_+\-]+)": \{\s*"%s": torch\.tensor\(\s*([-0-9.e+]+)' % key, txt):
            out[m.group(1)] = float(m.group(2))
        return out

    num = r"[-+]?\d\.\d+e[-+]?\d+"

    def floats_after(txt, start, key, stop):
        blk = txt[txt.index(key, start):]
        return [float(x) for x in re.findall(num, blk[: blk.index(stop)])]

    # repulsion energies (fp64 literals) for the molecules whose geometries are in-tree: test_classical/test_repulsion/samples.py
    rtxt = open(REF / "test/test_classical/test_repulsion/samples.py").read()
    repulsion = {}
    for name in ("H2", "H2O", "SiH4", "LYS_xao", "MB16_43_01"):
        m = re.search(r'"%s": \{\s*"gfn1": torch\.tensor\(([-0-9.e+]+)\)' % re.escape(name), rtxt)
        repulsion[name] = float(m.group(1))
    # second / third order electrostatics at fixed charges: test_coulomb/samples.py
    ctxt = open(REF / "test/test_coulomb/samples.py").read()
    es2_shell = {}
    for name in ("SiH4", "LiH"):
        i = ctxt.index('"%s": {  # shell-resolved' % name)
        es2 = float(re.search(r'"es2": torch\.tensor\(\s*([-0-9.e+]+)', ctxt[i:]).group(1))
        es2_shell[name] = {"q_shell": floats_after(ctxt, i, '"q": torch.tensor(', "]"), "es2": es2,
                           "grad": floats_after(ctxt, i, '"grad": torch.tensor(', "dtype")}
    i = ctxt.index('"MB16_43_01": {')
    es3_atom = {"MB16_43_01": {"q_atom": floats_after(ctxt, i, '"q": torch.tensor(', "]"),
                               "es3": float(re.search(r'"es3": torch\.tensor\(\s*([-0-9.e+]+)', ctxt[i:]).group(1))}}

    # Fermi filling known answer: SiH4 orbital energies (test_wavefunction/samples.py:375-), occupations and electronic
    # free energy at 5000 K (test_wavefunction/test_filling.py:185-268)
    wtxt = open(REF / "test/test_wavefunction/samples.py").read()
    i = wtxt.index('"SiH4": {')
    eb = wtxt[wtxt.index('"emo": torch.tensor(', i):]
    emo = [float(x) for x in re.findall(r"[-+]?\d\.\d+(?:e[-+]?\d+)?", eb[: eb.index(")")])][:17]
    ftxt = open(REF / "test/test_wavefunction/test_filling.py").read()
    j = ftxt.index("5000.0: emo.new_tensor(\n            [")
    focc = [float(x) for x in re.findall(r"[-+]?\d\.\d+(?:e[-+]?\d+)?", ftxt[j + 10 : ftxt.index("]", j)])]
    fen = float(re.search(r"5000\.0: emo\.new_tensor\(([-0-9.e+]+)\)", ftxt).group(1))
    fermi = {"emo": emo, "nel": 8.0, "kelvin": 5000.0, "focc": focc, "fenergy": fen}

    # Wiberg bond orders and Mulliken atomic charges of the converged GFN1 density: test_wavefunction/samples.py
    wiberg = {}
    for name in ("H2", "LiH", "SiH4"):
        i = wtxt.index('"%s": {' % name)
        wb = wtxt[wtxt.index('"wiberg": torch.tensor(', i):]
        mc = wtxt[wtxt.index('"mulliken_charges": torch.tensor(', i):]
        fl = r"[-+]?\d\.\d+(?:e[-+]?\d+)?"
        wiberg[name] = {"wiberg": [float(x) for x in re.findall(fl, wb[: wb.index(")")])],
                        "mulliken_charges": [float(x) for x in re.findall(fl, mc[: mc.index(")")])]}

    energies = {
        "wiberg_gfn1": wiberg,
        "wiberg_source": "test/test_wavefunction/samples.py (H2, LiH, SiH4: flattened nat x nat Wiberg matrix, 5-digit Mulliken charges)",
        "fermi_sih4_5000K": fermi,
        "repulsion_gfn1": repulsion,
        "repulsion_source": "test/test_classical/test_repulsion/samples.py:50-420 (fp64 literals)",
        "es2_shell_gfn1": es2_shell,
        "es3_atom_gfn1": es3_atom,
        "coulomb_source": "test/test_coulomb/samples.py:60-170, 576-660 (charges, ES2 energy and nuclear gradient at fixed shell charges; ES3 with float32-precision atomic charges; the LiH energy literal is a 0.0 placeholder there)",
        "scf_gfn1_tblite": {k: v for k, v in literals(REF / "test/test_scf/samples.py", "egfn1").items() if k in mols},
        "total_gfn1_tblite": literals(REF / "test/test_singlepoint/samples.py", "egfn1"),
        "eeq_guess_CH": [-0.11593066900969, -0.03864355757833, -0.03864355757833, -0.03864355757833, 0.11593066900969, 0.11593066900969],
        "scf_gfn1_tblite_atoms": atom_refs(),
        "scf_gfn1_tblite_atoms_source": "test/test_scf/test_elements_gfn1.py:45-132 (tblite 0.2.1, neutral atoms Z=1..86, no repulsion/dispersion)",
        "note": "tblite values (fp64 literals of the reference tests); tblite uses 1 Eh = 27.21138505 eV",
    }
    (OUT / "energies.json").write_text(json.dumps(energies, indent=1))
    print({k: len(v["numbers"]) for k, v in mols.items()})
    print(sorted(arrays))
    print(energies["scf_gfn1_tblite"])


if __name__ == "__main__":
    main()
