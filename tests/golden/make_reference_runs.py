#!/usr/bin/env python
"""Run the reference's OWN code (dxtb v0.4.0, unmodified, baseline/_ref) and commit its outputs as fixtures.

    python oracle/build_ref.py && python tests/golden/make_reference_runs.py

dxtb is pure Python; its un-vendored utility dependencies are replaced by the stand-ins of oracle/shim (see the README
there for what that does and does not pin).  Everything else — index helper, STO-NG basis, McMurchie-Davidson overlap,
H0, repulsion, halogen bond, ES2/ES3, guess spreading, Fermi filling, the unrolled SCF loop with its Anderson mixer and
culling, energy assembly, autograd forces — is dxtb's code executed here, fp64 on CPU.

Writes tests/golden/reference_runs.npz:
  <case>/<mol>/energy            total energy (get_energy)
  <case>/<mol>/iterations        get_iterations (number of SCF map evaluations)
  <case>/<mol>/q_orb             get_charges (orbital-resolved Mulliken charges)
  <case>/<mol>/forces            get_forces: autograd through the unrolled SCF (the reference's force definition)
with cases
  default   dxtb defaults (EEQ guess, x_atol 1e-4 / 1e-5), exclude=["disp"]
  sad       the same with guess="sad" (no third-party EEQ model on the path)
  tight     x_atol = x_atol_max = f_atol = 1e-10, EEQ guess
  batch/*   padded batches: energy / iterations / forces per batch
  d3shim    default + D3(BJ) with the SYNTHETIC table (oracle.synthetic_d3_table; arithmetic of the tad-dftd3 stand-in)
"""
import json
import os
import sys
import tempfile
import time
import warnings
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from oracle.build_ref import reference_paths  # noqa: E402

sys.path[:0] = reference_paths()

import torch  # noqa: E402
from dxtb.calculators import GFN1Calculator  # noqa: E402

from halogen_mols import HALOGEN_MOLS, halogen_mol  # noqa: E402
from oracle import gfn1_oracle as O  # noqa: E402

DD = {"dtype": torch.float64, "device": torch.device("cpu")}
SINGLE = ["H", "H2", "LiH_readme", "H2O", "NO2", "CH4", "SiH4", "MB16_43_01", "caffeine", "nicotine", "AD7en+", "LYS_xao"]
OPTS = {
    "default": {"verbosity": 0, "exclude": ["disp"]},
    "sad": {"verbosity": 0, "exclude": ["disp"], "guess": "sad"},
    "tight": {"verbosity": 0, "exclude": ["disp"], "x_atol": 1e-10, "x_atol_max": 1e-10, "f_atol": 1e-10},
}


def geometries():
    mols = json.load(open(HERE / "molecules.json"))
    g = {n: (np.array(mols[n]["numbers"]), np.array(mols[n]["positions"]), float(mols[n]["charge"])) for n in SINGLE}
    for n in HALOGEN_MOLS:
        z, p = halogen_mol(n)
        g[n] = (z, p, 0.0)
    return g


def run(numbers, positions, chrg, opts):
    numbers = torch.as_tensor(numbers)
    pos = torch.as_tensor(positions, **DD).clone().requires_grad_(True)
    chrg_t = torch.as_tensor(chrg, **DD)
    calc = GFN1Calculator(numbers, opts=dict(opts), **DD)
    e = calc.get_energy(pos, chrg=chrg_t)
    (g,) = torch.autograd.grad(e.sum(), pos)
    it = calc.get_iterations(pos, chrg=chrg_t)
    q = calc.get_charges(pos, chrg=chrg_t)
    return dict(energy=e.detach().numpy(), forces=-g.numpy(), iterations=np.asarray(it), q_orb=q.detach().numpy())


def pad(geoms):
    nat = max(len(z) for z, _, _ in geoms)
    numbers = np.zeros((len(geoms), nat), dtype=np.int64)
    pos = np.zeros((len(geoms), nat, 3))
    for i, (z, p, _) in enumerate(geoms):
        numbers[i, : len(z)] = z
        pos[i, : len(z)] = p
    return numbers, pos, np.array([c for _, _, c in geoms])


def main():
    warnings.simplefilter("ignore")
    out = {}
    g = geometries()
    for case, opts in OPTS.items():
        for name, (z, p, c) in g.items():
            if case == "tight" and name in ("LYS_xao",):
                continue
            t0 = time.time()
            r = run(z, p, c, opts)
            for k, v in r.items():
                out[f"{case}/{name}/{k}"] = v
            print(f"{case:8s} {name:16s} E = {float(r['energy']):.12f}  iter = {int(r['iterations'])}  ({time.time() - t0:.1f} s)", flush=True)
    batches = {
        "mixed": ["LiH_readme", "H2O", "caffeine", "SiH4"],
        "halogen": ["CH3Br_NH3", "CH4", "CH2BrI_cluster"],
        "charged": ["AD7en+", "NO2", "H2"],
    }
    for bname, names in batches.items():
        numbers, pos, chrg = pad([g[n] for n in names])
        r = run(numbers, pos, chrg, OPTS["default"])
        out[f"batch/{bname}/names"] = np.array(names)
        for k, v in r.items():
            out[f"batch/{bname}/{k}"] = v
        print(f"batch    {bname:16s} E = {r['energy']}  iter = {r['iterations']}", flush=True)
    # D3(BJ) through the tad-dftd3 stand-in with the synthetic table (same table as bench.py)
    table = O.synthetic_d3_table()
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "d3_synth.npz")
        np.savez(path, **table)
        os.environ["TAD_DFTD3_SHIM_TABLE"] = path
        for name in ("H2O", "caffeine", "CH3Br_NH3"):
            z, p, c = g[name]
            r = run(z, p, c, {"verbosity": 0})
            for k, v in r.items():
                out[f"d3shim/{name}/{k}"] = v
            print(f"d3shim   {name:16s} E = {float(r['energy']):.12f}  iter = {int(r['iterations'])}", flush=True)
    np.savez_compressed(HERE / "reference_runs.npz", **out)
    print("wrote", HERE / "reference_runs.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
