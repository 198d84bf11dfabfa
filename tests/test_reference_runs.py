"""Parity with the reference's OWN code: dxtb v0.4.0 executed unmodified (tests/golden/make_reference_runs.py, fixtures in
tests/golden/reference_runs.npz; dxtb's un-vendored utility dependencies replaced by oracle/shim).

* CPU (`-m "not gpu"`): the NumPy oracle reproduces dxtb's energies, charges and SCF iteration counts, and its analytic
  gradient equals dxtb's autograd forces at tight SCF thresholds; the gap at dxtb's default thresholds (analytic
  converged-SCF gradient vs autograd through the truncated SCF) is bounded and printed.
* GPU (`-m gpu`): the CUDA path against the same fixtures with the north-star tolerances.
* live (CPU, skipped when baseline/_ref is absent): dxtb is re-run here on two molecules and must reproduce its fixtures.
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from halogen_mols import HALOGEN_MOLS, halogen_mol
from oracle import gfn1_oracle as O

ROOT = Path(__file__).resolve().parent.parent
E_TOL, Q_TOL, F_TOL = 1e-9, 1e-7, 1e-7
CASES = {"default": {}, "sad": {"guess": "sad"}, "tight": {"x_atol": 1e-10, "x_atol_max": 1e-10}}
# Forces of the reference = autograd through the unrolled SCF.  At dxtb's DEFAULT thresholds (x_atol 1e-4 / 1e-5) the plain
# converged-SCF analytic gradient differs from them by up to 8.7e-6 Eh/bohr (first order in the SCF residual); with the
# first-order response of the residual (oracle _scf_response / CUDA scf_response) the gap is <= 3.1e-8 on every closed-shell
# fixture molecule.  Exception: the NO2 radical, where the derivative of the reference's own SCF trajectory is not converged
# in the symmetry-breaking charge mode (dv_K/dR differs from the converged response by 0.2, finite-difference checked); its
# autograd forces differ from the finite difference of its own energy by 3.6e-7 there.
F_GAP_DEFAULT = 1e-7
F_GAP_UNCONVERGED_TRAJECTORY = {"NO2": 5e-6}
F_GAP_NO_RESPONSE = 2e-5


@pytest.fixture(scope="module")
def runs():
    return np.load(ROOT / "tests" / "golden" / "reference_runs.npz")


def _names(runs, case):
    return sorted({k.split("/")[1] for k in runs.files if k.startswith(case + "/")})


def _geom(mols, n):
    if n in HALOGEN_MOLS:
        z, p = halogen_mol(n)
        return z, p, 0.0
    m = mols[n]
    return np.array(m["numbers"]), np.array(m["positions"]), float(m["charge"])


ALL = ["H", "H2", "LiH_readme", "H2O", "NO2", "CH4", "SiH4", "MB16_43_01", "caffeine", "nicotine", "AD7en+", "LYS_xao",
       "CH3Br_NH3", "CH3I_OCH2", "Br2_NH3", "CH2BrI_cluster"]


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("name", ALL)
def test_oracle_reproduces_dxtb(mols, runs, case, name):
    if f"{case}/{name}/energy" not in runs.files:
        pytest.skip("not in the fixture set")
    z, p, c = _geom(mols, name)
    r = O.singlepoint(z, p, c, opts=dict(exclude=("disp",), **CASES[case]), grad=True)
    assert abs(r.energy - float(runs[f"{case}/{name}/energy"])) < 1e-12
    assert np.abs(r.q_orb - runs[f"{case}/{name}/q_orb"][: len(r.q_orb)]).max() < 1e-9
    gap = np.abs(-r.gradient - runs[f"{case}/{name}/forces"]).max()
    if case == "tight":
        # AD7en+ stops one iteration apart at 1e-10 (the stop test sits in the eigensolver's rounding noise there)
        assert abs(r.iterations - int(runs[f"{case}/{name}/iterations"])) <= (1 if name == "AD7en+" else 0)
        assert gap < F_TOL
    else:
        assert r.iterations == int(runs[f"{case}/{name}/iterations"])
        assert gap < F_GAP_UNCONVERGED_TRAJECTORY.get(name, F_GAP_DEFAULT)
        r0 = O.singlepoint(z, p, c, opts=dict(exclude=("disp",), grad_response=False, **CASES[case]), grad=True)
        gap0 = np.abs(-r0.gradient - runs[f"{case}/{name}/forces"]).max()
        assert gap0 < F_GAP_NO_RESPONSE
        print(f"{case}/{name}: force gap to dxtb autograd at default thresholds {gap0:.2e} (converged-SCF formula) -> {gap:.2e} Eh/bohr")


def test_oracle_d3_arithmetic_vs_dxtb_with_shim_table(mols, runs):
    """dxtb's DispersionD3 wrapper (dispersion/d3.py:93-211) driving the tad-dftd3 stand-in with the synthetic table: the
    oracle's D3 restatement gives the same total energies and iteration counts (arithmetic only; the table is made up)."""
    table = O.synthetic_d3_table()
    for name in ("H2O", "caffeine", "CH3Br_NH3"):
        z, p, c = _geom(mols, name)
        r = O.singlepoint(z, p, c, d3_table=table, grad=True)
        assert abs(r.energy - float(runs[f"d3shim/{name}/energy"])) < 1e-12
        assert r.iterations == int(runs[f"d3shim/{name}/iterations"])
        assert np.abs(-r.gradient - runs[f"d3shim/{name}/forces"]).max() < F_GAP_DEFAULT


def test_reference_batch_semantics(mols, runs):
    """Padded batches in dxtb: per-system energies equal the single-molecule runs (culling), get_iterations = batch max."""
    for b in ("mixed", "halogen", "charged"):
        names = [str(n) for n in runs[f"batch/{b}/names"]]
        e = runs[f"batch/{b}/energy"]
        its = []
        for i, n in enumerate(names):
            assert abs(e[i] - float(runs[f"default/{n}/energy"])) < 1e-11
            its.append(int(runs[f"default/{n}/iterations"]))
        assert int(runs[f"batch/{b}/iterations"]) == max(its)


def test_live_reference_reproduces_fixtures(mols, runs):
    from oracle.build_ref import available, reference_paths

    if not available():
        pytest.skip("baseline/_ref not installed (python oracle/build_ref.py)")
    import subprocess

    code = f"""
import sys, json, warnings
sys.path[:0] = {reference_paths()!r}
warnings.simplefilter('ignore')
import torch
from dxtb.calculators import GFN1Calculator
mols = json.load(open({str(ROOT / 'tests' / 'golden' / 'molecules.json')!r}))
out = {{}}
for n in ('LiH_readme', 'H2O'):
    m = mols[n]
    calc = GFN1Calculator(torch.tensor(m['numbers']), opts={{'verbosity': 0, 'exclude': ['disp']}}, dtype=torch.float64)
    pos = torch.tensor(m['positions'], dtype=torch.float64, requires_grad=True)
    e = calc.get_energy(pos)
    (g,) = torch.autograd.grad(e, pos)
    out[n] = [e.item(), int(calc.get_iterations(pos)), (-g).tolist()]
print('RESULT' + json.dumps(out))
"""
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    out = json.loads(r.stdout.split("RESULT")[-1])
    for n, (e, it, f) in out.items():
        assert abs(e - float(runs[f"default/{n}/energy"])) < 1e-12
        assert it == int(runs[f"default/{n}/iterations"])
        assert np.abs(np.array(f) - runs[f"default/{n}/forces"]).max() < 1e-10


# ---------------------------------------------------------------------------------------------------------
# CUDA path vs dxtb
# ---------------------------------------------------------------------------------------------------------
def _pad(geoms, dev):
    nat = max(len(z) for z, _, _ in geoms)
    numbers = torch.zeros((len(geoms), nat), dtype=torch.long)
    pos = torch.zeros((len(geoms), nat, 3), dtype=torch.float64)
    for i, (z, p, _) in enumerate(geoms):
        numbers[i, : len(z)] = torch.as_tensor(z)
        pos[i, : len(z)] = torch.as_tensor(p)
    chrg = torch.tensor([c for _, _, c in geoms], dtype=torch.float64)
    return numbers.to(dev), pos.to(dev), chrg.to(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("case", list(CASES))
def test_cuda_path_reproduces_dxtb(mols, runs, case):
    """One padded batch of every fixture molecule: energies 1e-9 Eh, charges 1e-7 e, identical iteration counts; forces
    within 1e-7 Eh/bohr of dxtb's autograd forces at tight thresholds (and within the documented gap at the defaults)."""
    from dxtb_b200 import GFN1Calculator

    dev = torch.device("cuda:0")
    names = [n for n in ALL if f"{case}/{n}/energy" in runs.files]
    geoms = [_geom(mols, n) for n in names]
    numbers, pos, chrg = _pad(geoms, dev)
    opts = {"exclude": ["disp"], **CASES[case]}
    calc = GFN1Calculator(numbers, opts=opts, device=dev, dtype=torch.float64)
    p = pos.clone().requires_grad_(True)
    e = calc.get_energy(p, chrg)
    (g,) = torch.autograd.grad(e.sum(), p)
    q = calc.get_charges()
    it = calc.get_iterations()
    for i, n in enumerate(names):
        nat = len(geoms[i][0])
        ref_q = runs[f"{case}/{n}/q_orb"]
        assert abs(float(e[i]) - float(runs[f"{case}/{n}/energy"])) < E_TOL, n
        assert np.abs(q[i, : len(ref_q)].cpu().numpy() - ref_q).max() < Q_TOL, n
        gap = np.abs(-g[i, :nat].cpu().numpy() - runs[f"{case}/{n}/forces"]).max()
        if case == "tight":
            assert abs(int(it[i]) - int(runs[f"{case}/{n}/iterations"])) <= (1 if n == "AD7en+" else 0), n
            assert gap < F_TOL, n
        else:
            assert int(it[i]) == int(runs[f"{case}/{n}/iterations"]), n
            assert gap < F_GAP_UNCONVERGED_TRAJECTORY.get(n, F_GAP_DEFAULT), (n, gap)


@pytest.mark.gpu
def test_cuda_d3_vs_dxtb_with_shim_table(mols, runs):
    from dxtb_b200 import GFN1Calculator

    dev = torch.device("cuda:0")
    names = ["H2O", "caffeine", "CH3Br_NH3"]
    geoms = [_geom(mols, n) for n in names]
    numbers, pos, chrg = _pad(geoms, dev)
    calc = GFN1Calculator(numbers, device=dev, dtype=torch.float64, d3_reference=O.synthetic_d3_table())
    e = calc.get_energy(pos, chrg)
    for i, n in enumerate(names):
        assert abs(float(e[i]) - float(runs[f"d3shim/{n}/energy"])) < E_TOL
        assert int(calc.get_iterations()[i]) == int(runs[f"d3shim/{n}/iterations"])
