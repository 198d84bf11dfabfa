"""Occupied-subspace solve of the intermediate SCF iterations (dxtb_b200/csrc/xtb_scf_subspace.cuh) against the path that
diagonalises in every iteration (opts["scf_subspace"] = False, what the reference does: scf/unrolling/base.py:141-175).

The trajectory may differ by the residual of the intermediate solves only (1e-10), the final solve is the same full
eigendecomposition: energies 1e-10 Eh, charges / forces 1e-8, identical iteration counts; and the path must really be taken
(fewer Jacobi sweeps) where it applies and refused where occupations are fractional."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _conformers(base, nb, seed):
    rng = np.random.default_rng(seed)
    return base[None] + rng.normal(0.0, 0.05, size=(nb,) + base.shape)


def _run(numbers, pos, chrg, opts, dev, override=None):
    from dxtb_b200 import GFN1Calculator

    calc = GFN1Calculator(numbers, opts=opts, device=dev, dtype=torch.float64)
    if override is not None:
        calc._use_smem_override = override
    p = pos.clone().requires_grad_(True)
    e = calc.get_energy(p, chrg)
    (g,) = torch.autograd.grad(e.sum(), p)
    st = calc.cache["status"]
    return dict(e=e.detach().cpu().numpy(), g=g.cpu().numpy(), q=calc.get_charges().cpu().numpy(), it=calc.get_iterations().cpu().numpy(),
                sweeps=(st >> 8).cpu().numpy(), status=(st & 255).cpu().numpy())


@pytest.mark.parametrize("name,nb,override", [("caffeine", 64, None), ("caffeine", 8, 0), ("caffeine", 8, 2), ("LYS_xao", 8, None),
                                              ("MB16_43_01", 8, None), ("nicotine", 8, None), ("capsaicin", 8, None), ("AD7en+", 4, None)])
def test_subspace_path_reproduces_full_diagonalisation(mols, name, nb, override):
    dev = _dev()
    m = mols[name]
    numbers = torch.tensor(m["numbers"])[None].expand(nb, -1).contiguous().to(dev)
    chrg = torch.full((nb,), float(m["charge"]), dtype=torch.float64, device=dev)
    pos = torch.from_numpy(_conformers(np.array(m["positions"]), nb, 11)).to(dev)
    a = _run(numbers, pos, chrg, {"exclude": ["disp"], "scf_subspace": True}, dev, override)
    b = _run(numbers, pos, chrg, {"exclude": ["disp"], "scf_subspace": False}, dev, override)
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    assert np.array_equal(a["it"], b["it"])
    assert np.abs(a["e"] - b["e"]).max() < 1e-10
    assert np.abs(a["q"] - b["q"]).max() < 1e-8
    assert np.abs(a["g"] - b["g"]).max() < 1e-8
    # the path is taken: most intermediate solves need no Jacobi sweep at all (AD7en+: the certified gap of the cation is
    # below 50 kT in part of its iterations, which then diagonalise)
    assert a["sweeps"].mean() < (0.9 if name == "AD7en+" else 0.4) * b["sweeps"].mean(), (a["sweeps"].mean(), b["sweeps"].mean())


def test_subspace_path_refused_for_fractional_occupations(mols):
    """Open shells, hot electrons (gap < 50 kT) and molecules below 33 AOs must take the full eigendecomposition:
    bit-identical results."""
    dev = _dev()
    for name, opts in (("NO2", {}), ("caffeine", {"fermi_etemp": 5000.0}), ("H", {}), ("H2O", {}), ("LiH", {})):
        m = mols[name]
        numbers = torch.tensor(m["numbers"])[None].to(dev)
        chrg = torch.full((1,), float(m["charge"]), dtype=torch.float64, device=dev)
        pos = torch.tensor(m["positions"], dtype=torch.float64)[None].to(dev)
        a = _run(numbers, pos, chrg, {"exclude": ["disp"], "scf_subspace": True, **opts}, dev)
        b = _run(numbers, pos, chrg, {"exclude": ["disp"], "scf_subspace": False, **opts}, dev)
        assert np.array_equal(a["e"], b["e"]) and np.array_equal(a["g"], b["g"]) and np.array_equal(a["it"], b["it"])
        assert np.array_equal(a["sweeps"], b["sweeps"])


def test_subspace_path_on_the_large_system_path(mols, monkeypatch):
    """Grid-wide version (xtb_scf_large.cu): caffeine and LYS_xao forced onto the large-system path, subspace solve on against
    off -- same iteration counts, energies 1e-10, forces 1e-8, and fewer sweeps."""
    dev = _dev()
    monkeypatch.setenv("DXTB_B200_LARGE_MIN_NAO", "1")
    for name in ("caffeine", "LYS_xao"):
        m = mols[name]
        numbers = torch.tensor(m["numbers"])[None].expand(2, -1).contiguous().to(dev)
        chrg = torch.full((2,), float(m["charge"]), dtype=torch.float64, device=dev)
        pos = torch.from_numpy(_conformers(np.array(m["positions"]), 2, 5)).to(dev)
        a = _run(numbers, pos, chrg, {"exclude": ["disp"], "scf_subspace": True}, dev)
        b = _run(numbers, pos, chrg, {"exclude": ["disp"], "scf_subspace": False}, dev)
        assert (a["status"] == 0).all() and (b["status"] == 0).all()
        assert np.array_equal(a["it"], b["it"])
        assert np.abs(a["e"] - b["e"]).max() < 1e-10
        assert np.abs(a["g"] - b["g"]).max() < 1e-8
        assert a["sweeps"].mean() < 0.6 * b["sweeps"].mean(), (a["sweeps"].mean(), b["sweeps"].mean())


def test_persistent_launch_with_eigenvector_warm_start(mols, monkeypatch):
    """More molecules than SMs in a uniform bucket: persistent CTAs start every molecule after their first from the previous
    molecule's eigenvectors (xtb_scf_opts.persistent).  Against one CTA per molecule with the Cholesky start
    basis: identical iteration counts, energies 1e-11 Eh, forces 1e-9, fewer sweeps; and the launch is deterministic."""
    dev = _dev()
    nb = 2 * torch.cuda.get_device_properties(dev).multi_processor_count + 17
    m = mols["caffeine"]
    numbers = torch.tensor(m["numbers"])[None].expand(nb, -1).contiguous().to(dev)
    chrg = torch.zeros(nb, dtype=torch.float64, device=dev)
    pos = torch.from_numpy(_conformers(np.array(m["positions"]), nb, 3)).to(dev)
    a = _run(numbers, pos, chrg, {"exclude": ["disp"]}, dev)
    a2 = _run(numbers, pos, chrg, {"exclude": ["disp"]}, dev)
    monkeypatch.setenv("DXTB_B200_WARM_START", "0")
    b = _run(numbers, pos, chrg, {"exclude": ["disp"]}, dev)
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    assert np.array_equal(a["it"], b["it"])
    assert np.abs(a["e"] - b["e"]).max() < 1e-11
    assert np.abs(a["g"] - b["g"]).max() < 1e-9
    assert a["sweeps"].mean() < b["sweeps"].mean() - 0.5
    assert np.array_equal(a["e"], a2["e"]) and np.array_equal(a["g"], a2["g"])  # fixed-stride walk: same bits every time
