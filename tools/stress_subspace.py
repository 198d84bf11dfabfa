#!/usr/bin/env python
"""Stress test of the occupied-subspace path (developer tool, GPU box): every closed-shell fixture molecule with STRONGLY perturbed
geometries (sigma up to 0.15 bohr: smaller gaps, harder SCF), opts["scf_subspace"] True against False -- iteration counts, status
words, energies, forces."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from dxtb_b200 import GFN1Calculator  # noqa: E402

mols = json.load(open(ROOT / "tests/golden/molecules.json"))
dev = torch.device("cuda:0")
names = ["MB16_43_01", "caffeine", "nicotine", "LYS_xao", "AD7en+", "capsaicin", "C60"]
bad = 0
for sigma in (0.05, 0.10, 0.15):
    for name in names:
        m = mols[name]
        nb = 64 if len(m["numbers"]) < 50 else 16
        rng = np.random.default_rng(int(1000 * sigma) + len(name))
        base = np.array(m["positions"])
        pos = torch.from_numpy(base[None] + rng.normal(0.0, sigma, size=(nb,) + base.shape)).to(dev)
        numbers = torch.tensor(m["numbers"])[None].expand(nb, -1).contiguous().to(dev)
        chrg = torch.full((nb,), float(m["charge"]), dtype=torch.float64, device=dev)
        out = {}
        for sub in (True, False):
            calc = GFN1Calculator(numbers, opts={"exclude": ["disp"], "scf_subspace": sub, "maxiter": 100}, device=dev, dtype=torch.float64)
            p = pos.clone().requires_grad_(True)
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                e = calc.get_energy(p, chrg)
            (g,) = torch.autograd.grad(e.sum(), p)
            st = calc.cache["status"]
            out[sub] = (e.detach().cpu().numpy(), g.cpu().numpy(), calc.get_iterations().cpu().numpy(), (st & 255).cpu().numpy(), (st >> 8).cpu().numpy())
        a, b = out[True], out[False]
        conv = (b[3] == 0)
        dit = int(np.abs(a[2] - b[2]).max())
        de = float(np.abs(a[0] - b[0])[conv].max()) if conv.any() else 0.0
        dg = float(np.abs(a[1] - b[1])[conv].max()) if conv.any() else 0.0
        same_status = bool((a[3] == b[3]).all())
        flag = "" if (dit == 0 and de < 1e-9 and dg < 1e-7 and same_status) else "   <-- CHECK"
        bad += bool(flag)
        print(f"sigma {sigma:.2f} {name:11s} nb {nb:3d}: iterations max diff {dit}  dE {de:.1e}  dF {dg:.1e}  status equal {same_status} "
              f"(not converged: {int((~conv).sum())})  sweeps {a[4].mean():.1f} vs {b[4].mean():.1f}{flag}", flush=True)
print("mismatches:", bad)
