#!/usr/bin/env python
"""Tiny workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel variant once, energy + forces.
    compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import json, os, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from dxtb_b200 import GFN1Calculator
mols = json.load(open(ROOT / "tests/golden/molecules.json"))
dev = torch.device("cuda:0")

def run(names, override=None, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    nat = max(len(mols[n]["numbers"]) for n in names)
    numbers = torch.zeros((len(names), nat), dtype=torch.long)
    pos = torch.zeros((len(names), nat, 3), dtype=torch.float64)
    chrg = torch.zeros(len(names), dtype=torch.float64)
    for i, n in enumerate(names):
        m = mols[n]
        k = len(m["numbers"])
        numbers[i, :k] = torch.tensor(m["numbers"]); pos[i, :k] = torch.tensor(m["positions"], dtype=torch.float64); chrg[i] = m["charge"]
    calc = GFN1Calculator(numbers.to(dev), opts={"exclude": ["disp"], "maxiter": int(os.environ.get("SAN_MAXITER", "3"))}, device=dev, dtype=torch.float64)
    calc._use_smem_override = override
    p = pos.to(dev).requires_grad_(True)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e = calc.get_energy(p, chrg.to(dev))
    (g,) = torch.autograd.grad(e.sum(), p)
    torch.cuda.synchronize()
    print(names, calc._variants, [round(float(x), 6) for x in e])

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "smem"):
    run(["H2O", "caffeine"])                                        # variant 1
if which in ("all", "hybrid"):
    run(["LYS_xao"])                                                # variant 2 (V pass pooled)
if which in ("all", "global"):
    run(["nicotine", "H2O"], override=0)                            # variant 0 (forced)
if which in ("all", "large"):
    run(["H2O", "caffeine"], DXTB_B200_LARGE_MIN_NAO=1)            # variant 3 (two molecules in flight on host threads)
if which in ("all", "persistent"):
    # more equally sized molecules than SMs: persistent CTAs + eigenvector warm start (round 2b)
    import numpy as np
    nsm = torch.cuda.get_device_properties(dev).multi_processor_count
    m = mols["nicotine"]
    nb = nsm + 12
    rng = np.random.default_rng(0)
    pos = torch.from_numpy(np.array(m["positions"])[None] + rng.normal(0.0, 0.05, size=(nb, len(m["numbers"]), 3))).to(dev).requires_grad_(True)
    calc = GFN1Calculator(torch.tensor(m["numbers"])[None].expand(nb, -1).contiguous().to(dev),
                          opts={"exclude": ["disp"], "maxiter": int(os.environ.get("SAN_MAXITER", "3"))}, device=dev, dtype=torch.float64)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        e = calc.get_energy(p := pos, torch.zeros(nb, dtype=torch.float64, device=dev))
    (g,) = torch.autograd.grad(e.sum(), p)
    torch.cuda.synchronize()
    print("nicotine x", nb, calc._variants, round(float(e.sum()), 6), "sweeps", float((calc.cache["status"] >> 8).float().mean()))
if which in ("all", "halogen"):
    # halogen-bond energy + gradient, SCF response on a converged-enough state (round 2)
    sys.path.insert(0, str(ROOT / "tests"))
    from halogen_mols import halogen_mol
    z, xyz = halogen_mol("CH3Br_NH3")
    os.environ["DXTB_B200_LARGE_MIN_NAO"] = "1000000"
    calc = GFN1Calculator(torch.tensor(z, device=dev), opts={"exclude": ["disp"]}, device=dev, dtype=torch.float64)
    p = torch.tensor(xyz, device=dev, requires_grad=True)
    e = calc.get_energy(p)
    (g,) = torch.autograd.grad(e, p)
    torch.cuda.synchronize()
    print("CH3Br_NH3", calc._variants, float(e), float(g.abs().max()))
