#!/usr/bin/env python
"""A/B of the occupied-subspace solve of the intermediate SCF iterations (developer tool, GPU box): for several molecules and
batch sizes, energies / forces / charges / iteration counts with opts["scf_subspace"] True against False, SCF kernel times and
mean Jacobi sweeps.

    python tools/ab_subspace.py [quick]
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from dxtb_b200 import GFN1Calculator  # noqa: E402

mols = json.load(open(ROOT / "tests/golden/molecules.json"))
dev = torch.device("cuda:0")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
cases = [("LiH", 4), ("H2O", 8), ("CH4", 8), ("SiH4", 8), ("MB16_43_01", 32), ("caffeine", 148), ("caffeine", 1024), ("LYS_xao", 148),
         ("nicotine", 148), ("capsaicin", 296), ("AD7en+", 148), ("NO2", 8), ("C60", 148)]
if not quick:
    cases += [("vancoh2", 16)]
worst = {"dE": 0.0, "dF": 0.0, "dq": 0.0, "dit": 0}
for name, nb in cases:
    m = mols[name]
    numbers = torch.tensor(m["numbers"])[None].expand(nb, -1).contiguous().to(dev)
    chrg = torch.full((nb,), float(m["charge"]), dtype=torch.float64, device=dev)
    res = {}
    for sub in (True, False):
        calc = GFN1Calculator(numbers, opts={"exclude": ["disp"], "scf_subspace": sub}, device=dev, dtype=torch.float64)
        ts = []
        for rep in range(3):
            p = torch.from_numpy(bench.conformers(np.array(m["positions"]), nb, rep)).to(dev).requires_grad_(True)
            calc.scf_events = []
            e = calc.get_energy(p, chrg)
            (g,) = torch.autograd.grad(e.sum(), p)
            torch.cuda.synchronize()
            ts.append(calc.scf_events[0][0].elapsed_time(calc.scf_events[0][1]))
        st = calc.cache["status"]
        res[sub] = dict(e=e.detach().cpu().numpy(), g=g.cpu().numpy(), q=calc.get_charges().cpu().numpy(), it=calc.get_iterations().cpu().numpy(),
                        t=min(ts[1:]), sw=float((st >> 8).float().mean()), bad=int((st & 255).ne(0).sum()), var=calc._variants)
    a, b = res[True], res[False]
    dE = float(np.abs(a["e"] - b["e"]).max())
    dF = float(np.abs(a["g"] - b["g"]).max())
    dq = float(np.abs(a["q"] - b["q"]).max())
    dit = int(np.abs(a["it"] - b["it"]).max())
    worst = {"dE": max(worst["dE"], dE), "dF": max(worst["dF"], dF), "dq": max(worst["dq"], dq), "dit": max(worst["dit"], dit)}
    print(f"{name:11s} nb={nb:5d} var={a['var']} scf {a['t']:8.2f} ms vs {b['t']:8.2f} ms ({b['t'] / a['t']:.2f}x)  sweeps {a['sw']:5.1f} vs {b['sw']:5.1f}  "
          f"iters {a['it'].mean():.2f} vs {b['it'].mean():.2f} (max diff {dit})  dE {dE:.1e} dF {dF:.1e} dq {dq:.1e}  status!=0: {a['bad']} / {b['bad']}", flush=True)
print("worst", worst)
