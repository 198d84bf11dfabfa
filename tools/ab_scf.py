#!/usr/bin/env python
"""A/B timing of the SCF kernel alone for library builds (developer tool): DXTB_B200_LIB=path python tools/ab_scf.py"""
import json, os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from dxtb_b200 import GFN1Calculator
mols = json.load(open(ROOT / "tests/golden/molecules.json"))
dev = torch.device("cuda:0")
out = []
for name, nb in [("caffeine", 1024), ("LYS_xao", 592), ("capsaicin", 592), ("C60", 296)]:
    m = mols[name]
    numbers = torch.tensor(m["numbers"])[None].expand(nb, -1).contiguous().to(dev)
    chrg = torch.full((nb,), float(m["charge"]), dtype=torch.float64, device=dev)
    calc = GFN1Calculator(numbers, opts={"exclude": ["disp"]}, device=dev, dtype=torch.float64)
    ts = []
    for rep in range(5):
        p = torch.from_numpy(bench.conformers(np.array(m["positions"]), nb, rep)).to(dev)
        calc.scf_events = []
        calc.get_energy(p, chrg)
        torch.cuda.synchronize()
        ts.append(calc.scf_events[0][0].elapsed_time(calc.scf_events[0][1]))
    st = calc.cache["status"]
    out.append(f"{name}{calc._variants} {min(ts[1:]):.2f} (sw {float((st >> 8).float().mean()):.1f} it {int(calc.get_iterations().sum())})")
print(os.environ.get("DXTB_B200_LIB", "default"), " | ".join(out))
