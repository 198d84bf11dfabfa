#!/usr/bin/env python
"""A/B of kernel-variant choices on mid-size molecules (developer tool): SP/s of energy+forces for batches of one molecule."""
import os, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from dxtb_b200 import GFN1Calculator

dev = torch.device("cuda:0")
for name, nb in (("LYS_xao", 1024), ("capsaicin", 1024), ("C60", 592), ("vancoh2", 296)):
    z, base = bench.load_caffeine(name)
    numbers = torch.tensor(z)[None].expand(nb, -1).contiguous().to(dev)
    calc = GFN1Calculator(numbers, device=dev, dtype=torch.float64, opts={"exclude": ["disp"]})
    ts = []
    for s in range(3):
        p = torch.from_numpy(bench.conformers(base, nb, s)).to(dev).requires_grad_(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e = calc.get_energy(p)
        (g,) = torch.autograd.grad(e.sum(), p)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"{name:10s} nao {int(calc.desc.nao[0]):4d} nb {nb:5d} variants {calc._variants}: {nb/min(ts):9.1f} SP/s")
