#!/usr/bin/env python
"""Small driver for ncu: a few SCF launches on one wave of caffeine conformers."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from dxtb_b200 import GFN1Calculator  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 148
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
numbers_np, base = bench.load_caffeine()
numbers = torch.tensor(numbers_np)[None].expand(nb, -1).contiguous().to(dev)
chrg = torch.zeros(nb, dtype=torch.float64, device=dev)
calc = GFN1Calculator(numbers, opts={"exclude": ["disp"]}, device=dev, dtype=torch.float64)
for s in range(reps):
    p = torch.from_numpy(bench.conformers(base, nb, s)).to(dev).requires_grad_(True)
    e = calc.get_energy(p, chrg)
    (g,) = torch.autograd.grad(e.sum(), p)
torch.cuda.synchronize()
st = calc.cache["status"]
print("ok", float(e.sum()), "mean sweeps", float((st >> 8).float().mean()), "mean iters", float(calc.get_iterations().float().mean()))
