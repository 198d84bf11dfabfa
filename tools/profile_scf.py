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

# ---- step breakdown (wall clock, synchronised) ----
import time
pp = torch.from_numpy(bench.conformers(base, nb, 99)).to(dev)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    p = pp.detach().requires_grad_(True)
    e = calc.get_energy(p, chrg)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    (g,) = torch.autograd.grad(e.sum(), p)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"forward {1e3*(t1-t0):.2f} ms  backward {1e3*(t2-t1):.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    p = pp.detach().requires_grad_(True)
    e = calc.get_energy(p, chrg)
    (g,) = torch.autograd.grad(e.sum(), p)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
