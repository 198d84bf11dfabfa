#!/usr/bin/env python
"""Time one 1/8 shard of the config-5 ragged batch on one GPU with different treatments of its big molecules."""
import os, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from dxtb_b200 import GFN1Calculator

dev = torch.device("cuda:0")
wl = bench.Workload(5, 8, 4096)
idx = wl.shard(0)
numbers = torch.from_numpy(wl.numbers[idx]).to(dev)
chrg = torch.from_numpy(wl.chrg[idx]).to(dev)
for maxcount, conc in ((0, 1), (64, 1), (64, 4), (64, 6), (64, 8)):
    os.environ["DXTB_B200_LARGE_MAX_COUNT"] = str(maxcount)
    os.environ["DXTB_B200_LARGE_CONCURRENCY"] = str(conc)
    calc = GFN1Calculator(numbers, device=dev, dtype=torch.float64, opts={"exclude": ["disp"]})
    ts = []
    for s in range(3):
        p = torch.from_numpy(wl.positions(s, idx)).to(dev).requires_grad_(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e = calc.get_energy(p, chrg)
        (g,) = torch.autograd.grad(e.sum(), p)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"large_max_count {maxcount:3d} concurrency {conc}: variants {calc._variants}  step {1e3*min(ts):.0f} ms  E0 {float(e[0]):.10f} sumE {float(e.sum()):.8f}")
