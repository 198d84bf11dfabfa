#!/usr/bin/env python
"""Per-step wall-clock vs device time of the single-point step (developer tool)."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from dxtb_b200 import GFN1Calculator

dev = torch.device("cuda:0")
resp = "--no-response" not in sys.argv
wl = bench.Workload(2, 1, 1024)
numbers = torch.from_numpy(wl.numbers).to(dev)
chrg = torch.zeros(1024, dtype=torch.float64, device=dev)
calc = GFN1Calculator(numbers, device=dev, dtype=torch.float64, d3_reference=bench._d3_table(), opts={"grad_response": resp})
pos = [torch.from_numpy(wl.positions(s, np.arange(1024))).to(dev) for s in range(12)]
for s in range(12):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    p = pos[s].detach().requires_grad_(True)
    e = calc.get_energy(p, chrg)
    e1.record()
    (g,) = torch.autograd.grad(e.sum(), p)
    e2.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"step {s}: wall {1e3*(t1-t0):.1f} ms  fwd {e0.elapsed_time(e1):.1f}  bwd {e1.elapsed_time(e2):.1f}")
