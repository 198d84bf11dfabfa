#!/usr/bin/env python
"""Developer check of the large-system SCF path (GPU box): vancoh2 through both paths, then sh3 against the oracle fixture."""
import json, os, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from dxtb_b200 import GFN1Calculator
mols = json.load(open(ROOT / "tests/golden/molecules.json"))
dev = torch.device("cuda:0")

def run(name, large_min, grad=True):
    m = mols[name]
    numbers = torch.tensor(m["numbers"])[None].to(dev)
    pos = torch.tensor(m["positions"], dtype=torch.float64, device=dev)[None]
    chrg = torch.tensor([float(m["charge"])], dtype=torch.float64, device=dev)
    os.environ["DXTB_B200_LARGE_MIN_NAO"] = str(large_min)
    calc = GFN1Calculator(numbers, opts={"exclude": ["disp"]}, device=dev, dtype=torch.float64)
    out = None
    for rep in range(2):
        p = pos.clone().requires_grad_(grad)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e = calc.get_energy(p, chrg)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        g = torch.autograd.grad(e.sum(), p)[0] if grad else None
        torch.cuda.synchronize(); t2 = time.perf_counter()
        st = int(calc.cache["status"][0])
        print(f"{name} variants={calc._variants} E={float(e[0]):.12f} iters={int(calc.get_iterations()[0])} status={st & 255} sweeps={st >> 8} "
              f"fwd {1e3*(t1-t0):.1f} ms bwd {1e3*(t2-t1):.1f} ms", flush=True)
        out = (float(e[0]), int(calc.get_iterations()[0]), None if g is None else g[0].cpu().numpy(), calc.get_atomic_charges()[0].cpu().numpy())
    return out

which = sys.argv[1:] or ["vancoh2", "sh3"]
if "vancoh2" in which:
    a = run("vancoh2", 10**6)
    b = run("vancoh2", 500)
    print("vancoh2 large vs one-CTA: dE %.3e  iters %d/%d  dF %.3e dq %.3e" % (abs(a[0] - b[0]), a[1], b[1], np.abs(a[2] - b[2]).max(), np.abs(a[3] - b[3]).max()))
if "sh3" in which:
    r = np.load(ROOT / "tests/golden/sh3_oracle.npz")
    a = run("ex_sh3", 10**6)
    print("sh3 vs oracle: dE %.3e  iters %d/%d  dF %.3e dq %.3e" % (abs(a[0] - float(r["energy"])), a[1], int(r["iterations"]),
          np.abs(a[2] - r["gradient"]).max(), np.abs(a[3] - r["q_atom"]).max()))
