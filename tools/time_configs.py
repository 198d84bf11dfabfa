#!/usr/bin/env python
"""Timing of the other BASELINE configs (developer tool, GPU box)."""
import json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from dxtb_b200 import GFN1Calculator
mols = json.load(open(ROOT / "tests/golden/molecules.json"))
dev = torch.device("cuda:0")
def run(name, nb, forces=True):
    m = mols[name]
    base = np.array(m["positions"])
    numbers = torch.tensor(m["numbers"])[None].expand(nb, -1).contiguous().to(dev)
    chrg = torch.full((nb,), float(m["charge"]), dtype=torch.float64, device=dev)
    calc = GFN1Calculator(numbers, opts={"exclude": ["disp"]}, device=dev, dtype=torch.float64)
    for rep in range(3):
        p = torch.from_numpy(bench.conformers(base, nb, rep)).to(dev).requires_grad_(forces)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e = calc.get_energy(p, chrg)
        if forces:
            (g,) = torch.autograd.grad(e.sum(), p)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = calc.cache["status"]
    print(f"{name:10s} nb={nb:5d} nao={int(calc.desc.nao[0]):4d} smem={calc._variants} {dt*1e3:9.1f} ms  {nb/dt:9.1f} SP/s  iters {float(calc.get_iterations().float().mean()):.1f} sweeps {float((st>>8).float().mean()):.1f}")
for name, nb in [("caffeine", 1024), ("LYS_xao", 592), ("capsaicin", 592), ("C60", 296), ("vancoh2", 148)]:
    run(name, nb)


def run_mixed():
    """config-5-like ragged batch: nat 2..176 in one calculator (size buckets -> kernel variants)."""
    names = ["H2O", "LiH", "nicotine", "capsaicin", "caffeine", "LYS_xao", "C60", "vancoh2"]
    counts = [256, 256, 128, 128, 128, 64, 32, 16]
    rng = np.random.default_rng(2)
    nmax = max(len(mols[n]["numbers"]) for n in names)
    nb = sum(counts)
    numbers = torch.zeros((nb, nmax), dtype=torch.int64)
    pos = torch.zeros((nb, nmax, 3), dtype=torch.float64)
    chrg = torch.zeros(nb, dtype=torch.float64)
    i = 0
    for n, c in zip(names, counts):
        m = mols[n]
        k = len(m["numbers"])
        for _ in range(c):
            numbers[i, :k] = torch.tensor(m["numbers"])
            pos[i, :k] = torch.from_numpy(np.array(m["positions"]) + rng.normal(0.0, 0.02, (k, 3)))
            chrg[i] = float(m["charge"])
            i += 1
    perm = torch.from_numpy(rng.permutation(nb))
    numbers, pos, chrg = numbers[perm].to(dev), pos[perm].to(dev), chrg[perm].to(dev)
    calc = GFN1Calculator(numbers, opts={"exclude": ["disp"]}, device=dev, dtype=torch.float64)
    for rep in range(2):
        p = pos.clone().requires_grad_(True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e = calc.get_energy(p, chrg)
        (g,) = torch.autograd.grad(e.sum(), p)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"mixed      nb={nb:5d} nat 2..176 variants={calc._variants} buckets={[(b['use_smem'], b['len']) for b in calc._buckets]} {dt*1e3:9.1f} ms  {nb/dt:9.1f} SP/s")


run("ex_sh3", 1)
run_mixed()
