#!/usr/bin/env python
"""Developer check (GPU box): stage-by-stage comparison of the CUDA path with the NumPy oracle."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from dxtb_b200 import GFN1Calculator  # noqa: E402
from oracle import gfn1_oracle as O  # noqa: E402

mols = json.load(open(ROOT / "tests/golden/molecules.json"))
names = sys.argv[1:] or ["H", "H2", "LiH", "H2O", "CH4", "SiH4", "NO2", "MB16_43_01", "LYS_xao", "caffeine"]
dev = torch.device("cuda:0")
opts = {"exclude": ["disp"]}
nat_max = max(len(mols[n]["numbers"]) for n in names)
numbers = torch.zeros((len(names), nat_max), dtype=torch.long)
positions = torch.zeros((len(names), nat_max, 3), dtype=torch.float64)
chrg = torch.zeros(len(names), dtype=torch.float64)
for i, n in enumerate(names):
    k = len(mols[n]["numbers"])
    numbers[i, :k] = torch.tensor(mols[n]["numbers"])
    positions[i, :k] = torch.tensor(mols[n]["positions"], dtype=torch.float64)
    chrg[i] = mols[n]["charge"]
numbers, positions, chrg = numbers.to(dev), positions.to(dev), chrg.to(dev)
calc = GFN1Calculator(numbers, opts=opts, device=dev, dtype=torch.float64)
pos = positions.clone().requires_grad_(True)
t = time.time()
e = calc.get_energy(pos, chrg)
torch.cuda.synchronize()
print("forward time %.3fs" % (time.time() - t), "use_smem", calc._variants)
ws = calc.cache["ws"]
(g,) = torch.autograd.grad(e.sum(), pos)
torch.cuda.synchronize()
d = calc.desc
print("status", calc.cache["status"].tolist(), "iters", ws.iterations.tolist())
for i, n in enumerate(names):
    m = mols[n]
    r = O.singlepoint(m["numbers"], np.array(m["positions"]), chrg=m["charge"], opts=dict(exclude=("disp",)), grad=True)
    a0, a1 = d.at_off[i], d.at_off[i + 1]
    s0, s1 = d.sh_off[i], d.sh_off[i + 1]
    o0, o1 = d.ao_off[i], d.ao_off[i + 1]
    nao = o1 - o0
    S = ws.S[d.mat_off[i] : d.mat_off[i + 1]].cpu().numpy().reshape(nao, nao)
    H = ws.H0[d.mat_off[i] : d.mat_off[i + 1]].cpu().numpy().reshape(nao, nao)
    P = ws.P[d.mat_off[i] : d.mat_off[i + 1]].cpu().numpy().reshape(nao, nao)
    W = ws.W[d.mat_off[i] : d.mat_off[i + 1]].cpu().numpy().reshape(nao, nao)
    mm = O.make_mol(m["numbers"])
    gam = O.gamma_shell(mm, np.array(m["positions"]))
    q0 = O.eeq_charges(mm, np.array(m["positions"]), m["charge"])
    erep, _ = O.repulsion(mm, np.array(m["positions"]))
    print(
        f"{n:12s} nao={nao:4d} it={int(ws.iterations[i])}/{r.iterations}"
        f" dE={float(e[i].detach()) - r.energy:+.2e} dS={np.abs(S - r.S).max():.1e} dH0={np.abs(H - r.H0).max():.1e}"
        f" dcn={np.abs(ws.cn[a0:a1].cpu().numpy() - r.cn).max():.1e}"
        f" drep={np.abs(ws.e_rep[a0:a1].cpu().numpy() - erep).max():.1e}"
        f" dgam={np.abs(ws.gamma[d.gam_off[i]:d.gam_off[i+1]].cpu().numpy().reshape(s1-s0, s1-s0) - gam).max():.1e}"
        f" dq0={np.abs(ws.q0_at[a0:a1].cpu().numpy() - q0).max():.1e}"
        f" dq={np.abs(ws.q_orb[o0:o1].cpu().numpy() - r.q_orb).max():.1e}"
        f" dP={np.abs(P - r.P).max():.1e} dW={np.abs(W - r.W).max():.1e}"
        f" dG={float(ws.fenergy[i]) - r.fenergy:+.1e}"
        f" dgrad={np.abs(g[i, : a1 - a0].cpu().numpy() - r.gradient).max():.1e}"
    )
