#!/usr/bin/env python
"""Developer tool: two conformers of a molecule (argv[2], default caffeine) through a developer build of the library
(DXTB_B200_LIB=..., built with DXTB_B200_NVCC_FLAGS="-DXTB_PROFILE_PHASES -DXTB_DEBUG_SUBSPACE": cycle split, certified gap per
sweep, residual of every fixed-point / response iteration for block 0); kernel variant forced by argv[1] (-1: default)."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from dxtb_b200 import GFN1Calculator
ov = int(sys.argv[1]) if len(sys.argv) > 1 else -1
name = sys.argv[2] if len(sys.argv) > 2 else "caffeine"
dev = torch.device("cuda:0")
numbers_np, base = bench.load_caffeine(name)
nb = 2
numbers = torch.tensor(numbers_np)[None].expand(nb, -1).contiguous().to(dev)
chrg = torch.zeros(nb, dtype=torch.float64, device=dev)
calc = GFN1Calculator(numbers, opts={"exclude": ["disp"]}, device=dev, dtype=torch.float64)
if ov >= 0:
    calc._use_smem_override = ov
p = torch.from_numpy(bench.conformers(base, nb, 0)).to(dev)
e = calc.get_energy(p, chrg)
torch.cuda.synchronize()
print("E", e.cpu().numpy(), "iters", calc.get_iterations().cpu().numpy(), "sweeps", (calc.cache["status"] >> 8).cpu().numpy())
