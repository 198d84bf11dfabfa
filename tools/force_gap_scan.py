#!/usr/bin/env python
"""Per-conformer force gap of the CUDA path (and of the oracle) to the reference's autograd forces on the config-2 batch."""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from dxtb_b200 import GFN1Calculator
from oracle import gfn1_oracle as O

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    dev = torch.device("cuda:0")
    wl = bench.Workload(2, 1, 1024)
    idx = np.arange(n)
    pos = wl.positions(7, idx)
    calc = GFN1Calculator(torch.from_numpy(wl.numbers[idx]).to(dev), device=dev, dtype=torch.float64, d3_reference=bench._d3_table())
    p = torch.from_numpy(pos).to(dev).requires_grad_(True)
    e = calc.get_energy(p)
    (g,) = torch.autograd.grad(e.sum(), p)
    g = g.cpu().numpy()
    arm = bench.CpuArm()
    _, jobs = arm.jobs(wl, 7, 8)
    jobs = jobs[: n // 8]
    arm.run(jobs[: arm.cores])
    _, out = arm.run(jobs)
    arm.close()
    f_ref = np.concatenate([o[1] for o in out])
    gap = np.abs(-g[: len(f_ref)] - f_ref).reshape(len(f_ref), -1).max(1)
    order = np.argsort(-gap)
    print("max gap", gap.max(), "median", np.median(gap), "n >1e-7:", int((gap > 1e-7).sum()))
    for i in order[:4]:
        r = O.singlepoint(wl.numbers[i], pos[i], 0.0, grad=True, d3_table=bench._d3_table())
        r0 = O.singlepoint(wl.numbers[i], pos[i], 0.0, grad=True, d3_table=bench._d3_table(), opts={"response_tol": 1e-12, "response_maxiter": 40})
        print(i, "cuda-dxtb %.2e  oracle-dxtb %.2e  oracle(tol 1e-12)-dxtb %.2e  cuda-oracle %.2e  it %d |dv| %.1e" % (
            gap[i], np.abs(r.gradient + f_ref[i]).max(), np.abs(r0.gradient + f_ref[i]).max(), np.abs(r.gradient - g[i]).max(), r.iterations,
            np.abs(r.v_orb - r.v_in).max()))


if __name__ == "__main__":  # the CPU arm spawns worker processes that re-import this module
    main()
