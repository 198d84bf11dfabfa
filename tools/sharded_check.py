#!/usr/bin/env python
"""Sharded single points under NCCL == the 1-GPU result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py

Every rank computes its shard of ONE fixed batch (config-3-like drug-like conformers with contiguous shards, and a ragged
config-5-like batch with cost-balanced shards), the per-molecule energies / forces / iteration counts are gathered with
dxtb_b200.parallel, and rank 0 recomputes the whole batch on its own GPU and compares: bit-for-bit for the uniform batch as
long as shard and whole batch run the same launch shape (from 148 equally sized molecules per GPU on the launch is persistent and
molecules start from their predecessor's eigenvectors instead of the Cholesky basis: then 1e-11 Eh / 1e-9 Eh/bohr); for the ragged one the
parity tolerances
(1e-9 Eh, 1e-7 Eh/bohr, equal iteration counts), because the size buckets -- and with them the kernel variant of a molecule,
e.g. one-CTA kernel vs large-system path for the few 550-AO molecules of a shard -- depend on what else is in the shard
(different eigensolver round-off, response solver stopped at 1e-8).  Prints one JSON line; exit code 1 on mismatch.
"""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from dxtb_b200 import GFN1Calculator  # noqa: E402
from dxtb_b200.parallel import gather_by_index  # noqa: E402


def single_points(wl, idx, dev):
    numbers = torch.from_numpy(wl.numbers[idx]).to(dev)
    chrg = torch.from_numpy(wl.chrg[idx]).to(dev)
    calc = GFN1Calculator(numbers, opts={"exclude": ["disp"]}, device=dev, dtype=torch.float64)
    p = torch.from_numpy(wl.positions(0, idx)).to(dev).requires_grad_(True)
    e = calc.get_energy(p, chrg)
    (g,) = torch.autograd.grad(e.sum(), p)
    return e.detach(), g, calc.get_iterations().to(torch.float64)


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    report, ok = {}, True
    for config, n, tol, ftol in ((3, 64 * world, 0.0 if 64 * world <= 128 else 1e-11, 0.0 if 64 * world <= 128 else 1e-9), (5, 48 * world, 1e-9, 1e-7)):
        wl = bench.Workload(config, world, n)
        parts = [wl.shard(r) for r in range(world)]
        e, g, it = single_points(wl, parts[rank], dev)
        e_all = gather_by_index(e, parts, wl.n_total)
        g_all = gather_by_index(g, parts, wl.n_total)
        it_all = gather_by_index(it, parts, wl.n_total)
        if rank == 0:
            e1, g1, it1 = single_points(wl, np.arange(wl.n_total), dev)
            de, dg = float((e_all - e1).abs().max()), float((g_all - g1).abs().max())
            same_it = bool(torch.equal(it_all, it1))
            report[f"config{config}"] = {"n": wl.n_total, "world": world, "max_abs_dE": de, "max_abs_dF": dg, "iterations_equal": same_it,
                                         "tolerance": [tol, ftol], "shard_sizes": [len(p) for p in parts]}
            ok = ok and de <= tol and dg <= ftol and same_it
    if rank == 0:
        report["ok"] = ok
        print(json.dumps(report), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
