"""Command line entry: ``python -m dxtb_b200 mol.xyz [more files] [--grad] [--json]``.

Mirrors the single-point part of the reference driver (``cli/driver.py:70-379``, options of ``cli/argparser.py:221-722``
that belong to the GFN1 path): several files form one batch, charges / spins default to ``.CHRG`` / ``.UHF`` next to
each file, results are printed and optionally written as JSON.
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import torch

from . import io


def parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="dxtb_b200", description="GFN1-xTB single points on a B200 (fp64).")
    p.add_argument("file", nargs="+", help="structure file(s): .xyz, coord/.coord/.tmol; several files run as one batch")
    p.add_argument("-c", "--chrg", type=int, default=None, help="total charge (default: .CHRG next to the file, else 0)")
    p.add_argument("--spin", "--uhf", dest="spin", type=int, default=None, help="unpaired electrons (default: .UHF, else 0)")
    p.add_argument("--method", default="gfn1", choices=["gfn1", "gfn1-xtb"], help="only GFN1-xTB is on this path")
    p.add_argument("--exclude", nargs="*", default=[], choices=["disp", "rep", "hal", "es2", "es3", "scf", "all"])
    p.add_argument("--etemp", "--fermi-etemp", "--fermi_etemp", dest="fermi_etemp", type=float, default=None)
    p.add_argument("--fermi-maxiter", "--fermi_maxiter", dest="fermi_maxiter", type=int, default=None)
    p.add_argument("--fermi-thresh", "--fermi_thresh", dest="fermi_thresh", type=float, default=None)
    p.add_argument("--maxiter", type=int, default=None)
    p.add_argument("--mixer", choices=["anderson", "simple", "broyden"], default=None)
    p.add_argument("--damp", type=float, default=None)
    p.add_argument("--guess", choices=["eeq", "sad"], default=None)
    p.add_argument("--xtol", type=float, default=None, help="SCF convergence threshold (x_atol; x_atol_max = xtol / 10)")
    p.add_argument("--int-cutoff", "--int_cutoff", dest="int_cutoff", type=float, default=None)
    p.add_argument("--force-convergence", "--force_convergence", dest="force_convergence", action="store_true")
    p.add_argument("--grad", "--forces", "--force", dest="grad", action="store_true", help="also compute nuclear gradients")
    p.add_argument("--json", nargs="?", const="dxtb_b200.json", default=None, help="write results to this JSON file")
    p.add_argument("--device", default="cuda:0")
    p.add_argument("--dtype", default="double", choices=["double", "float64", "dbl"])
    p.add_argument("--d3-reference", dest="d3_reference", default=None, help="npz with the D3 reference data (INTEGRATION.md)")
    p.add_argument("-v", "--verbose", action="count", default=0)
    return p


def options(args: argparse.Namespace) -> dict:
    o: dict = {}
    for key in ("fermi_etemp", "fermi_maxiter", "fermi_thresh", "maxiter", "mixer", "damp", "guess", "int_cutoff"):
        v = getattr(args, key)
        if v is not None:
            o[key] = v
    if args.xtol is not None:
        o["x_atol"], o["x_atol_max"] = args.xtol, args.xtol / 10.0
    if args.force_convergence:
        o["force_convergence"] = True
    if args.exclude:
        o["exclude"] = list(args.exclude)
    return o


def run(argv: list[str] | None = None) -> dict:
    """Parse, compute, print; returns the result dictionary (also what ``--json`` writes)."""
    from .calculators import GFN1Calculator

    args = parser().parse_args(argv)
    structures = [io.read_structure(f) for f in args.file]
    chrg = [args.chrg if args.chrg is not None else io.read_chrg(f) for f in args.file]
    spin = [args.spin if args.spin is not None else io.read_spin(f) for f in args.file]
    dev = torch.device(args.device)
    numbers, positions = io.pack(structures)
    numbers, positions = numbers.to(dev), positions.to(dev)
    c = torch.tensor(chrg, dtype=torch.float64, device=dev)
    s = torch.tensor(spin, dtype=torch.float64, device=dev) if any(spin) else None

    t0 = time.perf_counter()
    calc = GFN1Calculator(numbers, opts=options(args), device=dev, dtype=torch.float64, d3_reference=args.d3_reference)
    pos = positions.requires_grad_(args.grad)
    energy = calc.get_energy(pos, c, s)
    grad = torch.autograd.grad(energy.sum(), pos)[0] if args.grad else None
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0

    res: dict = {"method": "GFN1-xTB", "unit": {"energy": "Eh", "gradient": "Eh/bohr", "length": "bohr"}, "wall_s": wall, "systems": []}
    it = calc.get_iterations().cpu()
    q = calc.get_atomic_charges().cpu()
    for i, f in enumerate(args.file):
        nat = int(structures[i][0].numel())
        entry = {"file": str(f), "natoms": nat, "charge": chrg[i], "spin": spin[i], "energy": float(energy[i]),
                 "scf_iterations": int(it[i]), "charges": q[i, :nat].tolist()}
        if grad is not None:
            entry["gradient"] = grad[i, :nat].cpu().tolist()
        res["systems"].append(entry)
        print(f"{f}: E = {entry['energy']:.12f} Eh   ({entry['scf_iterations']} SCF iterations)")
        if grad is not None and args.verbose:
            for z, gvec in zip(structures[i][0].tolist(), entry["gradient"]):
                print(f"  {io.SYMBOLS[z]:2s} {gvec[0]: .10e} {gvec[1]: .10e} {gvec[2]: .10e}")
        elif grad is not None:
            print(f"  |gradient| = {float(grad[i, :nat].norm()):.6e} Eh/bohr")
    if args.json:
        Path(args.json).write_text(json.dumps(res, indent=1))
    return res


def main() -> None:
    try:
        run()
    except (ValueError, NotImplementedError, RuntimeError) as e:
        print(f"dxtb_b200: {e}", file=sys.stderr)
        sys.exit(1)
