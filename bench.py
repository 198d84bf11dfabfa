#!/usr/bin/env python
"""Benchmark of the GFN1-xTB fp64 single-point hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port on all host cores)

One "step" = energy + forces of one batch of 1024 perturbed caffeine conformers per GPU (BASELINE config 2
geometry recipe: N(0, 0.05 bohr) per coordinate, seeded), weak scaling over GPUs, no data-path collective.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "GFN1-xTB fp64 single-points/sec (energy+forces)"
UNIT = "single-points/s"
NB = 1024
CPU_SAMPLE = 320  # conformers of the cpu_baseline leg (about 10 s on one host core)
SIGMA = 0.05
NODISP_NOTE = ("D3(BJ) dispersion is computed on both arms with a SYNTHETIC reference table of the real shape (tad-dftd3's "
               "C6 data is third-party and unavailable offline): its cost is included, its energy is not physical")


MOLECULE = "caffeine"


def load_caffeine():
    m = json.load(open(ROOT / "tests" / "golden" / "molecules.json"))[MOLECULE]
    return np.array(m["numbers"]), np.array(m["positions"])


def conformers(base: np.ndarray, nb: int, seed: int) -> np.ndarray:
    import torch

    g = torch.Generator().manual_seed(seed)
    b = torch.tensor(base, dtype=torch.float64)
    return (b[None] + SIGMA * torch.randn((nb, *b.shape), generator=g, dtype=torch.float64)).numpy()


# --------------------------------------------------------------------------------------------------
# CPU arm: the NumPy oracle ("port"; the reference itself cannot be imported: tad-mctc/tad-dftd3/
# tad-multicharge are absent and there is no network)
# --------------------------------------------------------------------------------------------------
_D3 = None


def _d3_table():
    global _D3
    if _D3 is None:
        from oracle import gfn1_oracle as O

        _D3 = O.synthetic_d3_table()
    return _D3


def _oracle_one(args):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import gfn1_oracle as O

    numbers, pos = args
    r = O.singlepoint(numbers, pos, 0.0, grad=True, d3_table=_d3_table())
    return r.energy


def _jobs(nsample: int, seed: int):
    numbers, base = load_caffeine()
    pos = conformers(base, nsample, seed)
    return [(numbers, pos[i]) for i in range(nsample)]


def cpu_rate(nsample: int, nproc: int, seed: int = 12345) -> tuple[float, float]:
    """single points / s of the oracle on ``nsample`` conformers, single process and single BLAS thread."""
    jobs = _jobs(nsample, seed)
    try:
        from threadpoolctl import threadpool_limits
    except Exception:  # pragma: no cover
        threadpool_limits = None
    _oracle_one(jobs[0])  # warm the parameter / CGTO caches
    t = time.perf_counter()
    if threadpool_limits is not None:
        with threadpool_limits(limits=1):
            for j in jobs:
                _oracle_one(j)
    else:
        for j in jobs:
            _oracle_one(j)
    dt = time.perf_counter() - t
    return nsample / dt, dt


def run_reference(args) -> None:
    """CPU arm: the oracle port on ALL host cores (one worker process per core, pool created once)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    os.environ["OMP_NUM_THREADS"] = "1"
    cores = os.cpu_count() or 1
    nsample = int(os.environ.get("BENCH_REF_SAMPLE_PER_CORE", "8")) * cores  # bounded sample of the 1024-conformer batch per step
    times = []
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_oracle_one, _jobs(cores, 7), chunksize=1)  # start + warm every worker
        for s in range(args.warmup + args.steps):
            jobs = _jobs(nsample, 1000 + s)
            t = time.perf_counter()
            pool.map(_oracle_one, jobs, chunksize=2)
            if s >= args.warmup:
                times.append(time.perf_counter() - t)
    value = float(nsample * len(times) / np.sum(times))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"caffeine x{NB} conformers (C8H10N4O2, 24 atoms, nao 76), energy+forces; CPU arm times a bounded sample",
                   "sigma_bohr": SIGMA, "note": NODISP_NOTE,
                   "why_port": "dxtb itself cannot be imported (tad-mctc / tad-dftd3 / tad-multicharge are not installable offline)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{nsample} conformers per step x {args.steps} steps, oracle/gfn1_oracle.py, {cores} worker processes"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle-reason sampler (NVML, else nvidia-smi): one sample right after the last timed step was
    enqueued, while the GPU still executes the timed region.  Inside the launch loop the SM clock is measured on the
    device instead (xtb_clock_probe): NVML queries there stall kernel submission (measured 5..80 ms per query)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons = index, [], set()
        self.max_mhz = None
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def sample(self):
        t0 = time.perf_counter()
        try:
            if self._nvml is not None:
                nv = self._nvml
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for name, bit in (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown),
                                  ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                                  ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown),
                                  ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap)):
                    if r & bit:
                        self.reasons.add(name)
            else:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
        except Exception:
            pass
        self.cost_ms = getattr(self, "cost_ms", 0.0) + 1e3 * (time.perf_counter() - t0)

    def stop(self):
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "query_ms_total": round(getattr(self, "cost_ms", 0.0), 2)}


def measure_fp64_peak(dev) -> float:
    """cuBLAS DGEMM 4096^3, best of 6 (TFLOP/s): the fp64 denominator (MEASURED_PEAKS.json has none)."""
    import torch

    n = 4096
    a = torch.randn((n, n), dtype=torch.float64, device=dev)
    b = torch.randn((n, n), dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize(dev)
        best = max(best, 2.0 * n**3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from dxtb_b200 import GFN1Calculator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    numbers_np, base = load_caffeine()
    nb = args.nb
    numbers = torch.tensor(numbers_np)[None].expand(nb, -1).contiguous().to(dev)
    chrg = torch.zeros(nb, dtype=torch.float64, device=dev)
    calc = GFN1Calculator(numbers, device=dev, dtype=torch.float64, d3_reference=_d3_table())
    nstep = args.warmup + args.steps
    host = [torch.from_numpy(conformers(base, nb, 100 * rank + s)).pin_memory() for s in range(nstep)]
    devpos = [h.to(dev) for h in host]

    def step(p):
        p = p.detach().requires_grad_(True)
        e = calc.get_energy(p, chrg)
        (g,) = torch.autograd.grad(e.sum(), p)
        return e.detach(), g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing (value) --------------------------------------------------------
    for s in range(args.warmup):
        step(devpos[s])
    calc.scf_events = []
    # pre-created (and once recorded, i.e. materialised) timing events: nothing but launches inside the timed loop
    calc.scf_event_pool = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps + 1)]
    for a, b in calc.scf_event_pool:
        a.record()
        b.record()
    sampler = ClockSampler(local)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iter_tensors = []
    # SM clock under load: a 20 us on-device probe (clock64 / globaltimer) after every step -- no NVML call inside the launch
    # loop: an in-process NVML query (and even a one-shot nvidia-smi in another process) was measured to stall kernel
    # submission by 5..80 ms on some boxes.  Throttle reasons + the NVML clock are read once, right after the last step
    # has been enqueued, i.e. while the GPU is still executing the timed region.
    from dxtb_b200 import _abi
    probe = torch.zeros(args.steps, dtype=torch.float64, device=dev)
    for s in range(args.warmup, nstep):
        step(devpos[s])
        _abi.lib().xtb_clock_probe(probe[s - args.warmup :].data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        iter_tensors.append(calc.get_iterations())
    e1.record()
    if rank == 0:
        sampler.sample()
    barrier()
    clocks = sampler.stop()
    clocks["sm_mhz_on_device"] = [round(float(x), 1) for x in probe.cpu()]
    if clocks["sm_mhz"] is None:
        clocks["sm_mhz"] = float(np.median(clocks["sm_mhz_on_device"]))
    iters_total = sum(int(t.sum()) for t in iter_tensors) + 2 * nb * args.steps  # + final solve + start basis per molecule
    ms = e0.elapsed_time(e1)
    scf_ms = [a.elapsed_time(b) for a, b in calc.scf_events]
    calc.scf_events = None

    # ---- end-to-end timing: pinned host -> device -> energies + forces back on the host -------------
    for s in range(min(args.warmup, 2)):
        step(host[s].to(dev, non_blocking=True))
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    h2d = d2h = 0
    for s in range(args.warmup, nstep):
        p = host[s].to(dev, non_blocking=True)
        e, g = step(p)
        eh, gh = e.cpu(), g.cpu()
        h2d, d2h = p.numel() * 8, (eh.numel() + gh.numel()) * 8
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    value = world * nb * args.steps / (ms * 1e-3)
    e2e = world * nb * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        n = int(calc.desc.nao[0])
        flops_per_launch = 10.0 * n**3 * iters_total / args.steps  # W_iter = 10 n^3 per SCF map evaluation
        scf_avg_ms = float(np.mean(scf_ms))
        peak = measure_fp64_peak(dev)
        achieved = flops_per_launch / (scf_avg_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        except Exception:
            pass
        bytes_per_launch = (48.0 * n * n + 8.0 * float(calc.desc.nsh[0]) ** 2) * iters_total / args.steps
        cpu_v, cpu_dt = cpu_rate(CPU_SAMPLE, 1)  # ~10 s of single-thread oracle work
        # parity gate beside the throughput number: a few molecules of the last batch against the oracle
        from oracle import gfn1_oracle as O

        p_last = devpos[nstep - 1].detach().requires_grad_(True)
        e_last = calc.get_energy(p_last, chrg)
        (g_last,) = torch.autograd.grad(e_last.sum(), p_last)
        it_last = calc.get_iterations()
        de = dg = 0.0
        it_equal = True
        idx = [0, nb // 3, nb - 1]
        for i in idx:
            r = O.singlepoint(numbers_np, host[nstep - 1][i].numpy(), 0.0, grad=True, d3_table=_d3_table())
            de = max(de, abs(float(e_last[i]) - r.energy))
            dg = max(dg, float(np.abs(g_last[i].cpu().numpy() - r.gradient).max()))
            it_equal = it_equal and int(it_last[i]) == r.iterations
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{MOLECULE} x{nb} conformers per GPU ({len(numbers_np)} atoms, nao {n}), energy+forces (BASELINE config 2 geometry recipe)",
                       "sigma_bohr": SIGMA, "l2": "a new conformer batch every step; per-step working set (S,H0,P,W ~190 MB) exceeds L2",
                       "opts": "dxtb defaults (EEQ guess, Anderson, x_atol 1e-4/1e-5, 300 K, D3(BJ) with synthetic table)", "note": NODISP_NOTE},
            "roofline": {"bound": "tensor", "pipe": "fp64 (DFMA/DMMA)", "kernel": "k_scf", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": 17.909760e6 / 148 * nb if MOLECULE == "caffeine" else None,  # DRAM bytes of k_scf per molecule, profiles/r1_scf_r8_ncu_full.csv

                         "peak_source": "cuBLAS DGEMM 4096^3 best-of-6 measured in this run (MEASURED_PEAKS.json has no fp64 entry)",
                         "scf_kernel_ms": scf_avg_ms, "scf_share_of_step": scf_avg_ms / (ms / args.steps),
                         "hbm_equiv_gbs": bytes_per_launch / (scf_avg_ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": f"{CPU_SAMPLE} conformers energy+forces, oracle/gfn1_oracle.py single thread ({cpu_dt:.1f} s)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": 23 * args.steps,  # 13 forward + 9 backward kernels + the clock probe per step
            "clocks": clocks,
            "scf_iterations_mean": (iters_total / args.steps - 2 * nb) / nb,
            "parity": {"checked": len(idx), "max_abs_dE_Eh": de, "max_abs_dF_Eh_per_bohr": dg, "scf_iterations_equal": it_equal,
                       "against": "oracle/gfn1_oracle.py (tolerances: 1e-9 Eh, 1e-7 Eh/bohr)"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nb", type=int, default=NB)
    ap.add_argument("--molecule", default="caffeine", help="fixture geometry to make conformers of (default: BASELINE config 2)")
    args = ap.parse_args()
    global MOLECULE
    MOLECULE = args.molecule
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
