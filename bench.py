#!/usr/bin/env python
"""Benchmark of the GFN1-xTB fp64 single-point hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W              # our CUDA path, BASELINE config 2 (default, weak scaling)
    python bench.py --impl reference --gpus N --steps K ...    # CPU arm: dxtb itself (baseline/_ref behind oracle/shim) on all cores
    python bench.py --gpus N --config 3|5 ...                  # north-star multi-GPU workloads (strong scaling, sharded + gathered)
    python bench.py --config 4                                 # one 1000-atom system (large-system path)

Workloads (all synthetic geometries, seeded):
  config 2 (default)  1024 perturbed caffeine conformers PER GPU (24 atoms, nao 76), N(0, 0.05 bohr) per coordinate; weak scaling
  config 3            ONE batch of 8192 drug-like conformers (capsaicin C18H27NO3 and its thioether analogue C18H27NO2S, 49 atoms,
                      nao 142/147) sharded over the ranks with parallel.shard_bounds; results gathered; strong scaling
  config 4            sh3 (1027 atoms, nao 3104), one molecule per GPU ("replicas only")
  config 5            ONE ragged batch of 4096 molecules of 16..176 atoms (compositions of examples/molecules and the reference's
                      test set, perturbed), cost-balanced with parallel.shard_by_cost; results gathered; strong scaling
A "step" = energy + forces of the rank's shard of one batch.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import re
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
SIGMA = 0.05
REF_BATCH = 8  # conformers per dxtb call in the CPU arm (dxtb's batched call is ~2x faster per system than single calls)
D3_NOTE = ("D3(BJ) dispersion is computed on both arms with a SYNTHETIC reference table of the real shape (tad-dftd3's "
           "C6 data is third-party and unavailable offline): its cost is included, its energy is not physical")


# --------------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------------
def _mols():
    return json.load(open(ROOT / "tests" / "golden" / "molecules.json"))


def _thio_capsaicin(z: np.ndarray, p: np.ndarray):
    """Capsaicin with the ether oxygen (atom 1) replaced by sulfur, moved 1.2 bohr outwards along the C-O-C bisector."""
    d = np.linalg.norm(p - p[1], axis=1)
    nb = np.flatnonzero((d < 3.0) & (d > 0))
    u = p[1] - p[nb].mean(0)
    z2, p2 = z.copy(), p.copy()
    z2[1] = 16
    p2[1] = p[1] + 1.2 * u / np.linalg.norm(u)
    return z2, p2


def _perturb(base: np.ndarray, seeds: np.ndarray) -> np.ndarray:
    """base (nat, 3) -> (len(seeds), nat, 3); conformer i depends on seeds[i] only (independent of the sharding)."""
    out = np.empty((len(seeds), *base.shape))
    for i, s in enumerate(seeds):
        out[i] = base + SIGMA * np.random.default_rng(int(s)).standard_normal(base.shape)
    return out


class Workload:
    """numbers (n_total, nat_pad) int64, positions(step) -> (n_total, nat_pad, 3) float64, for the WHOLE job."""

    def __init__(self, config: int, world: int, nb: int | None, molecule: str = "caffeine"):
        m = _mols()
        self.config, self.world = config, world
        if config == 2:
            self.strong = False
            z, p = np.array(m[molecule]["numbers"]), np.array(m[molecule]["positions"])
            per = nb or NB
            self.templates = [(z, p)]
            self.assign = np.zeros(per * world, dtype=np.int64)
            self.name = f"{molecule} x{per} conformers per GPU ({len(z)} atoms), energy+forces (BASELINE config 2 geometry recipe)"
        elif config == 3:
            self.strong = True
            z, p = np.array(m["capsaicin"]["numbers"]), np.array(m["capsaicin"]["positions"])
            n = nb or 8192
            self.templates = [(z, p), _thio_capsaicin(z, p)]
            self.assign = (np.arange(n) % 2).astype(np.int64)
            self.name = (f"{n} drug-like conformers (capsaicin C18H27NO3 / thioether analogue C18H27NO2S, 49 atoms, CHNOS), ONE batch "
                         "sharded over the GPUs, energy+forces (BASELINE config 3)")
        elif config == 4:
            self.strong = False
            z, p = np.array(m["ex_sh3"]["numbers"]), np.array(m["ex_sh3"]["positions"])
            self.templates = [(z, p)]
            self.assign = np.zeros((nb or 1) * world, dtype=np.int64)
            self.name = "sh3 (1027 atoms, nao 3104), one molecule per GPU, energy+forces (BASELINE config 4)"
        elif config == 5:
            self.strong = True
            names = ["MB16_43_01", "caffeine", "nicotine", "LYS_xao", "AD7en+", "capsaicin", "C60", "vancoh2"]
            self.templates = [(np.array(m[k]["numbers"]), np.array(m[k]["positions"])) for k in names]
            self.charges = [float(m[k]["charge"]) for k in names]
            n = nb or 4096
            rng = np.random.default_rng(5)
            # mostly small molecules, a tail of big ones (10..200 atoms)
            self.assign = rng.choice(len(names), size=n, p=[0.16, 0.2, 0.16, 0.14, 0.12, 0.12, 0.07, 0.03]).astype(np.int64)
            self.name = (f"{n} molecules of 16..176 atoms (compositions {', '.join(names)}; perturbed geometries), ONE ragged batch "
                         "cost-balanced over the GPUs, energy+forces (BASELINE config 5)")
        else:
            raise ValueError("config must be 2, 3, 4 or 5")
        self.n_total = len(self.assign)
        self.nat_pad = max(len(z) for z, _ in self.templates)
        self.numbers = np.zeros((self.n_total, self.nat_pad), dtype=np.int64)
        for t, (z, _) in enumerate(self.templates):
            self.numbers[self.assign == t, : len(z)] = z
        ch = getattr(self, "charges", [0.0] * len(self.templates))
        self.chrg = np.array([ch[t] for t in self.assign])

    def positions(self, step: int, idx: np.ndarray) -> np.ndarray:
        """Geometries of molecules ``idx`` (global ids) at ``step``."""
        out = np.zeros((len(idx), self.nat_pad, 3))
        for t, (z, p) in enumerate(self.templates):
            sel = np.flatnonzero(self.assign[idx] == t)
            if sel.size:
                out[sel, : len(z)] = _perturb(p, 1_000_003 * (step + 1) + idx[sel])
        return out

    def shard(self, rank: int) -> np.ndarray:
        """Global molecule ids of ``rank``."""
        from dxtb_b200.parallel import shard_bounds, shard_by_cost

        if self.config == 5:
            import torch

            nao = np.array([_nao(z) for z, _ in self.templates])[self.assign]
            return shard_by_cost(torch.from_numpy(nao.astype(np.float64) ** 3), self.world)[rank].numpy()
        a, b = shard_bounds(self.n_total, self.world, rank)
        return np.arange(a, b)


def _nao(z: np.ndarray) -> int:
    from dxtb_b200.param import gfn1_param

    par = gfn1_param()
    return int(sum((2 * par.ang[int(a), : int(par.nshell[int(a)])] + 1).sum() for a in z))


def load_caffeine(molecule: str = "caffeine"):
    m = _mols()[molecule]
    return np.array(m["numbers"]), np.array(m["positions"])


def conformers(base: np.ndarray, nb: int, seed: int) -> np.ndarray:
    """nb perturbed copies of ``base`` (tools/*.py)."""
    return _perturb(base, 7919 * (seed + 1) + np.arange(nb))


_D3 = None


def _d3_table():
    global _D3
    if _D3 is None:
        from oracle import gfn1_oracle as O

        _D3 = O.synthetic_d3_table()
    return _D3


# --------------------------------------------------------------------------------------------------
# CPU arm.  The reference itself: dxtb v0.4.0 installed unmodified into baseline/_ref (oracle/build_ref.py) with its
# un-vendored utility dependencies replaced by oracle/shim; one worker process per host core, one torch thread each
# (intra-op threading does not speed dxtb up: 3.3 -> 3.5 SP/s from 1 to 8 threads), every worker runs dxtb's batched
# get_energy + autograd forces on REF_BATCH systems.  Fallback (kind "port") when baseline/_ref is absent: the NumPy oracle.
# --------------------------------------------------------------------------------------------------
_CALCS: dict = {}


def _ref_init(paths, d3_path):
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["TAD_DFTD3_SHIM_TABLE"] = d3_path
    import warnings

    warnings.simplefilter("ignore")
    sys.path[:0] = paths
    import torch

    torch.set_num_threads(1)
    import dxtb  # noqa: F401


def _ref_job(job):
    """One dxtb call: numbers (B, nat), positions (B, nat, 3), charges (B,) -> energies; forces are computed and dropped."""
    import torch
    from dxtb.calculators import GFN1Calculator

    numbers, pos, chrg = job
    key = numbers.tobytes()
    if key not in _CALCS:
        _CALCS[key] = GFN1Calculator(torch.from_numpy(numbers), opts={"verbosity": 0}, dtype=torch.float64)
    calc = _CALCS[key]
    calc.reset()
    p = torch.from_numpy(pos).requires_grad_(True)
    e = calc.get_energy(p, torch.from_numpy(chrg))
    (g,) = torch.autograd.grad(e.sum(), p)
    return e.detach().numpy(), (-g).numpy()


def _port_job(job):
    from oracle import gfn1_oracle as O

    numbers, pos, chrg = job
    out = [O.singlepoint(numbers[i], pos[i], float(chrg[i]), grad=True, d3_table=_d3_table()) for i in range(len(numbers))]
    return np.array([r.energy for r in out]), None


class CpuArm:
    def __init__(self, cores: int | None = None):
        import multiprocessing as mp
        import tempfile

        from oracle.build_ref import available, reference_paths

        self.cores = cores or (os.cpu_count() or 1)
        self.kind = "reference" if available() else "port"
        self._tmp = tempfile.TemporaryDirectory()
        d3_path = os.path.join(self._tmp.name, "d3_synthetic.npz")
        np.savez(d3_path, **_d3_table())
        os.environ["OMP_NUM_THREADS"] = "1"
        if self.kind == "reference":
            self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_ref_init, initargs=(reference_paths(), d3_path))
            self.fn = _ref_job
        else:
            self.pool = mp.get_context("fork").Pool(self.cores)
            self.fn = _port_job
        self.what = ("dxtb 0.4.0 (baseline/_ref, unmodified; utility deps = oracle/shim), batched get_energy + autograd forces"
                     if self.kind == "reference" else "oracle/gfn1_oracle.py (NumPy port; baseline/_ref not installed)")

    def jobs(self, wl: Workload, step: int, per_worker: int):
        """cores x per_worker systems of the workload (the first ids of it), grouped by composition into dxtb batches."""
        n = min(wl.n_total, self.cores * per_worker)
        idx = np.arange(n)
        pos = wl.positions(step, idx)
        jobs = []
        for t in np.unique(wl.assign[idx]):
            sel = np.flatnonzero(wl.assign[idx] == t)
            nat = len(wl.templates[t][0])
            for k in range(0, len(sel), per_worker):
                s = sel[k : k + per_worker]
                jobs.append((np.ascontiguousarray(wl.numbers[s][:, :nat]), np.ascontiguousarray(pos[s][:, :nat]), wl.chrg[s].copy()))
        return n, jobs

    def run(self, jobs):
        t = time.perf_counter()
        out = self.pool.map(self.fn, jobs, chunksize=1)
        return time.perf_counter() - t, out

    def close(self):
        self.pool.close()
        self.pool.join()
        self._tmp.cleanup()


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = Workload(args.config, 1, args.nb, args.molecule)
    per_worker = int(os.environ.get("BENCH_REF_SAMPLE_PER_CORE", str(REF_BATCH if args.config in (2, 3) else 2)))
    arm = CpuArm()
    n, jobs = arm.jobs(wl, 0, per_worker)
    arm.run(jobs[: arm.cores])  # start + warm every worker (imports, parameter tables, calculators)
    times = []
    for s in range(args.warmup + args.steps):
        n, jobs = arm.jobs(wl, 100 + s, per_worker)
        dt, _ = arm.run(jobs)
        if s >= args.warmup:
            times.append(dt)
    arm.close()
    value = float(n * len(times) / np.sum(times))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
        "scaling": "strong" if wl.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.name + "; the CPU arm times a bounded sample of it per step", "sigma_bohr": SIGMA, "note": D3_NOTE},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "per_core": value / arm.cores, "kind": arm.kind,
                         "sample": f"{n} systems per step x {args.steps} steps ({per_worker} per dxtb call), {arm.what}, "
                                   f"{arm.cores} worker processes x 1 thread"},
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


_OURS = re.compile(r"\bkl?_[a-z0-9_]+")


def count_launches(fn, dev) -> tuple[int, dict]:
    """Kernels of dxtb_b200/_C.so (names k_* / kl_*) launched by one call of ``fn``, counted by the CUDA profiler (CUPTI
    through torch.profiler) in an extra, untimed step; every timed step launches the same sequence."""
    import torch
    from torch.profiler import ProfilerActivity, profile

    torch.cuda.synchronize(dev)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize(dev)
    names: dict[str, int] = {}
    other = 0
    for ev in prof.events():
        if "cuda" not in str(ev.device_type).lower():
            continue
        m = _OURS.search(ev.name)
        if m and not ev.name.startswith(("Memcpy", "Memset")):
            names[m.group(0)] = names.get(m.group(0), 0) + 1
        else:
            other += 1
    return sum(names.values()), {"ours": names, "library_or_copy": other}


def scf_traffic() -> tuple[float | None, str | None]:
    """DRAM bytes (read + write) of one k_scf launch per molecule from the committed ncu --set full capture."""
    p = ROOT / "profiles" / "k_scf_traffic.json"
    if not p.exists():
        return None, None
    d = json.loads(p.read_text())
    return float(d["dram_bytes_per_molecule"]), d.get("source")


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from dxtb_b200 import GFN1Calculator, _abi
    from dxtb_b200.parallel import gather_by_index

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = Workload(args.config, world, args.nb, args.molecule)
    parts = [wl.shard(r) for r in range(world)]
    mine = parts[rank]
    nb = len(mine)
    numbers = torch.from_numpy(wl.numbers[mine]).to(dev)
    chrg = torch.from_numpy(wl.chrg[mine]).to(dev)
    if args.config == 4 and nb == 1:
        numbers = numbers[0]
    calc = GFN1Calculator(numbers, device=dev, dtype=torch.float64, d3_reference=_d3_table())
    nstep = args.warmup + args.steps

    def geom(s):
        p = torch.from_numpy(wl.positions(s, mine))
        return p[0] if numbers.ndim == 1 else p

    host = [geom(s).pin_memory() for s in range(nstep)]
    devpos = [h.to(dev) for h in host]
    chrg_arg = chrg[0] if numbers.ndim == 1 else chrg

    def step(p):
        p = p.detach().requires_grad_(True)
        e = calc.get_energy(p, chrg_arg)
        (g,) = torch.autograd.grad(e.sum(), p)
        return e.detach(), g

    def gather(e, g):
        """The only communication of a sharded single point: per-molecule results to every rank (NCCL all_gather)."""
        if world == 1:
            return e, g
        return gather_by_index(e.reshape(nb), parts, wl.n_total), gather_by_index(g.reshape(nb, -1, 3), parts, wl.n_total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing (value) --------------------------------------------------------
    for s in range(args.warmup):
        gather(*step(devpos[s]))
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
    # loop (measured to stall kernel submission by 5..80 ms on some boxes).  Throttle reasons + the NVML clock are read once,
    # right after the last step has been enqueued, i.e. while the GPU is still executing the timed region.
    probe = torch.zeros(args.steps, dtype=torch.float64, device=dev)
    for s in range(args.warmup, nstep):
        gather(*step(devpos[s]))
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
    ms = e0.elapsed_time(e1)
    scf_ms = [a.elapsed_time(b) for a, b in calc.scf_events]
    calc.scf_events = None

    # ---- end-to-end timing: pinned host -> device -> energies + forces of the whole job back on the host -------------
    eh = gh = None  # pinned result buffers (allocated once, like the pinned inputs)
    for s in range(min(args.warmup, 2)):
        e, g = gather(*step(host[s].to(dev, non_blocking=True)))
        if eh is None:
            eh = torch.empty(e.shape, dtype=e.dtype, pin_memory=True)
            gh = torch.empty(g.shape, dtype=g.dtype, pin_memory=True)
    if eh is None:
        e, g = gather(*step(host[0].to(dev, non_blocking=True)))
        eh, gh = torch.empty(e.shape, dtype=e.dtype, pin_memory=True), torch.empty(g.shape, dtype=g.dtype, pin_memory=True)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    h2d = d2h = 0
    for s in range(args.warmup, nstep):
        p = host[s].to(dev, non_blocking=True)
        e, g = gather(*step(p))
        eh.copy_(e.detach(), non_blocking=True)
        gh.copy_(g, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()  # energies and forces of this step are on the host
        h2d, d2h = p.numel() * 8, (eh.numel() + gh.numel()) * 8
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)

    my_ms = ms
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        per_rank = [torch.zeros(2, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(per_rank, torch.tensor([my_ms / args.steps, float(np.mean(scf_ms))], dtype=torch.float64, device=dev))
        per_rank = [[round(float(x[0]), 3), round(float(x[1]), 3)] for x in per_rank]
    else:
        per_rank = [[round(my_ms / args.steps, 3), round(float(np.mean(scf_ms)), 3)]]
    n_job = wl.n_total  # systems the whole job processes per step (all ranks)
    value = n_job * args.steps / (ms * 1e-3)
    e2e = n_job * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        nao = calc.desc.nao.astype(np.float64)
        it = torch.stack(iter_tensors).to(torch.float64).mean(0).cpu().numpy()  # mean map evaluations per molecule
        # 10 n^3 per SCF map evaluation (+ the final solve), n^3/3 for the Cholesky start basis
        flops_per_launch = float((10.0 * nao**3 * (it + 1.0) + nao**3 / 3.0).sum())
        bytes_per_launch = float(((48.0 * nao**2 + 8.0 * calc.desc.nsh.astype(np.float64) ** 2) * (it + 1.0)).sum())
        scf_avg_ms = float(np.mean(scf_ms))
        peak = measure_fp64_peak(dev)
        achieved = flops_per_launch / (scf_avg_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        except Exception:
            pass
        launches, launch_names = count_launches(lambda: (step(devpos[nstep - 1]), _abi.lib().xtb_clock_probe(
            probe.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)), dev)
        traffic_per_mol, traffic_src = scf_traffic() if args.config == 2 and args.molecule == "caffeine" else (None, None)

        # parity gate beside the throughput number: molecules of the last batch against the oracle
        from oracle import gfn1_oracle as O

        p_last = devpos[nstep - 1].detach().requires_grad_(True)
        e_last = calc.get_energy(p_last, chrg_arg)
        (g_last,) = torch.autograd.grad(e_last.sum(), p_last)
        it_last = calc.get_iterations().reshape(-1)
        e_last, g_last = e_last.reshape(-1), g_last.reshape(nb, -1, 3)
        de = dg = 0.0
        it_equal = True
        order = np.argsort(nao)
        idx = sorted({int(order[0]), int(order[len(order) // 2]), int(order[-1])}) if args.config != 4 else []
        if args.config == 2:
            idx = [0, nb // 3, nb - 1]
        pos_np = host[nstep - 1].numpy().reshape(nb, -1, 3)
        for i in idx:
            z = wl.numbers[mine[i]]
            r = O.singlepoint(z, pos_np[i][: (z > 0).sum()], float(wl.chrg[mine[i]]), grad=True, d3_table=_d3_table())
            de = max(de, abs(float(e_last[i]) - r.energy))
            dg = max(dg, float(np.abs(g_last[i, : r.gradient.shape[0]].cpu().numpy() - r.gradient).max()))
            it_equal = it_equal and int(it_last[i]) == r.iterations

        # CPU baseline of the SAME workload: the reference itself on all host cores, ~10-30 s of CPU work (N = 1 only)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            arm = CpuArm()
            per_worker = REF_BATCH if args.config in (2, 3) else 1
            n, jobs = arm.jobs(wl, nstep - 1, per_worker)
            if args.config == 4:
                jobs, n = jobs[:1], 1
            arm.run(jobs[: arm.cores] if args.config != 4 else [])  # warm the workers
            dt, out = arm.run(jobs)
            if arm.kind == "reference" and args.config == 2:  # the live reference is a second parity checker
                e_ref = np.concatenate([o[0].reshape(-1) for o in out])
                f_ref = np.concatenate([o[1].reshape(-1, *o[1].shape[-2:]) for o in out])
                cpu_de = float(np.abs(e_last[: len(e_ref)].detach().cpu().numpy() - e_ref).max())
                gaps = np.abs(-g_last[: len(f_ref)].cpu().numpy() - f_ref).reshape(len(f_ref), -1).max(1)
                cpu_df = float(gaps.max())
                cpu_df_stats = {"median": float(np.median(gaps)), "fraction_within_1e-7": float((gaps < 1e-7).mean())}
            else:
                cpu_de = cpu_df = cpu_df_stats = None
            arm.close()
            cpu = {"value": n / dt, "unit": UNIT, "cores": arm.cores, "per_core": n / dt / arm.cores, "kind": arm.kind,
                   "sample": f"{n} systems of the last batch, energy+forces, {arm.what}, {arm.cores} worker processes x 1 thread ({dt:.1f} s)",
                   "max_abs_dE_vs_cuda_Eh": cpu_de, "max_abs_dF_vs_cuda_Eh_per_bohr": cpu_df, "dF_vs_cuda": cpu_df_stats,
                   "note": "forces of dxtb are autograd through the unrolled SCF at default thresholds; the CUDA forces are the "
                           "analytic gradient + first-order response of the SCF residual; what remains is second order "
                           "(derivative of the reference's own unconverged trajectory), see DESIGN.md section 5"}

        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if wl.strong else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl.name, "baseline_config": args.config, "systems_per_step_all_ranks": n_job,
                       "nao_min_max": [int(nao.min()), int(nao.max())], "sigma_bohr": SIGMA,
                       "l2": "a new geometry batch every step; per-step working set (S, H0, P, W) exceeds L2" if args.config != 4 else
                             "one 3104-AO system: 3 x 82 MB matrices, new geometry every step",
                       "opts": "dxtb defaults (EEQ guess, Anderson, x_atol 1e-4/1e-5, 300 K, D3(BJ) with synthetic table)", "note": D3_NOTE,
                       "kernel_variants": calc._variants},
            "roofline": {"bound": "tensor", "pipe": "fp64 (DFMA/DMMA)", "kernel": "k_scf" if 3 not in calc._variants else "k_scf / kl_*",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic_per_mol * nb if traffic_per_mol is not None else None, "traffic_source": traffic_src,
                         "peak_source": "cuBLAS DGEMM 4096^3 best-of-6 measured in this run (MEASURED_PEAKS.json has no fp64 entry)",
                         "algorithmic_flops_per_launch": flops_per_launch,
                         "flop_convention": "10 n^3 per SCF map evaluation and final solve + n^3/3 Cholesky start basis (SURVEY 8d)",
                         "scf_kernel_ms": scf_avg_ms, "scf_share_of_step": scf_avg_ms / (my_ms / args.steps),
                         "hbm_equiv_gbs": bytes_per_launch / (scf_avg_ms * 1e-3) / 1e9, "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches * args.steps, "gpu_launches_per_step": launch_names,
            "clocks": clocks,
            "scf_iterations_mean": float(it.mean()),
            "eigensolver": {"jacobi_sweeps_mean": float((calc.cache["status"] >> 8).double().mean()),
                            "occupied_subspace_path": bool(calc.opts["scf_subspace"]),
                            "note": "intermediate iterations of closed-shell molecules with a certified gap >= 50 kT solve for the occupied "
                                    "subspace (Riccati fixed point) instead of diagonalising; the roofline convention still counts the "
                                    "reference's 10 n^3 per map evaluation, so `achieved` is algorithmic, not executed, flops"},
            "per_rank_ms": {"step_and_scf_kernel": per_rank, "systems": [len(p) for p in parts]},
            "parity": {"checked": len(idx), "max_abs_dE_Eh": de, "max_abs_dF_Eh_per_bohr": dg, "scf_iterations_equal": it_equal,
                       "against": "oracle/gfn1_oracle.py (tolerances: 1e-9 Eh, 1e-7 Eh/bohr)"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json config (1-based as listed there)")
    ap.add_argument("--nb", type=int, default=None, help="systems per GPU (config 2/4) or in the whole batch (config 3/5)")
    ap.add_argument("--molecule", default="caffeine", help="config 2: fixture geometry to make conformers of")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
