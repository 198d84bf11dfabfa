#!/usr/bin/env python
"""Oracle results for EVERY molecule of the BASELINE config-2 batch (1024 perturbed caffeine conformers, bench.Workload(2)
at step 0, exclude=["disp"]): energies, SCF iteration counts and the atomic charges of every conformer plus forces of
every 16th, as a committed fixture for the full-size `-m gpu` gate (tests/test_gpu_parity.py).  ~30 s on 8 cores.

    python tests/golden/make_config2_oracle.py
"""
import multiprocessing as mp
import os
import sys

os.environ.setdefault("OMP_NUM_THREADS", "1")  # one BLAS thread per worker (8 x 8 oversubscription costs 30x)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("MKL_NUM_THREADS", "1")
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def one(args):
    from oracle import gfn1_oracle as O

    i, z, p = args
    r = O.singlepoint(z, p, 0.0, opts={"exclude": ("disp",)}, grad=(i % 16 == 0))
    return r.energy, r.iterations, r.q_at, (r.gradient if i % 16 == 0 else None)


def main():
    wl = bench.Workload(2, 1, 1024)
    pos = wl.positions(0, np.arange(1024))
    jobs = [(i, wl.numbers[i], pos[i]) for i in range(1024)]
    with mp.Pool() as pool:
        out = pool.map(one, jobs, chunksize=8)
    np.savez_compressed(ROOT / "tests/golden/config2_oracle.npz",
                        energy=np.array([o[0] for o in out]), iterations=np.array([o[1] for o in out], dtype=np.int32),
                        q_at=np.array([o[2] for o in out]), gradient=np.array([o[3] for o in out[::16]]),
                        positions_checksum=float(np.abs(pos).sum()))
    it = np.array([o[1] for o in out])
    print("iterations: min", it.min(), "max", it.max(), "mean", it.mean())


if __name__ == "__main__":
    main()
