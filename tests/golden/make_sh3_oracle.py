#!/usr/bin/env python
"""Oracle result for BASELINE config 4 (sh3, 1027 atoms, nao 3104) as a committed fixture.

The NumPy oracle needs minutes for this system, so the GPU parity test reads its result from
tests/golden/sh3_oracle.npz instead of running it.  NOT a reference golden: it is the output of
oracle/gfn1_oracle.py (pinned against tblite on the smaller systems, tests/test_oracle_golden.py).

    python tests/golden/make_sh3_oracle.py
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import gfn1_oracle as O  # noqa: E402

m = json.load(open(ROOT / "tests/golden/molecules.json"))["ex_sh3"]
t0 = time.time()
r = O.singlepoint(np.array(m["numbers"]), np.array(m["positions"]), float(m["charge"]), grad=True)
print("oracle sh3: E =", repr(r.energy), "iterations", r.iterations, "time %.1f s" % (time.time() - t0), flush=True)
np.savez_compressed(ROOT / "tests/golden/sh3_oracle.npz", energy=r.energy, iterations=r.iterations, q_atom=r.q_at, e_rep=r.e_rep, e_scf=r.e_scf,
                    gradient=r.gradient)
