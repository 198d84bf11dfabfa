"""CPU restatement of the occupied-subspace policy of k_scf (dxtb_b200/csrc/xtb_scf_subspace.cuh) on the oracle's SCF:
tools/subspace_sim.py runs the SCF of a molecule with every iteration diagonalised (cyclic Jacobi in the previous eigenbasis,
the round-1 kernel) and with the certified-gap + Riccati fixed point + Newton inverse policy.  The policy must give the same
iteration count, intermediate charges within 2e-9 e of scipy's eigh in EVERY iteration, and need fewer sweeps.  (The CUDA
path is checked against the all-diagonalising CUDA path and the oracle / dxtb fixtures in the -m gpu tests.)"""
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))


@pytest.mark.parametrize("name", ["H2O", "CH4", "SiH4"])
def test_subspace_policy_reproduces_the_diagonalising_scf(name):
    import subspace_sim as sim

    full = sim.run(name, 3, False)
    fast = sim.run(name, 3, True)
    assert fast["iters"] == full["iters"]
    assert fast["maxerrq"] < 2e-9 and full["maxerrq"] < 2e-9
    assert fast["fastit"] >= fast["iters"] - 1  # every intermediate solve but possibly the first took the subspace path
    assert fast["sw"] + fast["final_sw"] < 0.6 * (full["sw"] + full["final_sw"])


def test_gap_certificate_is_a_lower_bound():
    """Gershgorin inside the two diagonal blocks + Cauchy interlacing: the certified gap never exceeds the true gap."""
    import numpy as np

    rng = np.random.default_rng(0)
    for _ in range(50):
        n, no = 24, 9
        d = np.sort(rng.normal(size=n))
        d[no:] += 1.0
        e = rng.normal(scale=rng.choice([1e-3, 1e-2, 5e-2]), size=(n, n))
        a = np.diag(d) + 0.5 * (e + e.T) * (1 - np.eye(n))
        off = np.abs(a - np.diag(np.diag(a)))
        cert = (np.diag(a)[no:] - off[no:, no:].sum(1)).min() - (np.diag(a)[:no] + off[:no, :no].sum(1)).max()
        w = np.linalg.eigvalsh(a)
        assert cert <= w[no] - w[no - 1] + 1e-12
