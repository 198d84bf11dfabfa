"""Multi-GPU: dxtb_b200.parallel under NCCL, gathered sharded results == the 1-GPU result (needs >= 2 GPUs on the box)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
def test_sharded_single_points_equal_one_gpu_result():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (run under gpurun --gpus 2/8)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={min(n, 8)}", "--master-addr", "127.0.0.1",
           "--master-port", "29511", str(ROOT / "tools" / "sharded_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rep = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert rep["ok"], rep
    assert rep["config3"]["max_abs_dE"] <= rep["config3"]["tolerance"][0] and rep["config3"]["max_abs_dF"] <= rep["config3"]["tolerance"][1]
