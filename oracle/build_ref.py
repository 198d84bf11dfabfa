#!/usr/bin/env python
"""Build recipe for running the UNMODIFIED reference (dxtb v0.4.0) as the CPU oracle / CPU baseline.

    python oracle/build_ref.py

1. installs /root/reference (pure Python) without its dependencies into baseline/_ref (git-ignored, NOT gpurun-ignored,
   so it travels to the GPU box):  pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy>
   (the copy under /tmp is needed because the build writes egg-info into the source tree and /root/reference is read-only;
   with dependencies the install fails: h5py, tad-mctc, tad-dftd3, tad-dftd4, tad-multicharge are not in the wheelhouse);
2. checks that ``import dxtb`` works behind the stand-in dependency packages of oracle/shim (committed; see its README).

``reference_paths()`` is what tests/golden/make_reference_runs.py and bench.py --impl reference put on sys.path.
TEST INFRASTRUCTURE ONLY: nothing under dxtb_b200/ uses any of this.
"""
from __future__ import annotations

import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
TARGET = ROOT / "baseline" / "_ref"
SHIM = ROOT / "oracle" / "shim"
SOURCE = Path("/root/reference")


def reference_paths() -> list[str]:
    return [str(SHIM), str(TARGET)]


def available() -> bool:
    return (TARGET / "dxtb" / "__init__.py").exists()


def install(force: bool = False) -> bool:
    """Returns True if baseline/_ref holds the reference afterwards."""
    if available() and not force:
        return True
    if not SOURCE.exists():
        return False
    with tempfile.TemporaryDirectory() as tmp:
        src = Path(tmp) / "dxtb_src"
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns(".git", "test", "docs", "examples"))
        if TARGET.exists():
            shutil.rmtree(TARGET)
        TARGET.parent.mkdir(exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", str(TARGET), str(src)]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            print(r.stdout)
            return False
    return available()


def check_import() -> str:
    code = ("import sys; sys.path[:0] = %r; import dxtb; from dxtb.calculators import GFN1Calculator; print(dxtb.__version__)"
            % reference_paths())
    r = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("import dxtb failed behind oracle/shim:\n" + r.stdout)
    return r.stdout.strip().splitlines()[-1]


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print("baseline/_ref:", "installed" if ok else "NOT available (no /root/reference here and no previous install)")
    if ok:
        print("import dxtb ->", check_import())
