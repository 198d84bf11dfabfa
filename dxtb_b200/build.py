"""Build the CUDA extension in-tree with nvcc for sm_100a (no JIT cache, no torch extension loader).

The resulting ``dxtb_b200/_C.so`` is a plain C-ABI shared library (see ``include/xtb_b200.h``).
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
INCLUDE = ROOT.parent / "include"
SO_PATH = ROOT / "_C.so"
# secondary build of the SCF kernel (two CTAs per SM, developer knob DXTB_B200_2CTA).  Also tried: two CTAs of 256 threads
# (-DXTB_NT=256 -DXTB_NGRP=8, 128 registers each, no spills) -- hybrid variant for caffeine 21.4 ms against 14.5 ms for the
# shared-memory variant at 1 CTA/SM (energy only, 1024 conformers), capsaicin 61.9 against 56.1 ms.
SECONDARY_FLAGS = ["-DXTB_SECONDARY", "-DXTB_MINB=2", "-DXTB_JACOBI_SMALL_SIN=0.0"]
# (source, object suffix, extra flags): xtb_scf.cu is built twice, see the comment at its top
SOURCES = [("xtb_geometry.cu", "", []), ("xtb_integrals.cu", "", []), ("xtb_scf.cu", "", []), ("xtb_scf_large.cu", "", []), ("xtb_scf.cu", ".2cta", SECONDARY_FLAGS)]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "-diag-suppress", "177",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the dxtb_b200 CUDA extension cannot be built")


def needs_build() -> bool:
    if not SO_PATH.exists():
        return True
    t = SO_PATH.stat().st_mtime
    deps = [CSRC / s[0] for s in SOURCES] + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    return any(d.stat().st_mtime > t for d in deps)


def build_extension(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu for sm_100a and link ``_C.so``. Returns the path of the library."""
    if not force and not needs_build():
        return SO_PATH
    nvcc = nvcc_path()
    objdir = ROOT / "build"
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for s, suffix, extra in SOURCES:
        obj = objdir / (s + suffix + ".o")
        dev_flags = os.environ.get("DXTB_B200_NVCC_FLAGS", "").split()  # developer builds, e.g. -DXTB_PROFILE_PHASES
        cmd = [nvcc, *NVCC_FLAGS, *extra, *dev_flags, f"-I{INCLUDE}", f"-I{CSRC}", "-c", str(CSRC / s), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd))
        procs.append((s + suffix, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{out}")
        if verbose and out:
            print(out)
    cmd = [nvcc, "-shared", "-o", str(SO_PATH), *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return SO_PATH


if __name__ == "__main__":
    print(build_extension(force=True, verbose=True))
