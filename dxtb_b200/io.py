"""On-disk formats -> tensors for the calculator: xyz, Turbomole coord, .CHRG / .UHF.

Host-side mirror of what the reference gets from ``tad_mctc.io.read`` at ``cli/driver.py:70-146``: atomic numbers
(int64) and positions in bohr (float64); the total charge / number of unpaired electrons default to the ``.CHRG`` /
``.UHF`` files next to the structure file.
"""
from __future__ import annotations

from pathlib import Path

import torch

__all__ = ["AA2AU", "SYMBOLS", "read_xyz", "read_coord", "read_structure", "read_chrg", "read_spin", "pack"]

AA2AU = 1.0 / 0.529177210903  # CODATA 2018 bohr radius in Angstrom

SYMBOLS = (
    "X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr "
    "Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir "
    "Pt Au Hg Tl Pb Bi Po At Rn"
).split()
_Z = {s.lower(): i for i, s in enumerate(SYMBOLS)}


def _number(token: str) -> int:
    t = token.strip()
    if t.isdigit():
        z = int(t)
    else:
        z = _Z.get("".join(ch for ch in t if ch.isalpha()).lower(), 0)
    if not 1 <= z <= 86:
        raise ValueError(f"unknown element '{token}' (GFN1-xTB covers Z = 1..86)")
    return z


def read_xyz(path: str | Path, frame: int | None = 0) -> tuple[torch.Tensor, torch.Tensor]:
    """xyz file (Angstrom).  ``frame``: index of the frame of a multi-frame file, ``None`` -> all frames stacked as
    ``(nframes, nat)`` / ``(nframes, nat, 3)`` (frames must hold the same number of atoms, e.g. a conformer ensemble)."""
    lines = Path(path).read_text().splitlines()
    frames, i = [], 0
    while i < len(lines):
        if not lines[i].strip():
            i += 1
            continue
        try:
            nat = int(lines[i].split()[0])
        except ValueError as e:
            raise ValueError(f"{path}: line {i + 1}: expected the number of atoms") from e
        body = lines[i + 2 : i + 2 + nat]
        if len(body) < nat:
            raise ValueError(f"{path}: frame {len(frames)} is truncated ({len(body)} of {nat} atoms)")
        nums, pos = [], []
        for ln in body:
            t = ln.split()
            if len(t) < 4:
                raise ValueError(f"{path}: malformed atom line '{ln}'")
            nums.append(_number(t[0]))
            pos.append([float(x) * AA2AU for x in t[1:4]])
        frames.append((nums, pos))
        i += 2 + nat
    if not frames:
        raise ValueError(f"{path}: no structure found")
    if frame is not None:
        nums, pos = frames[frame]
        return torch.tensor(nums, dtype=torch.int64), torch.tensor(pos, dtype=torch.float64)
    if len({len(f[0]) for f in frames}) != 1:
        raise ValueError(f"{path}: frames with different numbers of atoms cannot be stacked; read them one by one")
    return (torch.tensor([f[0] for f in frames], dtype=torch.int64), torch.tensor([f[1] for f in frames], dtype=torch.float64))


def read_coord(path: str | Path) -> tuple[torch.Tensor, torch.Tensor]:
    """Turbomole ``$coord`` block (bohr by default; ``$coord angs`` in Angstrom)."""
    nums, pos, scale, on = [], [], 1.0, False
    for ln in Path(path).read_text().splitlines():
        s = ln.strip()
        if s.startswith("$"):
            if on:
                break
            if s.lower().startswith("$coord"):
                on = True
                opt = s.lower().split()[1:]
                if any(o.startswith("ang") for o in opt):
                    scale = AA2AU
                elif any(o.startswith("frac") for o in opt):
                    raise NotImplementedError("fractional coordinates (periodic systems) are outside the GFN1 molecular path")
            continue
        if on and s and not s.startswith("#"):
            t = s.split()
            if len(t) < 4:
                raise ValueError(f"{path}: malformed atom line '{ln}'")
            pos.append([float(x) * scale for x in t[:3]])
            nums.append(_number(t[3]))
    if not nums:
        raise ValueError(f"{path}: no $coord block found")
    return torch.tensor(nums, dtype=torch.int64), torch.tensor(pos, dtype=torch.float64)


def read_structure(path: str | Path) -> tuple[torch.Tensor, torch.Tensor]:
    """Dispatch on the file name: ``*.xyz`` -> xyz, ``coord`` / ``*.coord`` / ``*.tmol`` -> Turbomole."""
    p = Path(path)
    name = p.name.lower()
    if name.endswith(".xyz"):
        return read_xyz(p)
    if name == "coord" or name.endswith((".coord", ".tmol")):
        return read_coord(p)
    raise ValueError(f"{path}: unknown structure format (supported: .xyz, coord/.coord/.tmol)")


def _read_int_file(directory: Path, name: str) -> int:
    f = directory / name
    if not f.is_file():
        return 0
    txt = f.read_text().split()
    if not txt:
        return 0
    return int(float(txt[0]))


def read_chrg(structure_path: str | Path) -> int:
    """Total charge from the ``.CHRG`` file next to the structure file (0 if absent), cli/driver.py:93-115."""
    return _read_int_file(Path(structure_path).resolve().parent, ".CHRG")


def read_spin(structure_path: str | Path) -> int:
    """Number of unpaired electrons from the ``.UHF`` file next to the structure file (0 if absent)."""
    return _read_int_file(Path(structure_path).resolve().parent, ".UHF")


def pack(structures: list[tuple[torch.Tensor, torch.Tensor]]) -> tuple[torch.Tensor, torch.Tensor]:
    """Zero-pad a list of (numbers, positions) to ``(nb, nat_max)`` / ``(nb, nat_max, 3)`` (the reference's batch layout)."""
    nmax = max(int(n.numel()) for n, _ in structures)
    numbers = torch.zeros((len(structures), nmax), dtype=torch.int64)
    positions = torch.zeros((len(structures), nmax, 3), dtype=torch.float64)
    for i, (n, p) in enumerate(structures):
        numbers[i, : n.numel()] = n
        positions[i, : n.numel()] = p
    return numbers, positions
