"""CPU-side checks: the C-ABI library loads and exports every declared symbol, struct layouts agree with the
header, and the host-side descriptor / option handling mirrors the reference interface."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from dxtb_b200 import _abi
from dxtb_b200.batch import BatchDescriptor
from dxtb_b200.param import gfn1_param
from oracle import gfn1_oracle as O

ROOT = Path(__file__).resolve().parent.parent


def _ensure_built():
    from dxtb_b200.build import build_extension

    build_extension()


def test_library_exports_every_declared_symbol():
    _ensure_built()
    header = (ROOT / "include" / "xtb_b200.h").read_text()
    declared = set(re.findall(r"^(?:int|int64_t)\s+(xtb_\w+)\s*\(", header, flags=re.M))
    assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
    lib = _abi.lib()
    for name in declared:
        assert hasattr(lib, name)
    assert lib.xtb_version() >= 100
    assert lib.xtb_sizeof_batch() == ctypes.sizeof(_abi.XtbBatch)
    assert lib.xtb_sizeof_scf_opts() == ctypes.sizeof(_abi.XtbScfOpts)


def test_descriptor_matches_oracle_index_maps(mols):
    names = ["H2O", "SiH4", "LYS_xao", "MB16_43_01"]
    nat = max(len(mols[n]["numbers"]) for n in names)
    numbers = torch.zeros((len(names), nat), dtype=torch.long)
    for i, n in enumerate(names):
        numbers[i, : len(mols[n]["numbers"])] = torch.tensor(mols[n]["numbers"])
    d = BatchDescriptor(numbers, torch.device("cpu"))
    t = d._t
    for i, n in enumerate(names):
        m = O.make_mol(mols[n]["numbers"])
        s0, s1, o0, o1 = d.sh_off[i], d.sh_off[i + 1], d.ao_off[i], d.ao_off[i + 1]
        assert (t["sh_atom"][s0:s1].numpy() == m.sh_atom).all()
        assert (t["sh_l"][s0:s1].numpy() == m.sh_l).all()
        assert (t["sh_ao"][s0:s1].numpy() == m.sh_ao).all()
        assert (t["ao_sh"][o0:o1].numpy() == m.ao_sh).all()
        by_l = t["sh_by_l"][s0:s1].numpy()
        assert sorted(by_l.tolist()) == list(range(m.nsh))
        assert (np.diff(m.sh_l[by_l]) >= 0).all()
        assert d.mat_off[i + 1] - d.mat_off[i] == m.nao**2
    assert d.struct.nao_max == max(O.make_mol(mols[n]["numbers"]).nao for n in names)


def test_cgto_tables_match_oracle():
    par, opar = gfn1_param(), O.params()
    for z in (1, 6, 7, 8, 16, 17, 35, 79):
        for k in range(int(par.nshell[z])):
            a1, c1 = par.cgto(z, k)
            a2, c2 = opar.cgto(z, k)
            assert np.array_equal(a1, a2) and np.allclose(c1, c2, rtol=0, atol=1e-15)
    # H 2s is orthogonal to 1s and normalised (basis/ortho.py)
    a1, c1 = par.cgto(1, 0)
    a2, c2 = par.cgto(1, 1)
    s12 = (np.sqrt(np.pi / (a1[:, None] + a2[None, :])) ** 3 * c1[:, None] * c2[None, :]).sum()
    s22 = (np.sqrt(np.pi / (a2[:, None] + a2[None, :])) ** 3 * c2[:, None] * c2[None, :]).sum()
    assert abs(s12) < 1e-8 and abs(s22 - 1) < 1e-12  # 1s STO-4G is normalised to ~1e-9 only


def test_hscale_table_matches_oracle():
    par, opar = gfn1_param(), O.params()
    tab = par.hscale_table()
    for t1 in range(6):
        for t2 in range(6):
            assert tab[t1, t2] == opar.hscale(t1 % 3, t1 < 3, t2 % 3, t2 < 3)


def test_calculator_rejects_what_the_path_does_not_cover():
    from dxtb_b200 import GFN1Calculator
    from dxtb_b200.exceptions import DtypeError, MissingD3ReferenceError

    numbers = torch.tensor([3, 1])
    with pytest.raises(DtypeError):
        GFN1Calculator(numbers.to(torch.float32), device="cuda")
    with pytest.raises(NotImplementedError):
        GFN1Calculator(numbers, device="cpu", opts={"exclude": ["disp"]})  # no CPU fallback
    with pytest.raises(NotImplementedError):
        GFN1Calculator(numbers, device="cuda", dtype=torch.float32)
    if torch.cuda.is_available():
        with pytest.raises(MissingD3ReferenceError):
            GFN1Calculator(numbers, device="cuda")
        with pytest.raises(NotImplementedError):
            GFN1Calculator(numbers, device="cuda", opts={"exclude": ["disp"], "scf_mode": "implicit"})


def test_product_does_not_import_oracle():
    for p in (ROOT / "dxtb_b200").rglob("*.py"):
        assert "oracle" not in p.read_text().replace("the oracle", ""), p
