"""GFN1-xTB parametrisation tables for the host-side descriptor builder.

Mirrors what dxtb's ``ParamModule`` getters provide to the hot path (reference:
``param/module/utils.py:323-573``, ``basis/slater.py:69-135``, ``basis/ortho.py:36-110``,
``xtb/gfn1.py:66-165``), but as dense per-element NumPy tables indexed by atomic number so that a
whole batch can be gathered with vectorised indexing.
"""
from __future__ import annotations

import json
import math
from functools import lru_cache
from pathlib import Path

import numpy as np

DATA = Path(__file__).resolve().parent / "data" / "gfn1_param.json"
MAX_Z = 86
MAX_SHELL = 3
MAX_PRIM = 7


class GFN1Param:
    """Dense element tables of the GFN1-xTB parameter set (``param/gfn1/gfn1-xtb.toml``)."""

    def __init__(self, path: str | Path = DATA):
        raw = json.loads(Path(path).read_text())
        self.meta = raw["meta"]
        tp = raw["third_party"]
        c = tp["codata2018"]
        self.AA2AU = 1.0 / (c["bohr_m"] * 1e10)
        self.EV2AU = c["ev_j"] / c["hartree_j"]
        self.KELVIN2AU = c["kb_j_per_k"] / c["hartree_j"]

        z1 = MAX_Z + 1
        self.nshell = np.zeros(z1, dtype=np.int32)
        self.ang = np.zeros((z1, MAX_SHELL), dtype=np.int32)
        self.pqn = np.zeros((z1, MAX_SHELL), dtype=np.int32)
        self.valence = np.zeros((z1, MAX_SHELL), dtype=bool)
        self.level = np.zeros((z1, MAX_SHELL))  # Hartree
        self.kcn = np.zeros((z1, MAX_SHELL))  # Hartree
        self.shpoly = np.zeros((z1, MAX_SHELL))
        self.refocc = np.zeros((z1, MAX_SHELL))
        self.eta = np.zeros((z1, MAX_SHELL))  # gam * lgam
        self.slater = np.zeros((z1, MAX_SHELL))
        self.ngauss = np.zeros((z1, MAX_SHELL), dtype=np.int32)
        self.gam3 = np.zeros(z1)
        self.zeff = np.zeros(z1)
        self.arep = np.zeros(z1)
        self.xbond = np.zeros(z1)
        self.en = np.zeros(z1)
        for zs, e in raw["element"].items():
            z = int(zs)
            n = len(e["ang"])
            self.nshell[z] = n
            self.ang[z, :n] = e["ang"]
            self.pqn[z, :n] = e["pqn"]
            seen = set()
            for k, l in enumerate(e["ang"]):  # param/module/utils.py:488-542
                self.valence[z, k] = l not in seen
                seen.add(l)
            self.level[z, :n] = np.array(e["levels_ev"]) * self.EV2AU  # xtb/base.py:154
            self.kcn[z, :n] = np.array(e["kcn_ev"]) * self.EV2AU  # xtb/base.py:155
            self.shpoly[z, :n] = e["shpoly"]
            self.refocc[z, :n] = e["refocc"]
            self.eta[z, :n] = np.array(e["lgam"]) * e["gam"]  # secondorder.py:838-839
            self.slater[z, :n] = e["slater"]
            self.ngauss[z, :n] = e["ngauss"]
            self.gam3[z] = e["gam3"]
            self.zeff[z] = e["zeff"]
            self.arep[z] = e["arep"]
            self.xbond[z] = e["xbond"]
            self.en[z] = e["en"]
        self.atomic_rad = np.concatenate([[0.0], np.array(tp["atomic_radii_angstrom"]) * self.AA2AU])
        self.cov_d3 = np.concatenate([[0.0], np.array(tp["cov_2009_angstrom"]) * self.AA2AU * 4.0 / 3.0])
        q = tp["eeq2019"]
        self.eeq_chi = np.concatenate([[0.0], q["chi"]])
        self.eeq_eta = np.concatenate([[0.0], q["eta"]])
        self.eeq_kcn = np.concatenate([[0.0], q["kcn"]])
        self.eeq_rad = np.concatenate([[0.0], q["rad"]])

        h = raw["hamiltonian"]
        self.kpol = h["kpol"]
        self.enscale = h["enscale"]
        self.kshell = h["shell"]
        self.kpair = np.ones((z1, z1))  # param/module/utils.py:442-486
        for a, b, v in h["kpair"]:
            self.kpair[a, b] = v
            self.kpair[b, a] = v
        self.rep_kexp = raw["repulsion"]["kexp"]
        self.xb_damp = raw["halogen"]["damping"]
        self.xb_rscale = raw["halogen"]["rscale"]
        self.gexp = raw["charge"]["gexp"]
        self.d3 = raw["dispersion_d3"]
        self._sto = {int(n): (np.array(v["coeff"]), np.array(v["alpha"])) for n, v in raw["sto_ng"].items()}

    def with_ev2au(self, ev2au: float) -> "GFN1Param":
        """Copy with another eV->Hartree factor (tblite's goldens use 1 Eh = 27.21138505 eV)."""
        import copy

        new = copy.copy(self)
        new.level = self.level * (ev2au / self.EV2AU)
        new.kcn = self.kcn * (ev2au / self.EV2AU)
        new.EV2AU = ev2au
        return new

    # ------------------------------------------------------------------------------------------
    def hscale_table(self) -> np.ndarray:
        """6x6 table over shell type = l + 3*(non-valence) (xtb/gfn1.py:66-165)."""
        lab = "spd"
        out = np.zeros((6, 6))
        for t1 in range(6):
            for t2 in range(6):
                l1, v1, l2, v2 = t1 % 3, t1 < 3, t2 % 3, t2 < 3
                k11 = self.kshell.get(lab[l1] * 2, 1.0) if v1 else self.kpol
                k22 = self.kshell.get(lab[l2] * 2, 1.0) if v2 else self.kpol
                val = (k11 + k22) / 2.0
                if v1 and v2:
                    val = self.kshell.get(lab[l1] + lab[l2], self.kshell.get(lab[l2] + lab[l1], val))
                out[t1, t2] = val
        return out

    def _slater_to_gauss(self, ng: int, n: int, l: int, zeta: float):
        """STO-NG expansion with normalisation (basis/slater.py:69-135)."""
        itype = n + (0, 4, 7, 9, 10)[l] - 1
        if n == 6 and ng == 6:
            itype = 15 + l
        ctab, atab = self._sto[ng]
        alpha = atab[itype] * (zeta * zeta)
        dfact = (1.0, 1.0, 3.0, 15.0, 105.0)[l]
        # basis/slater.py:52-54, 130-134: the reference keeps its double-factorial table in float32, so sqrt((2l-1)!!) is
        # taken in float32 before it divides the fp64 coefficients (d shells: sqrt(3) -> 1.7320507764816284).  Together with
        # the float32 cartesian->spherical matrix (md/trafo.py:37-38, 69-75; XTB_TRAFO_S3 in xtb_integrals.cu) the net effect
        # is a relative +1.8e-8 on the d(z2) component only: up to 9e-9 on S elements, 5e-10 Eh on MB16_43_01.  Replicated
        # for parity with dxtb's fp64 path.
        coeff = ctab[itype] * (2.0 / math.pi * alpha) ** 0.75 * np.sqrt(4.0 * alpha) ** l / float(np.sqrt(np.float32(dfact)))
        return alpha, coeff

    def cgto(self, z: int, k: int) -> tuple[np.ndarray, np.ndarray]:
        """Primitive exponents / contraction coefficients of shell ``k`` of element ``z``; a non-valence
        shell (H 2s) is orthonormalised against the preceding one (basis/bas.py:149-193, basis/ortho.py:78-110)."""
        cache = self.__dict__.setdefault("_cgto_cache", {})
        if (z, k) in cache:
            return cache[(z, k)]
        alpha, coeff = self._slater_to_gauss(int(self.ngauss[z, k]), int(self.pqn[z, k]), int(self.ang[z, k]), float(self.slater[z, k]))
        if not self.valence[z, k]:
            ai, ci = self.cgto(z, k - 1)

            def sint(a1, a2, c1, c2):
                o = 1.0 / (a1[:, None] + a2[None, :])  # basis/ortho.py:57-59
                return float((np.sqrt(math.pi * o) ** 3 * c1[:, None] * c2[None, :]).sum())

            ovl = sint(ai, alpha, ci, coeff)
            alpha = np.concatenate([alpha, ai])
            coeff = np.concatenate([coeff, -ovl * ci])
            coeff = coeff / math.sqrt(sint(alpha, alpha, coeff, coeff))
        if alpha.size > MAX_PRIM:
            raise NotImplementedError("more than 7 primitives per shell")
        cache[(z, k)] = (alpha, coeff)
        return alpha, coeff


@lru_cache(maxsize=1)
def gfn1_param() -> GFN1Param:
    return GFN1Param()
