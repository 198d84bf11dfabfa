"""Device-resident, immutable batch descriptor (ragged / CSR layout).

Replaces dxtb's ``IndexHelper`` (``basis/indexhelper.py:336-491``) and the per-species parameter gathers of
``BaseHamiltonian.__init__`` (``xtb/base.py:107-182``), ``ES2``/``ES3``/``Repulsion``/``Halogen`` caches for the
hot path: everything the kernels need about *which* atoms / shells / orbitals exist is computed once on
the host with vectorised NumPy and uploaded.  Zero padding of ``numbers`` is removed here: kernels never see
padding.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi
from .param import MAX_PRIM, GFN1Param, gfn1_param

INT_CUTOFF = 50.0  # constants/defaults.py:93
REP_CUTOFF = 25.0  # constants/xtb.py:33
XB_CUTOFF = 20.0  # constants/xtb.py:37
CN_CUTOFF = 25.0  # tad-mctc ncoord default
KCN_D3 = 16.0
D3_DISP_CUTOFF = 50.0  # tad-dftd3 defaults.D3_DISP_CUTOFF
D3_WF = 4.0  # tad-dftd3 Gaussian weighting factor


def _excl_cumsum(x: np.ndarray, dtype=np.int64) -> np.ndarray:
    out = np.zeros(x.size + 1, dtype=dtype)
    np.cumsum(x, out=out[1:])
    return out


class BatchDescriptor:
    """CSR description of a batch of molecules + all per-atom/per-shell GFN1 parameters on ``device``."""

    def __init__(self, numbers: torch.Tensor, device: torch.device, par: GFN1Param | None = None,
                 exclude: tuple[str, ...] = (), int_cutoff: float = INT_CUTOFF, d3_table: dict | None = None):
        par = par or gfn1_param()
        self.par = par
        self.device = device
        num = numbers.detach().cpu().numpy().astype(np.int64)
        self.single = num.ndim == 1
        if self.single:
            num = num[None, :]
        if num.ndim != 2:
            raise ValueError("numbers must have shape (nat,) or (nb, nat)")
        if (num < 0).any() or (num > 86).any():
            raise ValueError("atomic numbers must be in 0..86 (0 = padding)")
        self.nb, self.nat_pad = num.shape
        mask = num > 0
        nat = mask.sum(1).astype(np.int64)
        if (nat == 0).any():
            raise ValueError("empty molecule in batch")
        z = num[mask]  # (nat_tot,) row-major order keeps molecules contiguous
        self.mask_np = mask
        at_mol = np.repeat(np.arange(self.nb), nat)
        at_off = _excl_cumsum(nat)

        at_nsh = par.nshell[z].astype(np.int64)
        nsh_tot = int(at_nsh.sum())
        sh_atom_g = np.repeat(np.arange(z.size), at_nsh)
        at_sh0_g = _excl_cumsum(at_nsh)
        sh_k = np.arange(nsh_tot) - at_sh0_g[sh_atom_g]
        sh_z = z[sh_atom_g]
        sh_l = par.ang[sh_z, sh_k].astype(np.int64)
        sh_mol = at_mol[sh_atom_g]
        nsh = np.bincount(sh_mol, minlength=self.nb)
        sh_off = _excl_cumsum(nsh)
        sh_nao = 2 * sh_l + 1
        sh_ao_g = _excl_cumsum(sh_nao)
        nao = np.bincount(sh_mol, weights=sh_nao, minlength=self.nb).astype(np.int64)
        ao_off = _excl_cumsum(nao)
        nao_tot = int(nao.sum())
        ao_sh_g = np.repeat(np.arange(nsh_tot), sh_nao)
        ao_mol = sh_mol[ao_sh_g]

        self.nat, self.nsh, self.nao = nat, nsh, nao
        self.at_off, self.sh_off, self.ao_off = at_off, sh_off, ao_off
        self.mat_off = _excl_cumsum(nao * nao)
        self.gam_off = _excl_cumsum(nsh * nsh)
        self.eeq_off = _excl_cumsum((nat + 1) * (nat + 1))
        self.eeq_large = [int(i) for i in np.flatnonzero(nat >= 256)]  # XTB_EEQ_LARGE_NAT (include/xtb_b200.h)
        self.nat_tot, self.nsh_tot, self.nao_tot = int(z.size), nsh_tot, nao_tot
        self.z = z

        # molecule-local ids
        sh_atom = sh_atom_g - at_off[sh_mol]
        sh_ao = sh_ao_g[:-1] - ao_off[sh_mol]
        ao_sh = ao_sh_g - sh_off[ao_mol]
        at_sh0 = at_sh0_g[:-1] - sh_off[at_mol]
        order = np.lexsort((np.arange(nsh_tot), sh_l, sh_mol))
        sh_by_l = order - sh_off[sh_mol[order]]
        nsh_l = np.bincount(sh_mol * 3 + sh_l, minlength=3 * self.nb).reshape(self.nb, 3)

        # CGTO table over the unique (element, shell) pairs present
        key = sh_z * 4 + sh_k
        ukey, inv = np.unique(key, return_inverse=True)
        cg = np.zeros((ukey.size, _abi.CGTO))
        for r, kk in enumerate(ukey):
            alpha, coeff = par.cgto(int(kk // 4), int(kk % 4))
            cg[r, 0] = alpha.size
            cg[r, 1 : 1 + alpha.size] = alpha
            cg[r, 1 + MAX_PRIM : 1 + MAX_PRIM + alpha.size] = coeff
        sh_type = sh_l + 3 * (~par.valence[sh_z, sh_k]).astype(np.int64)

        species, at_species = np.unique(z, return_inverse=True)
        kpair = par.kpair[np.ix_(species, species)]

        at_par = np.zeros((z.size, _abi.ATPAR))
        at_par[:, _abi.AT_RAD] = par.atomic_rad[z]
        at_par[:, _abi.AT_RCOV] = par.cov_d3[z]
        at_par[:, _abi.AT_EN] = par.en[z]
        at_par[:, _abi.AT_AREP] = par.arep[z]
        at_par[:, _abi.AT_ZEFF] = 0.0 if "rep" in exclude else par.zeff[z]  # excluded repulsion: zero energy and gradient
        at_par[:, _abi.AT_GAM3] = 0.0 if "es3" in exclude else par.gam3[z]
        at_par[:, _abi.AT_XBOND] = 0.0 if "hal" in exclude else par.xbond[z]
        at_par[:, _abi.AT_EEQ_CHI] = par.eeq_chi[z]
        at_par[:, _abi.AT_EEQ_ETA] = par.eeq_eta[z]
        at_par[:, _abi.AT_EEQ_KCN] = par.eeq_kcn[z]
        at_par[:, _abi.AT_EEQ_RAD] = par.eeq_rad[z]
        if d3_table is not None:
            at_par[:, _abi.AT_R4R2] = np.asarray(d3_table["r4r2"], dtype=np.float64)[z]
        sh_par = np.zeros((nsh_tot, _abi.SHPAR))
        sh_par[:, _abi.SH_LEVEL] = par.level[sh_z, sh_k]
        sh_par[:, _abi.SH_KCN] = par.kcn[sh_z, sh_k]
        sh_par[:, _abi.SH_SHPOLY] = par.shpoly[sh_z, sh_k]
        sh_par[:, _abi.SH_ETA] = par.eta[sh_z, sh_k]
        sh_par[:, _abi.SH_REFOCC] = par.refocc[sh_z, sh_k]

        # sum of reference occupations per molecule (scf/iterator.py:172-173)
        self.nel0 = np.bincount(sh_mol, weights=par.refocc[sh_z, sh_k], minlength=self.nb)

        def dev(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a.astype(dt))).to(device)

        i32, i64, f64 = np.int32, np.int64, np.float64
        self._t = dict(
            at_off=dev(at_off, i32), sh_off=dev(sh_off, i32), ao_off=dev(ao_off, i32),
            mat_off=dev(self.mat_off, i64), gam_off=dev(self.gam_off, i64), eeq_off=dev(self.eeq_off, i64),
            at_z=dev(z, i32), at_species=dev(at_species, i32), at_sh0=dev(at_sh0, i32), at_nsh=dev(at_nsh, i32),
            at_par=dev(at_par, f64),
            sh_atom=dev(sh_atom, i32), sh_l=dev(sh_l, i32), sh_ao=dev(sh_ao, i32), sh_cgto=dev(inv, i32),
            sh_type=dev(sh_type, i32), sh_by_l=dev(sh_by_l, i32), nsh_l=dev(nsh_l, i32), sh_par=dev(sh_par, f64),
            ao_sh=dev(ao_sh, i32), cgto=dev(cg, f64), kpair=dev(kpair, f64),
        )
        if d3_table is not None:  # per-species slices of the D3 reference data
            refcn = np.asarray(d3_table["cn"], dtype=np.float64)[species]
            c6 = np.asarray(d3_table["c6"], dtype=np.float64)[np.ix_(species, species)]
            if refcn.shape[1] != 7 or c6.shape[2:] != (7, 7):
                raise ValueError("D3 reference table must have 7 references per element")
            self._t["d3_refcn"] = dev(refcn, f64)
            self._t["d3_c6"] = dev(c6, f64)
        # scatter/gather maps between the padded (nb, nat_pad) layout and the ragged one
        self.atom_index = torch.from_numpy(np.flatnonzero(mask.reshape(-1))).to(device)  # ragged -> flat padded
        self.at_mol = dev(at_mol, i64)
        self.ao_mol = dev(ao_mol, i64)
        ao_local = np.arange(nao_tot) - ao_off[ao_mol]
        self.nao_pad = int(nao.max())
        self.ao_index = dev(ao_mol * self.nao_pad + ao_local, i64)
        self._ao_atom_local = (sh_atom[ao_sh_g]).astype(np.int64)  # molecule-local atom of every AO (ragged order)

        s = _abi.XtbBatch()
        s.nb, s.nat_tot, s.nsh_tot, s.nao_tot = self.nb, self.nat_tot, nsh_tot, nao_tot
        s.nat_max, s.nsh_max, s.nao_max = int(nat.max()), int(nsh.max()), int(nao.max())
        s.nspecies, s.ncgto = int(species.size), int(ukey.size)
        s.has_xb = int(bool((at_par[:, _abi.AT_XBOND] != 0.0).any()))
        s.nsh_l_max = (C.c_int32 * 4)(*[int(x) for x in nsh_l.max(axis=0)], 0)
        s.mat_total, s.gam_total, s.eeq_total = int(self.mat_off[-1]), int(self.gam_off[-1]), int(self.eeq_off[-1])
        for name, t in self._t.items():
            setattr(s, name, t.data_ptr())
        s.hscale = (C.c_double * 36)(*par.hscale_table().reshape(-1).tolist())
        s.enscale, s.rep_kexp, s.xb_damp, s.xb_rscale, s.gexp = par.enscale, par.rep_kexp, par.xb_damp, par.xb_rscale, par.gexp
        s.int_cutoff, s.rep_cutoff, s.xb_cutoff, s.cn_cutoff, s.kcn_d3 = int_cutoff, REP_CUTOFF, XB_CUTOFF, CN_CUTOFF, KCN_D3
        s.d3_s6, s.d3_s8, s.d3_a1, s.d3_a2 = par.d3["s6"], par.d3["s8"], par.d3["a1"], par.d3["a2"]
        s.d3_cutoff, s.d3_wf = D3_DISP_CUTOFF, D3_WF
        self.has_d3 = d3_table is not None
        self.struct = s

    @property
    def ptr(self):
        return C.addressof(self.struct)

    # ---- layout conversion helpers ---------------------------------------------------------------
    def gather_atoms(self, padded: torch.Tensor) -> torch.Tensor:
        """(nb, nat_pad, ...) -> (nat_tot, ...)"""
        tail = padded.shape[2:] if not self.single else padded.shape[1:]
        flat = padded.reshape(self.nb * self.nat_pad, *tail)
        return flat.index_select(0, self.atom_index).contiguous()

    def scatter_atoms(self, ragged: torch.Tensor) -> torch.Tensor:
        """(nat_tot, ...) -> (nb, nat_pad, ...) zero padded (or (nat, ...) for a single molecule)."""
        out = ragged.new_zeros((self.nb * self.nat_pad, *ragged.shape[1:]))
        out.index_copy_(0, self.atom_index, ragged)
        out = out.reshape(self.nb, self.nat_pad, *ragged.shape[1:])
        return out[0] if self.single else out

    def scatter_matrices(self, ragged: torch.Tensor) -> torch.Tensor:
        """(mat_total,) ragged nao x nao blocks -> (nb, nao_pad, nao_pad) zero padded (built lazily; not on the hot path)."""
        if not hasattr(self, "_mat_index"):
            nao = torch.from_numpy(self.nao).to(self.device)
            mol = torch.repeat_interleave(torch.arange(self.nb, device=self.device), nao * nao)
            local = torch.arange(int(self.mat_off[-1]), device=self.device) - torch.from_numpy(self.mat_off[:-1]).to(self.device)[mol]
            n = nao[mol]
            self._mat_index = mol * self.nao_pad * self.nao_pad + (local // n) * self.nao_pad + local % n
        out = ragged.new_zeros((self.nb * self.nao_pad * self.nao_pad,))
        out.index_copy_(0, self._mat_index, ragged)
        out = out.reshape(self.nb, self.nao_pad, self.nao_pad)
        return out[0] if self.single else out

    def scatter_orbitals(self, ragged: torch.Tensor) -> torch.Tensor:
        out = ragged.new_zeros((self.nb * self.nao_pad,))
        out.index_copy_(0, self.ao_index, ragged)
        out = out.reshape(self.nb, self.nao_pad)
        return out[0] if self.single else out
