"""CPU oracle: a NumPy fp64 restatement of dxtb's GFN1-xTB single-point path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dxtb_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs do, and only as the checker / CPU baseline.

Every function cites the reference file:line (relative to ``/root/reference/src/dxtb/_src``,
dxtb v0.4.0) whose arithmetic it restates.  Arithmetic that lives in third-party packages that
are absent from the reference tree (tad-mctc 0.7.0: ``cdist``, ``cn_d3``/``exp_count``,
``storch.eighb``, radii, unit constants; tad-multicharge 0.5.0: EEQ charges; tad-dftd3 0.6.0: D3)
is restated from the published algorithms; see DESIGN.md "Oracle pinning" for what pins each.

Parity status: PINNED for overlap / H0 / CN-derivatives (float32 goldens of the reference's own
tests), SCF and total energies (tblite fp64 literals of the reference's tests, where dispersion is
zero or excluded), EEQ guess for H/C (reference KAT).  UNPINNED: D3 dispersion (reference C6
table is third-party data that is not available offline), EEQ parameters of elements other than
H and C, SCF iteration counts (no reference test asserts them).
"""
from __future__ import annotations

import json
import math
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np
from scipy.linalg import eigh as _scipy_eigh
from scipy.special import erf as _erf

_DATA = Path(__file__).resolve().parent.parent / "dxtb_b200" / "data" / "gfn1_param.json"

EPS = float(np.finfo(np.float64).eps)
TINY = float(np.finfo(np.float64).tiny)

# defaults: constants/defaults.py:93 (INTCUTOFF), constants/xtb.py:33-37, tad-mctc CN cutoff
INT_CUTOFF = 50.0
REP_CUTOFF = 25.0
XB_CUTOFF = 20.0
CN_CUTOFF = 25.0
KCN_D3 = 16.0
EEQ_KCN = 7.5
EEQ_CN_MAX = 8.0
EEQ_CN_CUTOFF = 25.0

SQRTPI3 = math.sqrt(math.pi) ** 3

# integral/driver/pytorch/impls/md/trafo.py:144-161 (NLM_CART) and :63-79 (TRAFO)
NLM_CART = (
    np.array([[0, 0, 0]]),
    np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0]]),  # py, pz, px
    np.array([[2, 0, 0], [0, 2, 0], [0, 0, 2], [1, 1, 0], [1, 0, 1], [0, 1, 1]]),
)
# trafo.py:37-38, 65-75: TRAFO is built with torch.tensor(...) at the default dtype, i.e. sqrt(3) and sqrt(3)/2 are rounded
# to float32 before they are cast to the working precision
_S3 = float(np.float32(math.sqrt(3.0)))
_S3_4 = float(np.float32(math.sqrt(3.0) * 0.5))
TRAFO = (
    np.array([[1.0]]),
    np.eye(3),
    np.array(
        [
            [-0.5, -0.5, 1.0, 0.0, 0.0, 0.0],
            [0.0, 0.0, 0.0, 0.0, _S3, 0.0],
            [0.0, 0.0, 0.0, 0.0, 0.0, _S3],
            [_S3_4, -_S3_4, 0.0, 0.0, 0.0, 0.0],
            [0.0, 0.0, 0.0, _S3, 0.0, 0.0],
        ]
    ),
)


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------
class Params:
    """GFN1-xTB parameter blob (param/gfn1/gfn1-xtb.toml restated as JSON by tools/make_param_blob.py)."""

    def __init__(self, path: Path | str = _DATA):
        d = json.loads(Path(path).read_text())
        self.raw = d
        tp = d["third_party"]
        c = tp["codata2018"]
        self.aa2au = 1.0 / (c["bohr_m"] * 1e10)
        self.ev2au = c["ev_j"] / c["hartree_j"]
        self.kelvin2au = c["kb_j_per_k"] / c["hartree_j"]
        # index by atomic number (0 = padding)
        self.atomic_rad = np.array([0.0] + tp["atomic_radii_angstrom"]) * self.aa2au  # xtb/base.py:138
        self.cov_d3 = np.array([0.0] + tp["cov_2009_angstrom"]) * self.aa2au * 4.0 / 3.0
        e = tp["eeq2019"]
        self.eeq_chi = np.array([0.0] + e["chi"])
        self.eeq_eta = np.array([0.0] + e["eta"])
        self.eeq_kcn = np.array([0.0] + e["kcn"])
        self.eeq_rad = np.array([0.0] + e["rad"])
        self.elem = {int(z): v for z, v in d["element"].items()}
        h = d["hamiltonian"]
        self.kpol = h["kpol"]
        self.enscale = h["enscale"]
        self.kshell = h["shell"]
        self.kpair = np.ones((87, 87))
        for a, b, v in h["kpair"]:
            self.kpair[a, b] = v
            self.kpair[b, a] = v
        self.kexp = d["repulsion"]["kexp"]
        self.xb_damp = d["halogen"]["damping"]
        self.xb_rscale = d["halogen"]["rscale"]
        self.gexp = d["charge"]["gexp"]
        self.d3 = d["dispersion_d3"]
        self.sto = {int(n): (np.array(v["coeff"]), np.array(v["alpha"])) for n, v in d["sto_ng"].items()}
        self._cgto_cache: dict[tuple[int, int], tuple[np.ndarray, np.ndarray]] = {}

    # basis/slater.py:69-135
    def slater_to_gauss(self, ng: int, n: int, l: int, zeta: float):
        itype = n + [0, 4, 7, 9, 10][l] - 1
        if n == 6 and ng == 6:
            itype = 15 + l
        coeff_t, alpha_t = self.sto[ng]
        alpha = alpha_t[itype] * zeta**2
        dfact = [1.0, 1.0, 3.0, 15.0, 105.0][l]
        # slater.py:52-54: `dfactorial` is a float32 tensor, so its square root is taken in float32 (l = 2: 1.7320507764816284)
        coeff = coeff_t[itype] * ((2.0 / math.pi * alpha) ** 0.75 * np.sqrt(4 * alpha) ** l / float(np.sqrt(np.float32(dfact))))
        return alpha, coeff

    # basis/ortho.py:36-110 and the trigger in basis/bas.py:182-188
    def cgto(self, z: int, ish: int):
        key = (z, ish)
        if key in self._cgto_cache:
            return self._cgto_cache[key]
        e = self.elem[z]
        alpha, coeff = self.slater_to_gauss(e["ngauss"][ish], e["pqn"][ish], e["ang"][ish], e["slater"][ish])
        if not self.valence(z)[ish]:
            ai, ci = self.cgto(z, ish - 1)
            aj, cj = alpha, coeff

            def gint(a1, a2, c1, c2):
                o = 1.0 / (a1[:, None] + a2[None, :])
                return float((np.sqrt(math.pi * o) ** 3 * c1[:, None] * c2[None, :]).sum())

            ovl = gint(ai, aj, ci, cj)
            alpha = np.concatenate([aj, ai])
            coeff = np.concatenate([cj, -ovl * ci])
            coeff = coeff / math.sqrt(gint(alpha, alpha, coeff, coeff))
        self._cgto_cache[key] = (alpha, coeff)
        return alpha, coeff

    # param/module/utils.py:488-542: first shell of each angular momentum is valence
    def valence(self, z: int) -> list[bool]:
        seen, out = set(), []
        for l in self.elem[z]["ang"]:
            out.append(l not in seen)
            seen.add(l)
        return out

    # xtb/gfn1.py:66-165
    def hscale(self, l1: int, v1: bool, l2: int, v2: bool) -> float:
        lab = "spdfg"
        k11 = self.kpol if not v1 else self.kshell.get(lab[l1] * 2, 1.0)
        k22 = self.kpol if not v2 else self.kshell.get(lab[l2] * 2, 1.0)
        if v1 and v2:
            a, b = lab[l1] + lab[l2], lab[l2] + lab[l1]
            if a in self.kshell:
                return self.kshell[a]
            if b in self.kshell:
                return self.kshell[b]
        return (k11 + k22) / 2.0


_PARAMS: Params | None = None


def params() -> Params:
    global _PARAMS
    if _PARAMS is None:
        _PARAMS = Params()
    return _PARAMS


# --------------------------------------------------------------------------------------
# index maps (basis/indexhelper.py:336-491)
# --------------------------------------------------------------------------------------
@dataclass
class Mol:
    numbers: np.ndarray  # (nat,) int, no padding
    sh_atom: np.ndarray = field(default=None)  # shell -> atom
    sh_l: np.ndarray = field(default=None)
    sh_idx: np.ndarray = field(default=None)  # shell index within its element
    sh_ao: np.ndarray = field(default=None)  # first AO of shell
    ao_sh: np.ndarray = field(default=None)  # AO -> shell
    ao_atom: np.ndarray = field(default=None)
    nat: int = 0
    nsh: int = 0
    nao: int = 0


def make_mol(numbers) -> Mol:
    par = params()
    numbers = np.asarray(numbers, dtype=np.int64)
    numbers = numbers[numbers > 0]
    m = Mol(numbers=numbers, nat=len(numbers))
    sh_atom, sh_l, sh_idx, sh_ao, ao_sh = [], [], [], [], []
    nao = 0
    for ia, z in enumerate(numbers):
        for ish, l in enumerate(par.elem[int(z)]["ang"]):
            sh_atom.append(ia)
            sh_l.append(l)
            sh_idx.append(ish)
            sh_ao.append(nao)
            ao_sh += [len(sh_atom) - 1] * (2 * l + 1)
            nao += 2 * l + 1
    m.sh_atom = np.array(sh_atom)
    m.sh_l = np.array(sh_l)
    m.sh_idx = np.array(sh_idx)
    m.sh_ao = np.array(sh_ao)
    m.ao_sh = np.array(ao_sh)
    m.ao_atom = m.sh_atom[m.ao_sh]
    m.nsh = len(sh_atom)
    m.nao = nao
    return m


def _shell_param(m: Mol, key: str) -> np.ndarray:
    par = params()
    return np.array([par.elem[int(m.numbers[a])][key][i] for a, i in zip(m.sh_atom, m.sh_idx)], dtype=np.float64)


def _cdist(pos: np.ndarray) -> np.ndarray:
    """tad-mctc ``storch.cdist`` (p=2): sqrt(clamp(sum((xi-xj)^2), min=eps))."""
    d = pos[:, None, :] - pos[None, :, :]
    return np.sqrt(np.maximum((d * d).sum(-1), EPS))


# --------------------------------------------------------------------------------------
# McMurchie-Davidson overlap (integral/driver/pytorch/impls/md/explicit.py:61-187, 190-319)
# --------------------------------------------------------------------------------------
def _ecoeffs(imax: int, jmax: int, x, rpi, rpj):
    """Hermite expansion coefficients E^{ij}_t by the MD recursion; the reference writes the same
    recursion out explicitly (explicit.py:360-422, 533-622, 830-937).  Returns dict[(i,j)] -> list over t."""
    one = np.ones_like(rpi)
    E = {(0, 0): [one]}

    def get(i, j, t):
        lst = E[(i, j)]
        return lst[t] if 0 <= t < len(lst) else 0.0

    for i in range(imax + 1):
        if i > 0:
            E[(i, 0)] = [
                (x * get(i - 1, 0, t - 1) if t > 0 else 0.0) + rpi * get(i - 1, 0, t) + (t + 1) * get(i - 1, 0, t + 1)
                for t in range(i + 1)
            ]
        for j in range(1, jmax + 1):
            E[(i, j)] = [
                (x * get(i, j - 1, t - 1) if t > 0 else 0.0) + rpj * get(i, j - 1, t) + (t + 1) * get(i, j - 1, t + 1)
                for t in range(i + j + 1)
            ]
    return E


def overlap_block(li, lj, ai, ci, aj, cj, vec, grad=False):
    """Spherical overlap block(s) for shell pairs of one (li,lj,cgto_i,cgto_j) class.

    ``vec`` (npair,3) is what the reference passes to ``md_explicit``: ``-(pos_i - pos_j)``
    (impls/overlap.py:225-226).  Returns S (npair, 2li+1, 2lj+1) and, if ``grad``, dS/dR_i
    (npair, 3, 2li+1, 2lj+1) (derivative w.r.t. the centre of shell i, explicit.py:196-203).
    """
    a = ai[None, :, None]
    b = aj[None, None, :]
    eij = a + b
    oij = 1.0 / eij
    xij = 0.5 * oij
    r2 = (vec * vec).sum(-1)[:, None, None]
    est = a * b * oij * r2
    sij = np.exp(-est) * SQRTPI3 * oij**1.5 * ci[None, :, None] * cj[None, None, :]
    v = vec[:, :, None, None]  # (npair,3,1,1)
    rpi = +v * (b * oij)[:, None]
    rpj = -v * (a * oij)[:, None]
    x = np.broadcast_to(xij[:, None], rpi.shape)
    E = _ecoeffs(li + (1 if grad else 0), lj, x, rpi, rpj)
    nlmi, nlmj = NLM_CART[li], NLM_CART[lj]
    npair = vec.shape[0]
    s3d = np.zeros((npair, len(nlmi), len(nlmj)))
    ds3d = np.zeros((npair, 3, len(nlmi), len(nlmj))) if grad else None
    for mi, (ix, iy, iz) in enumerate(nlmi):
        for mj, (jx, jy, jz) in enumerate(nlmj):
            ex = E[(ix, jx)][0][:, 0]
            ey = E[(iy, jy)][0][:, 1]
            ez = E[(iz, jz)][0][:, 2]
            s3d[:, mi, mj] = (sij * ex * ey * ez).sum((-2, -1))
            if grad:
                two_a = 2.0 * a

                def dcomp(i, j, k):
                    f = two_a * E[(i + 1, j)][0][:, k]
                    if i > 0:
                        f = f - i * E[(i - 1, j)][0][:, k]
                    return f

                ds3d[:, 0, mi, mj] = (sij * dcomp(ix, jx, 0) * ey * ez).sum((-2, -1))
                ds3d[:, 1, mi, mj] = (sij * ex * dcomp(iy, jy, 1) * ez).sum((-2, -1))
                ds3d[:, 2, mi, mj] = (sij * ex * ey * dcomp(iz, jz, 2)).sum((-2, -1))
    ti, tj = TRAFO[li], TRAFO[lj]
    s = np.einsum("ij,pjk,lk->pil", ti, s3d, tj)
    if not grad:
        return s, None
    ds = np.einsum("ij,pxjk,lk->pxil", ti, ds3d, tj)
    return s, ds


def _pair_classes(m: Mol, pos: np.ndarray, cutoff: float):
    """Lower-triangular off-atom shell pairs grouped by (element-shell i, element-shell j) class
    (basis/bas.py:195-278 + impls/overlap.py:188-233)."""
    dist = _cdist(pos)
    I, J = np.tril_indices(m.nsh, -1)
    A, B = m.sh_atom[I], m.sh_atom[J]
    d = dist[A, B]
    keep = (A != B) & (d < cutoff) & (d > 0.1)
    I, J, A, B = I[keep], J[keep], A[keep], B[keep]
    zi, zj = m.numbers[A], m.numbers[B]
    key = ((zi * 8 + m.sh_idx[I]) * 1024 + (zj * 8 + m.sh_idx[J])).astype(np.int64)
    out = []
    for k in np.unique(key):
        sel = key == k
        out.append((I[sel], J[sel]))
    return out


def overlap(m: Mol, pos: np.ndarray, cutoff: float = INT_CUTOFF, grad: bool = False):
    """impls/overlap.py:147-243 (and :246-349 for the gradient dS_{mu nu}/dR_{centre of mu})."""
    par = params()
    S = np.zeros((m.nao, m.nao))
    dS = np.zeros((3, m.nao, m.nao)) if grad else None
    for I, J in _pair_classes(m, pos, cutoff):
        i0, j0 = I[0], J[0]
        li, lj = int(m.sh_l[i0]), int(m.sh_l[j0])
        ai, ci = par.cgto(int(m.numbers[m.sh_atom[i0]]), int(m.sh_idx[i0]))
        aj, cj = par.cgto(int(m.numbers[m.sh_atom[j0]]), int(m.sh_idx[j0]))
        vec = pos[m.sh_atom[I]] - pos[m.sh_atom[J]]
        s, ds = overlap_block(li, lj, ai, ci, aj, cj, -vec, grad=grad)
        ni, nj = 2 * li + 1, 2 * lj + 1
        rows = (m.sh_ao[I][:, None] + np.arange(ni)[None, :])[:, :, None]
        cols = (m.sh_ao[J][:, None] + np.arange(nj)[None, :])[:, None, :]
        S[rows, cols] = s
        if grad:
            for x in range(3):
                dS[x][rows, cols] = ds[:, x]
    S = np.tril(S, -1) + np.triu(S.T)
    np.fill_diagonal(S, 1.0)
    if grad:
        dS = np.stack([np.tril(dS[x], -1) - np.triu(dS[x].T) for x in range(3)], axis=-1)  # (nao,nao,3)
    return S, dS


# --------------------------------------------------------------------------------------
# coordination number (tad-mctc ncoord.cn_d3 with exp_count; called at xtb/gfn1.py:57-64)
# --------------------------------------------------------------------------------------
def cn_d3(m: Mol, pos: np.ndarray, grad: bool = False):
    par = params()
    dist = _cdist(pos)
    rc = par.cov_d3[m.numbers]
    r0 = rc[:, None] + rc[None, :]
    mask = ~np.eye(m.nat, dtype=bool) & (dist <= CN_CUTOFF)
    dsafe = np.where(mask, dist, 1.0)
    cf = np.where(mask, 1.0 / (1.0 + np.exp(-KCN_D3 * (r0 / dsafe - 1.0))), 0.0)
    cn = cf.sum(-1)
    if not grad:
        return cn, None
    # d cf/dR = -k r0/R^2 * exp(..)/(1+exp(..))^2
    ex = np.exp(-KCN_D3 * (r0 / dsafe - 1.0))
    dcf = np.where(mask, -KCN_D3 * r0 / dsafe**2 * ex / (1.0 + ex) ** 2, 0.0)
    rij = pos[:, None, :] - pos[None, :, :]
    dcfdr = (dcf / dsafe)[:, :, None] * rij  # d cf_AB / d R_A
    return cn, dcfdr


# --------------------------------------------------------------------------------------
# H0 (xtb/base.py:251-360, xtb/gfn1.py:66-183)
# --------------------------------------------------------------------------------------
def _h0_shell_factors(m: Mol, pos: np.ndarray, cn: np.ndarray):
    par = params()
    dist = _cdist(pos)
    z_sh = m.numbers[m.sh_atom]
    val = np.array([par.valence(int(z))[i] for z, i in zip(z_sh, m.sh_idx)])
    se = _shell_param(m, "levels_ev") * par.ev2au - _shell_param(m, "kcn_ev") * par.ev2au * cn[m.sh_atom]
    shpoly = _shell_param(m, "shpoly")
    rad = par.atomic_rad[m.numbers]
    offatom = m.sh_atom[:, None] != m.sh_atom[None, :]
    rr_at = np.where(~np.eye(m.nat, dtype=bool), np.sqrt(dist / (rad[:, None] + rad[None, :])), 0.0)
    rr = rr_at[m.sh_atom[:, None], m.sh_atom[None, :]]
    tmp_a = 1.0 + shpoly[:, None] * rr
    tmp_b = 1.0 + shpoly[None, :] * rr
    var_pi = tmp_a * tmp_b
    en = np.array([par.elem[int(z)]["en"] for z in z_sh])
    var_x = np.where(offatom, 1.0 + par.enscale * (en[:, None] - en[None, :]) ** 2, 0.0)
    hs = np.array([[par.hscale(int(l1), bool(v1), int(l2), bool(v2)) for l2, v2 in zip(m.sh_l, val)] for l1, v1 in zip(m.sh_l, val)])
    kp = par.kpair[z_sh[:, None], z_sh[None, :]]
    var_k = np.where(val[:, None] & val[None, :], hs * kp * var_x, hs)
    var_h = 0.5 * (se[:, None] + se[None, :])
    hsh = np.where(offatom, var_pi * var_k * var_h, var_h)
    return dict(hsh=hsh, var_pi=var_pi, var_k=var_k, tmp_a=tmp_a, tmp_b=tmp_b, shpoly=shpoly, rr=rr,
                offatom=offatom, dist=dist, kcn=_shell_param(m, "kcn_ev") * par.ev2au)


def h0(m: Mol, pos: np.ndarray, S: np.ndarray, cn: np.ndarray | None = None):
    if cn is None:
        cn, _ = cn_d3(m, pos)
    f = _h0_shell_factors(m, pos, cn)
    hcore = f["hsh"][m.ao_sh[:, None], m.ao_sh[None, :]] * S
    return 0.5 * (hcore + hcore.T)


# --------------------------------------------------------------------------------------
# classical terms
# --------------------------------------------------------------------------------------
def repulsion(m: Mol, pos: np.ndarray, grad: bool = False):
    """components/classicals/repulsion/base.py:172-243, 269-334, 337-406; rep.py:66-98."""
    par = params()
    arep = np.array([par.elem[int(z)]["arep"] for z in m.numbers])
    zeff = np.array([par.elem[int(z)]["zeff"] for z in m.numbers])
    mask = ~np.eye(m.nat, dtype=bool)
    a = np.where(mask, np.sqrt(arep[:, None] * arep[None, :] + TINY), 0.0)
    zz = zeff[:, None] * zeff[None, :] * mask
    k = par.kexp
    dist = np.where(mask, _cdist(pos), EPS)
    r1k = dist**k
    e = np.where(mask & (dist <= REP_CUTOFF), zz * np.exp(-a * r1k) / dist, 0.0)
    eat = 0.5 * e.sum(-1)
    if not grad:
        return eat, None
    g = np.where(mask, -(a * r1k * k + 1.0) * e / dist**2, 0.0)
    rij = pos[:, None, :] - pos[None, :, :]
    # E = 1/2 sum_AB e_AB ; dE/dR_A = sum_B de_AB/dR * rij/R
    return eat, (g[:, :, None] * rij).sum(1)


def halogen(m: Mol, pos: np.ndarray, grad: bool = False):
    """components/classicals/halogen/hal.py:209-364.  With ``grad`` also dE/dR of the triples (X, J, K): the
    reference differentiates the energy by autograd (classicals/base.py:118-156) with the neighbour list fixed
    (the index K of the halogen's nearest neighbour is integer data, hal.py:253-266), so the derivative acts on the
    three squared distances d2xj, d2xk, d2kj only."""
    par = params()
    halogens, bases = (17, 35, 53, 85), (7, 8, 15, 16)
    e = np.zeros(m.nat)
    g = np.zeros((m.nat, 3)) if grad else None
    rads = par.atomic_rad[m.numbers] * par.xb_rscale
    for x, zx in enumerate(m.numbers):
        if zx not in halogens:
            continue
        xb = par.elem[int(zx)]["xbond"]
        for j, zj in enumerate(m.numbers):
            if zj not in bases:
                continue
            if np.linalg.norm(pos[x] - pos[j]) > XB_CUTOFF:
                continue
            kbest, dbest = 0, np.finfo(np.float64).max
            for k in range(m.nat):
                r1 = np.linalg.norm(pos[x] - pos[k])
                if 0.0 < r1 < dbest:
                    kbest, dbest = k, r1
            r0 = rads[x] + rads[j]
            dxj, dxk, dkj = pos[j] - pos[x], pos[kbest] - pos[x], pos[kbest] - pos[j]
            d2xj, d2xk, d2kj = dxj @ dxj, dxk @ dxk, dkj @ dkj
            xy = math.sqrt(d2xk * d2xj)
            lj6 = (r0 / math.sqrt(d2xj)) ** 6
            lj12 = lj6**2
            lj = (lj12 - par.xb_damp * lj6) / (1.0 + lj12)
            cosa = (d2xk + d2xj - d2kj) / xy
            t = 0.5 - 0.25 * cosa
            fd = t**6
            e[x] += lj * fd * xb
            if grad:
                # chain rule over a = d2xj, b = d2xk, c = d2kj
                dlj_dlj6 = ((2.0 * lj6 - par.xb_damp) * (1.0 + lj12) - (lj12 - par.xb_damp * lj6) * 2.0 * lj6) / (1.0 + lj12) ** 2
                dlj_da = dlj_dlj6 * (-3.0 * lj6 / d2xj)
                dfd = -1.5 * t**5  # d fd / d cosa
                de_da = xb * (dlj_da * fd + lj * dfd * (1.0 / xy - 0.5 * cosa / d2xj))
                de_db = xb * lj * dfd * (1.0 / xy - 0.5 * cosa / d2xk)
                de_dc = xb * lj * dfd * (-1.0 / xy)
                g[j] += 2.0 * de_da * dxj - 2.0 * de_dc * dkj
                g[kbest] += 2.0 * de_db * dxk + 2.0 * de_dc * dkj
                g[x] += -2.0 * de_da * dxj - 2.0 * de_db * dxk
    return (e, g) if grad else e


# --------------------------------------------------------------------------------------
# D3(BJ) dispersion (tad-dftd3 0.6.0: dftd3 -> weight_references / atomic_c6 / dispersion with
# rational_damping; wrapper components/classicals/dispersion/d3.py:93-211).  The reference data
# (reference CNs, C6 table, sqrt(Z) r4/r2) are THIRD-PARTY DATA that is not available offline, so the
# table is an argument: dict(cn=(Z+1,7) with -1 for missing references, c6=(Z+1,Z+1,7,7), r4r2=(Z+1,)).
# --------------------------------------------------------------------------------------
D3_WF = 4.0
D3_DISP_CUTOFF = 50.0


def d3_weights(z, cn, table, grad=False):
    refcn = table["cn"][z]  # (nat, 7)
    mask = refcn >= 0
    dcn = np.where(mask, refcn - cn[:, None], 0.0)
    w = np.where(mask, np.exp(-D3_WF * dcn * dcn), 0.0)
    norm = w.sum(-1, keepdims=True) + EPS
    gw = w / norm
    if not grad:
        return gw, None
    dw = np.where(mask, 2.0 * D3_WF * dcn * w, 0.0)  # d w / d cn
    dgw = dw / norm - w * dw.sum(-1, keepdims=True) / norm**2
    return gw, dgw


def d3_dispersion(numbers, pos, table, cn=None, grad=False):
    """Atom-resolved D3(BJ) energy (s9 = 0) and, optionally, (dE/dR direct part, dE/dCN)."""
    par = params()
    z = np.asarray(numbers)
    m = make_mol(z)
    if cn is None:
        cn, _ = cn_d3(m, pos)
    s6, s8, a1, a2 = par.d3["s6"], par.d3["s8"], par.d3["a1"], par.d3["a2"]
    gw, dgw = d3_weights(z, cn, table, grad=grad)
    rc6 = table["c6"][z[:, None], z[None, :]]  # (nat,nat,7,7)
    c6 = np.einsum("ia,jb,ijab->ij", gw, gw, rc6)
    n = len(z)
    offd = ~np.eye(n, dtype=bool)
    dist = np.where(offd, _cdist(pos), EPS)
    qq = 3.0 * table["r4r2"][z][:, None] * table["r4r2"][z][None, :]
    r0 = a1 * np.sqrt(qq) + a2
    ok = offd & (dist <= D3_DISP_CUTOFF)
    t6 = np.where(ok, 1.0 / (dist**6 + r0**6), 0.0)
    t8 = np.where(ok, 1.0 / (dist**8 + r0**8), 0.0)
    f = s6 * t6 + s8 * qq * t8
    e = -0.5 * (c6 * f).sum(-1)
    if not grad:
        return e, None, None
    df = np.where(ok, -(s6 * 6.0 * dist**5 * t6**2 + s8 * qq * 8.0 * dist**7 * t8**2), 0.0)  # d f / d R
    rij = pos[:, None, :] - pos[None, :, :]
    g_direct = ((-c6 * df / dist)[:, :, None] * rij).sum(1)
    dc6 = np.einsum("ia,jb,ijab->ij", dgw, gw, rc6)  # d c6_ij / d cn_i
    dedcn = -(dc6 * f).sum(-1)
    return e, g_direct, dedcn


def synthetic_d3_table(seed: int = 7, zmax: int = 86):
    """A made-up table with the SHAPE of tad-dftd3's reference data: for testing the arithmetic only."""
    rng = np.random.default_rng(seed)
    cn = -np.ones((zmax + 1, 7))
    nref = rng.integers(2, 6, size=zmax + 1)
    for z in range(1, zmax + 1):
        cn[z, : nref[z]] = np.sort(rng.uniform(0.0, 4.5, size=nref[z]))
        cn[z, 0] = 0.0
    a = rng.uniform(2.0, 60.0, size=(zmax + 1, 7))
    c6 = np.sqrt(a[:, None, :, None] * a[None, :, None, :]) * (1.0 + 0.1 * rng.uniform(size=(zmax + 1, zmax + 1, 7, 7)))
    c6 = 0.5 * (c6 + c6.transpose(1, 0, 3, 2))
    r4r2 = rng.uniform(1.5, 9.0, size=zmax + 1)
    return {"cn": cn, "c6": c6, "r4r2": r4r2}


# --------------------------------------------------------------------------------------
# second/third-order electrostatics
# --------------------------------------------------------------------------------------
def gamma_shell(m: Mol, pos: np.ndarray):
    """components/interactions/coulomb/secondorder.py:799-870 with average.py:41-56 (harmonic)."""
    par = params()
    gexp = par.gexp
    h = _shell_param(m, "lgam") * np.array([par.elem[int(z)]["gam"] for z in m.numbers[m.sh_atom]])
    dist = _cdist(pos)
    offatom = ~np.eye(m.nat, dtype=bool)
    dg = np.where(offatom, (dist + EPS) ** gexp, EPS)[m.sh_atom[:, None], m.sh_atom[None, :]]
    h1 = 1.0 / (h + EPS)
    avg = 2.0 / (h1[:, None] + h1[None, :])
    return 1.0 / (dg + avg ** (-gexp)) ** (1.0 / gexp)


def gam3(m: Mol):
    par = params()
    return np.array([par.elem[int(z)]["gam3"] for z in m.numbers])


# --------------------------------------------------------------------------------------
# EEQ guess (tad-multicharge 0.5.0 get_eeq_charges; called at scf/guess.py:118-120)
# --------------------------------------------------------------------------------------
def eeq_charges(m: Mol, pos: np.ndarray, chrg: float):
    par = params()
    z = m.numbers
    n = m.nat
    dist = _cdist(pos)
    offd = ~np.eye(n, dtype=bool)
    rc = par.cov_d3[z]
    r0 = rc[:, None] + rc[None, :]
    mask = offd & (dist <= EEQ_CN_CUTOFF)
    dsafe = np.where(offd, dist, 1.0)
    cf = np.where(mask, 0.5 * (1.0 + _erf(-EEQ_KCN * (dsafe / r0 - 1.0))), 0.0)
    cn = cf.sum(-1)
    cn = math.log(1.0 + math.exp(EEQ_CN_MAX)) - np.log(1.0 + np.exp(EEQ_CN_MAX - cn))
    rhs = np.zeros(n + 1)
    rhs[:n] = -par.eeq_chi[z] + np.sqrt(np.maximum(cn, EPS)) * par.eeq_kcn[z]
    rhs[n] = chrg
    rad = par.eeq_rad[z]
    gam = 1.0 / np.sqrt(rad[:, None] ** 2 + rad[None, :] ** 2)
    A = np.zeros((n + 1, n + 1))
    A[:n, :n] = np.where(offd, _erf(dsafe * gam) / dsafe, 0.0)
    A[np.arange(n), np.arange(n)] = par.eeq_eta[z] + math.sqrt(2.0 / math.pi) / rad
    A[:n, n] = 1.0
    A[n, :n] = 1.0
    x = np.linalg.solve(A, rhs)
    return x[:n]


def guess_orbital_charges(m: Mol, qat: np.ndarray):
    """scf/guess.py:122-182: atom charge split equally over shells, then equally over the shell's AOs."""
    nsh_at = np.bincount(m.sh_atom, minlength=m.nat)
    qsh = qat[m.sh_atom] / nsh_at[m.sh_atom]
    return qsh[m.ao_sh] / (2 * m.sh_l[m.ao_sh] + 1)


# --------------------------------------------------------------------------------------
# Fermi filling (wavefunction/filling.py:201-366)
# --------------------------------------------------------------------------------------
def fermi_occupation(nel: np.ndarray, emo: np.ndarray, kt: float, maxiter: int = 200, thr: float | None = None):
    """nel (2,), emo (nao,) ascending -> occupation (2,nao)."""
    n = emo.shape[0]
    if abs(nel.sum()) < EPS:
        return np.zeros((2, n))
    thr = math.sqrt(EPS) if thr is None else thr
    idx = np.arange(1, n + 1)[None, :] - nel[:, None]
    homo = np.argmax(idx >= -1e-15 * 5, axis=-1)
    lumo_missing = (n - 1) <= homo
    lumo = np.where(lumo_missing, homo, homo + 1)
    ef = np.where(nel != 0, 0.5 * (emo[homo] + emo[lumo]), 0.0)[:, None]
    not_empty = (nel != 0)[:, None]
    e2 = np.where(not_empty, emo[None, :], 0.0)
    for _ in range(maxiter):
        ex = (e2 - ef) / kt
        small = ex < 50
        et = np.exp(np.where(small, ex, 0.0))
        f = np.where(small, 1.0 / (et + 1.0), 0.0)
        df = np.where(small, et / (kt * (et + 1.0) ** 2), EPS)
        nn = f.sum(-1, keepdims=True)
        resid = homo[:, None] - nn + 1
        ef = ef + resid / df.sum(-1, keepdims=True)
        if np.all(np.abs(resid) <= thr):
            return np.where(not_empty, f, 0.0)
    raise RuntimeError("Fermi energy failed to converge.")


# --------------------------------------------------------------------------------------
# Anderson mixer (scf/mixer/anderson.py:163-317; convergence scf/mixer/base.py:229-256)
# --------------------------------------------------------------------------------------
class Anderson:
    def __init__(self, n, damp=0.5, damp_init=0.1, generations=5, diagonal_offset=0.01, soft_start=True):
        self.damp, self.damp_init, self.gen, self.off, self.soft = damp, damp_init, generations, diagonal_offset, soft_start
        self.step = 0
        self.x_hist = np.zeros((generations + 1, n))
        self.f = np.zeros((generations + 1, n))
        self.delta = None

    def iter(self, x_new, x_old):
        if self.step == 0:
            self.x_hist[0] = x_old
        self.step += 1
        self.f[0] = x_new - x_old
        if self.step > self.gen or (self.step > 1 and not self.soft):
            n = min(self.step - 1, self.gen)
            df = self.f[0][None, :] - self.f[1 : n + 1]
            a = df @ df.T
            b = df @ self.f[0]
            a[np.diag_indices(n)] *= 1.0 + self.off**2
            th = np.linalg.solve(a, b)
            x_bar = th @ (self.x_hist[1 : n + 1] - self.x_hist[0][None, :]) + self.x_hist[0]
            f_bar = th @ (-df) + self.f[0]
            x_mix = x_bar + self.damp * f_bar
        else:
            x_mix = self.x_hist[0] + self.f[0] * self.damp_init
        self.f = np.roll(self.f, 1, 0)
        self.x_hist = np.roll(self.x_hist, 1, 0)
        self.x_hist[0] = x_mix
        self.delta = self.f[1].copy()
        return x_mix

    def converged(self, x_tol, x_tol_max):
        return (np.linalg.norm(self.delta) < x_tol) and (np.abs(self.delta).max() < x_tol_max)


class Simple:
    """scf/mixer/simple.py:88-151."""

    def __init__(self, n, damp=0.5, **_):
        self.damp = damp
        self.delta = None
        self.step = 0

    def iter(self, x_new, x_old):
        self.step += 1
        self.delta = x_new - x_old
        return x_old + self.damp * self.delta

    converged = Anderson.converged


# --------------------------------------------------------------------------------------
# single point (calculators/types/energy.py:78-441; scf/iterator.py:51-189; scf/base.py; unrolling/default.py:71-136)
# --------------------------------------------------------------------------------------
DEFAULT_OPTS = dict(
    maxiter=100, mixer="anderson", damp=0.5, damp_init=0.1, damp_generations=5, damp_diagonal_offset=0.01,
    damp_soft_start=True, x_atol=1e-4, x_atol_max=1e-5, fermi_etemp=300.0, fermi_maxiter=200, fermi_thresh=None,
    guess="eeq", exclude=(), int_cutoff=INT_CUTOFF,
    # gradient: add the first-order response of the not fully converged SCF state (see _scf_response)
    grad_response=True, response_maxiter=12, response_tol=1e-8,
)


@dataclass
class Result:
    energy: float = 0.0
    e_atom: np.ndarray = None
    e_scf: float = 0.0
    e_rep: float = 0.0
    e_xb: float = 0.0
    e_disp: float = 0.0
    fenergy: float = 0.0
    q_orb: np.ndarray = None
    q_sh: np.ndarray = None
    q_at: np.ndarray = None
    iterations: int = 0
    converged: bool = True
    S: np.ndarray = None
    H0: np.ndarray = None
    P: np.ndarray = None
    W: np.ndarray = None
    v_orb: np.ndarray = None
    v_in: np.ndarray = None  # input potential of the final solve (the un-mixed potential of the last iteration)
    C: np.ndarray = None
    emo: np.ndarray = None
    occ: np.ndarray = None
    cn: np.ndarray = None
    gradient: np.ndarray = None  # dE/dR (analytic, converged-SCF)
    gradient_parts: dict = None  # "h0_dedr": overlap-derivative + scaling-function part of the electronic gradient


def _potential(m: Mol, q_orb, gam, g3):
    """scf/base.py:702-727 -> interactions/base.py:134-181, secondorder.py:414-442, thirdorder.py:303-331."""
    q_sh = np.bincount(m.ao_sh, weights=q_orb, minlength=m.nsh)
    q_at = np.bincount(m.sh_atom, weights=q_sh, minlength=m.nat)
    v_sh = gam @ q_sh + (g3 * q_at**2)[m.sh_atom]
    return v_sh[m.ao_sh], q_sh, q_at


def singlepoint(numbers, positions, chrg: float = 0.0, opts: dict | None = None, grad: bool = False,
                d3_energy=None, d3_table=None) -> Result:
    """One GFN1-xTB single point with dxtb's default path.  ``d3_energy``: optional callable
    (numbers, positions) -> atomwise dispersion energies (the D3 table is third-party data)."""
    par = params()
    o = dict(DEFAULT_OPTS)
    o.update(opts or {})
    excl = set(o["exclude"])
    m = make_mol(numbers)
    pos = np.asarray(positions, dtype=np.float64)[: m.nat]
    res = Result()

    # classicals (energy.py:154-166)
    e_at = np.zeros(m.nat)
    g_tot = np.zeros((m.nat, 3))
    if "rep" not in excl:
        er, gr = repulsion(m, pos, grad=grad)
        e_at += er
        res.e_rep = er.sum()
        if grad:
            g_tot += gr
    if "hal" not in excl:
        ex = halogen(m, pos, grad=grad)
        if grad:
            ex, gx = ex
            g_tot += gx
        e_at += ex
        res.e_xb = ex.sum()
    if "disp" not in excl and d3_energy is not None:
        ed = d3_energy(m.numbers, pos)
        e_at += ed
        res.e_disp = ed.sum()
    d3_dedcn = None
    if "disp" not in excl and d3_table is not None:
        ed, gd, d3_dedcn = d3_dispersion(m.numbers, pos, d3_table, grad=grad)
        e_at += ed
        res.e_disp = ed.sum()
        if grad:
            g_tot += gd

    # integrals (energy.py:170-268)
    S, dS = overlap(m, pos, cutoff=o["int_cutoff"], grad=grad)
    cn, dcfdr = cn_d3(m, pos, grad=grad)
    H0 = h0(m, pos, S, cn)
    res.S, res.H0, res.cn = S, H0, cn

    gam = gamma_shell(m, pos) if "es2" not in excl else np.zeros((m.nsh, m.nsh))
    g3 = gam3(m) if "es3" not in excl else np.zeros(m.nat)

    # reference occupation (scf/iterator.py:147-189)
    n0 = _shell_param(m, "refocc")[m.ao_sh] / (2 * m.sh_l[m.ao_sh] + 1)
    nel = n0.sum() - chrg
    nuhf = float(np.remainder(np.round(nel), 2))
    diff = min(nuhf, nel)
    nb = (nel - diff) / 2.0
    nab = np.round(np.array([nb + diff, nb]))  # scf/base.py:878 rounds the aufbau sum
    kt = o["fermi_etemp"] * par.kelvin2au

    # guess (scf/guess.py:35-120)
    if o["guess"] == "eeq":
        q0 = guess_orbital_charges(m, eeq_charges(m, pos, chrg))
    else:
        q0 = np.zeros(m.nao)

    state = {}

    def fcn(v):  # iterate_potential, scf/base.py:651-675
        F = H0 - 0.5 * S * (v[:, None] + v[None, :])
        emo, C = _scipy_eigh(F, S)
        if kt >= 3e-7:
            occ = fermi_occupation(nab, emo, kt, o["fermi_maxiter"], o["fermi_thresh"])
        else:
            occ = np.stack([(np.arange(m.nao) < nab[0]) * 1.0, (np.arange(m.nao) < nab[1]) * 1.0])
        focc = occ.sum(0)
        P = (C * focc[None, :]) @ C.T
        e_orb = np.einsum("ik,ki->i", P, H0)
        q = n0 - np.einsum("ik,ki->i", P, S)
        vnew, q_sh, q_at = _potential(m, q, gam, g3)
        state.update(F=F, emo=emo, C=C, occ=occ, P=P, e_orb=e_orb, q=q, q_sh=q_sh, q_at=q_at)
        return vnew

    guess_v, _, _ = _potential(m, q0, gam, g3)
    mixer_cls = Anderson if o["mixer"] in ("anderson", "broyden") else Simple
    mixer = mixer_cls(m.nao, damp=o["damp"], damp_init=o["damp_init"], generations=o["damp_generations"],
                      diagonal_offset=o["damp_diagonal_offset"], soft_start=o["damp_soft_start"])
    iters = 1
    v_new = fcn(guess_v)
    converged = True
    if o["maxiter"] > 0:
        v = mixer.iter(v_new, guess_v)
        converged = False
        for _ in range(o["maxiter"]):
            v_new = fcn(v)
            iters += 1
            v = mixer.iter(v_new, v)
            if mixer.converged(o["x_atol"], o["x_atol_max"]):
                converged = True
                break
    # converged_to_charges: one more solve with the un-mixed potential (scf/base.py:497-501)
    fcn(v_new)
    res.v_in = v_new.copy()
    st = state
    res.iterations, res.converged = iters, converged
    res.q_orb, res.q_sh, res.q_at = st["q"], st["q_sh"], st["q_at"]
    # potential of the FINAL charges: what the reference hands to its analytic gradient (scf/base.py:468)
    v_fin = _potential(m, st["q"], gam, g3)[0]
    res.P, res.emo, res.occ, res.v_orb, res.C = st["P"], st["emo"], st["occ"], v_fin, st["C"]

    # energies (scf/base.py:514-534, 558-607; interactions/base.py:305-360)
    v_es2 = gam @ st["q_sh"]
    e_at += np.bincount(m.ao_atom, weights=st["e_orb"], minlength=m.nat)
    e_at += np.bincount(m.sh_atom, weights=0.5 * st["q_sh"] * v_es2, minlength=m.nat)
    e_at += g3 * st["q_at"] ** 3 / 3.0
    occ = st["occ"]
    o1 = np.maximum(occ, EPS)
    o2 = np.maximum(1.0 - occ, EPS)
    G = float(np.log(o1**o1 * o2**o2).sum() * kt)
    e_at += G / m.nat
    res.fenergy = G
    res.e_atom = e_at
    res.energy = float(e_at.sum())
    res.e_scf = res.energy - res.e_rep - res.e_xb - res.e_disp

    if grad:
        focc = occ.sum(0)
        W = (st["C"] * (focc * st["emo"])[None, :]) @ st["C"].T
        res.W = W
        parts: dict = {}
        P_eff, W_eff, v_grad, y_sh = st["P"], W, v_fin, None
        if o["grad_response"] and o["maxiter"] > 0:
            Z, ZW, y, Ky = _scf_response(m, S, st["C"], st["emo"], occ, kt, gam, g3, st["q_at"], v_fin - res.v_in,
                                         o["response_maxiter"], o["response_tol"])
            P_eff, W_eff, v_grad = st["P"] + Z, W + ZW, v_fin + Ky
            y_sh = np.bincount(m.ao_sh, weights=y, minlength=m.nsh)
            parts["response_y"] = y
        g_tot += _electronic_gradient(m, pos, S, dS, P_eff, W_eff, v_grad, cn, dcfdr, st["q_sh"], gam, parts, y_sh=y_sh)
        res.gradient_parts = parts
        if d3_dedcn is not None:  # CN chain rule of the dispersion energy (same exp-count CN as H0)
            g_tot += (dcfdr * (d3_dedcn[:, None] + d3_dedcn[None, :])[:, :, None]).sum(1)
        res.gradient = g_tot
    return res


def _scf_response(m: Mol, S, C, emo, occ, kt, gam, g3, q_at, dv, maxiter=12, tol=1e-8):
    """First-order correction of the analytic gradient for a NOT fully converged SCF state.

    The reference's forces are autograd through the unrolled SCF (calculators/types/autograd.py:80-201): the exact
    derivative of the energy it computes, E = E[v_in(R), R], where v_in is the un-mixed potential that enters the final
    solve (scf/base.py:497-501) and v_out = V(q_out) the potential of the resulting charges.  With Omega = sum f eps + G
    stationary in the orbitals and occupations,
        dE/dR = [Hellmann-Feynman + Pulay terms] + (v_out - v_in) . dq_out/dR ,
    and the last term (first order in the SCF residual dv = v_out - v_in; 1e-6..1e-5 Eh/bohr at dxtb's default thresholds)
    is what the converged-SCF formula of analytical.py:63-222 leaves out.  dq_out/dR is expanded with the response of the
    converged fixed point (coupled-perturbed equations in adjoint / Z-vector form):
        y = (1 - chi K)^-1 chi dv,   u = dv + K y,
        chi w = -diag(Z_w S),  Z_w = C [ (C^T A_w C) o G ] C^T,  A_w = -1/2 S o (w (+) w),  G_ij = (f_j - f_i)/(e_j - e_i),
        K = dV/dq = gamma + 2 Gamma q_A  (secondorder.py:414-442, thirdorder.py:303-331),
    so that the correction is Tr[Z_u dF/dR] - Tr[ZW_u dS/dR] - sum (K y)_mu P_mu,nu dS_mu,nu/dR + y . dV/dR|_q, i.e. the
    standard gradient evaluated with P + Z_u, W + ZW_u, v_out + K y and the extra ES2 cross term (y_sh).  Diagonal of G:
    Fermi-function derivative per spin channel with the Fermi-level shift projected out (wavefunction/filling.py:201-366).
    Returns Z_u, ZW_u, y, K y.  Residual error: (v_out - v_in) . chi . (dv_K/dR - dv*/dR), second order in the residual
    unless the trajectory's own derivative is unconverged (symmetry-breaking soft modes, e.g. the NO2 radical)."""
    n = len(emo)
    f = occ.sum(0)
    fp_s = -(occ * (1.0 - occ)) / kt if kt >= 3e-7 else np.zeros_like(occ)
    fp = fp_s.sum(0)
    de = emo[None, :] - emo[:, None]
    close = np.abs(de) <= 1e-9
    den = np.where(close, 1.0, de)
    G = np.where(close, 0.5 * (fp[None, :] + fp[:, None]), (f[None, :] - f[:, None]) / den)
    fe = f * emo
    Ge = np.where(close, 0.5 * ((f + emo * fp)[None, :] + (f + emo * fp)[:, None]), (fe[None, :] - fe[:, None]) / den)

    def kernel(y):  # K y on orbital-resolved vectors
        y_sh = np.bincount(m.ao_sh, weights=y, minlength=m.nsh)
        y_at = np.bincount(m.sh_atom, weights=y_sh, minlength=m.nat)
        return (gam @ y_sh + (2.0 * g3 * q_at * y_at)[m.sh_atom])[m.ao_sh]

    def respond(w, want_w=False):
        At = C.T @ (-0.5 * S * (w[:, None] + w[None, :])) @ C
        Zt = At * G
        d = np.diag(At).copy()
        zd = np.zeros(n)
        for s in range(2):
            norm = fp_s[s].sum()
            abar = (fp_s[s] * d).sum() / norm if abs(norm) > TINY else 0.0
            zd += (d - abar) * fp_s[s]
        np.fill_diagonal(Zt, zd)
        Z = C @ Zt @ C.T
        if not want_w:
            return -np.einsum("ij,ij->i", Z, S)
        ZWt = At * Ge
        np.fill_diagonal(ZWt, d * f + zd * emo)
        return Z, C @ ZWt @ C.T

    if np.abs(dv).max() < tol:
        return np.zeros((n, n)), np.zeros((n, n)), np.zeros(n), np.zeros(n)
    # iterate in potential space, w = K y:  w = K (z0 + chi w); Anderson from an empty history
    z0 = respond(dv)
    w = kernel(z0)
    y = z0
    mixer = Anderson(n, damp=0.5, damp_init=0.5, generations=5, diagonal_offset=0.01, soft_start=False)
    for _ in range(maxiter):
        y = z0 + respond(w)
        w_new = kernel(y)
        if np.abs(w_new - w).max() < tol:
            w = w_new
            break
        w = mixer.iter(w_new, w)
    return (*respond(dv + w, want_w=True), y, w)


def _electronic_gradient(m: Mol, pos, S, dS, P, W, v, cn, dcfdr, q_sh, gam, parts: dict | None = None, y_sh=None):
    """calculators/types/analytical.py:63-222 + xtb/gfn1.py:185-408 + secondorder.py:873-926 + ncoord/utils.py:30-52."""
    f = _h0_shell_factors(m, pos, cn)
    a2s = m.ao_sh
    offatom_ao = m.ao_atom[:, None] != m.ao_atom[None, :]
    hcore = f["hsh"][a2s[:, None], a2s[None, :]]
    ph = P * hcore
    # overlap-derivative term
    sval = np.where(offatom_ao, 2.0 * (ph - W) - P * (v[:, None] + v[None, :]), 0.0)
    g_orb = (dS * sval[:, :, None]).sum(1)  # (nao,3): derivative w.r.t. centre of mu
    g = np.zeros((m.nat, 3))
    np.add.at(g, m.ao_atom, g_orb)
    # scaling-function term
    dist_sh = f["dist"][m.sh_atom[:, None], m.sh_atom[None, :]]
    dvar_pi = np.where(f["offatom"], (f["tmp_a"] * f["shpoly"][None, :] + f["tmp_b"] * f["shpoly"][:, None]) * f["rr"] * 0.5
                       / np.where(f["offatom"], dist_sh, 1.0) ** 2, 0.0)
    phs = ph * S
    phs_sh = np.zeros((m.nsh, m.nsh))
    np.add.at(phs_sh, (a2s[:, None], a2s[None, :]), phs)
    dpi_sh = 2.0 * phs_sh * dvar_pi / f["var_pi"]
    dpi_at = np.zeros((m.nat, m.nat))
    np.add.at(dpi_at, (m.sh_atom[:, None], m.sh_atom[None, :]), dpi_sh)
    rij = pos[:, None, :] - pos[None, :, :]
    g += (dpi_at[:, :, None] * rij).sum(1)
    if parts is not None:
        parts["h0_dedr"] = g.copy()  # = dedr of GFN1Hamiltonian.get_gradient (xtb/gfn1.py:185-408), without the CN chain
    # CN term
    ps_sh = np.zeros((m.nsh, m.nsh))
    np.add.at(ps_sh, (a2s[:, None], a2s[None, :]), P * S)
    dhdcn = np.where(f["offatom"], -f["kcn"][None, :] * f["var_pi"] * f["var_k"], -f["kcn"][None, :])
    dedcn_sh = (ps_sh * dhdcn).sum(0)
    dedcn = np.bincount(m.sh_atom, weights=dedcn_sh, minlength=m.nat)
    # dE/dR_A = sum_B (dedcn_A + dedcn_B) dcf_AB/dR_A
    g_cn = (dcfdr * (dedcn[:, None] + dedcn[None, :])[:, :, None]).sum(1)
    g += g_cn
    if parts is not None:
        parts["h0_dedcn"] = dedcn.copy()  # = dedcn of GFN1Hamiltonian.get_gradient
        parts["h0_dcn"] = g_cn.copy()  # = ncoord get_dcn(dcndr, dedcn)
    # ES2: E = 1/2 q g q, d gamma/dR_A = -gamma^3 (R_A-R_B) for off-atom shell pairs (gexp = 2)
    offsh = f["offatom"]
    qq = q_sh[:, None] * q_sh[None, :]
    if y_sh is not None:  # y . dV/dR at fixed charges (response of the SCF potential, _scf_response)
        qq = qq + y_sh[:, None] * q_sh[None, :] + q_sh[:, None] * y_sh[None, :]
    dg = np.where(offsh, -(gam**3), 0.0) * qq
    dg_at = np.zeros((m.nat, m.nat))
    np.add.at(dg_at, (m.sh_atom[:, None], m.sh_atom[None, :]), dg)
    g += (dg_at[:, :, None] * rij).sum(1)
    return g
