"""ctypes binding of the C ABI declared in ``include/xtb_b200.h``.

This is the stub a dxtb maintainer would add to call the B200 kernels.  There is
no CPU fallback: importing this module without the compiled ``_C.so`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_SO = Path(__file__).resolve().parent / "_C.so"
if os.environ.get("DXTB_B200_LIB"):  # developer A/B builds
    _SO = Path(os.environ["DXTB_B200_LIB"])

ATPAR, SHPAR, CGTO, MAXPRIM = 12, 6, 16, 7
(AT_RAD, AT_RCOV, AT_EN, AT_AREP, AT_ZEFF, AT_GAM3, AT_XBOND, AT_EEQ_CHI, AT_EEQ_ETA, AT_EEQ_KCN, AT_EEQ_RAD, AT_R4R2) = range(12)
(SH_LEVEL, SH_KCN, SH_SHPOLY, SH_ETA, SH_REFOCC) = range(5)

STATUS_SCF_NOT_CONVERGED = 1
STATUS_FERMI_FAILED = 2
STATUS_JACOBI_NOT_CONVERGED = 4
STATUS_S_NOT_POSDEF = 8

_vp = C.c_void_p


class XtbBatch(C.Structure):
    _fields_ = [
        ("nb", C.c_int32), ("nat_tot", C.c_int32), ("nsh_tot", C.c_int32), ("nao_tot", C.c_int32),
        ("nat_max", C.c_int32), ("nsh_max", C.c_int32), ("nao_max", C.c_int32),
        ("nspecies", C.c_int32), ("ncgto", C.c_int32), ("has_xb", C.c_int32), ("nsh_l_max", C.c_int32 * 4),
        ("mat_total", C.c_int64), ("gam_total", C.c_int64), ("eeq_total", C.c_int64),
        ("at_off", _vp), ("sh_off", _vp), ("ao_off", _vp), ("mat_off", _vp), ("gam_off", _vp), ("eeq_off", _vp),
        ("at_z", _vp), ("at_species", _vp), ("at_sh0", _vp), ("at_nsh", _vp), ("at_par", _vp),
        ("sh_atom", _vp), ("sh_l", _vp), ("sh_ao", _vp), ("sh_cgto", _vp), ("sh_type", _vp), ("sh_by_l", _vp),
        ("nsh_l", _vp), ("sh_par", _vp), ("ao_sh", _vp), ("cgto", _vp), ("kpair", _vp),
        ("hscale", C.c_double * 36),
        ("enscale", C.c_double), ("rep_kexp", C.c_double), ("xb_damp", C.c_double), ("xb_rscale", C.c_double),
        ("gexp", C.c_double),
        ("int_cutoff", C.c_double), ("rep_cutoff", C.c_double), ("xb_cutoff", C.c_double), ("cn_cutoff", C.c_double),
        ("kcn_d3", C.c_double),
        ("d3_refcn", _vp), ("d3_c6", _vp),
        ("d3_s6", C.c_double), ("d3_s8", C.c_double), ("d3_a1", C.c_double), ("d3_a2", C.c_double),
        ("d3_cutoff", C.c_double), ("d3_wf", C.c_double),
    ]


class XtbScfOpts(C.Structure):
    _fields_ = [
        ("maxiter", C.c_int32), ("mixer", C.c_int32), ("generations", C.c_int32), ("soft_start", C.c_int32),
        ("fermi_maxiter", C.c_int32), ("want_density", C.c_int32), ("use_smem", C.c_int32), ("jacobi_max_sweeps", C.c_int32),
        ("subspace", C.c_int32), ("subspace_maxiter", C.c_int32),
        ("damp", C.c_double), ("damp_init", C.c_double), ("diag_offset", C.c_double),
        ("x_atol", C.c_double), ("x_atol_max", C.c_double), ("kt", C.c_double), ("fermi_thresh", C.c_double),
        ("jacobi_tol", C.c_double), ("jacobi_tol_iter", C.c_double), ("subspace_tol", C.c_double), ("subspace_gap", C.c_double),
        ("mol_list", _vp), ("list_len", C.c_int32), ("list_nao_max", C.c_int32), ("list_nsh_max", C.c_int32),
        ("list_nat_max", C.c_int32), ("persistent", C.c_int32), ("reserved0", C.c_int32),
    ]


EXPORTS = {
    "xtb_version": (C.c_int, []),
    "xtb_clock_probe": (C.c_int, [_vp, _vp]),
    "xtb_sizeof_batch": (C.c_int, []),
    "xtb_sizeof_scf_opts": (C.c_int, []),
    "xtb_geometry_fwd": (C.c_int, [_vp] * 6),
    "xtb_eeq_guess": (C.c_int, [_vp] * 6),
    "xtb_eeq_guess_large": (C.c_int, [_vp, C.c_int32, C.c_int32, C.c_int64, C.c_int64] + [_vp] * 5),
    "xtb_gamma_fwd": (C.c_int, [_vp] * 4),
    "xtb_overlap_h0_fwd": (C.c_int, [_vp] * 6),
    "xtb_scf_workspace_bytes": (C.c_int64, [_vp, _vp]),
    "xtb_scf_smem_bytes": (C.c_int64, [_vp]),
    "xtb_scf_smem_bytes_for": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "xtb_scf_smem_bytes_mode": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "xtb_scf_run": (C.c_int, [_vp] * 22),
    "xtb_scf_large_workspace_bytes": (C.c_int64, [C.c_int32] * 4),
    "xtb_scf_run_large": (C.c_int, [_vp, _vp] + [C.c_int32] * 4 + [_vp] * 19 + [C.c_int64, _vp]),
    "xtb_grad_bwd": (C.c_int, [_vp] * 16),
    "xtb_d3_fwd": (C.c_int, [_vp] * 6),
}

_lib = None


def lib() -> C.CDLL:
    """Load ``_C.so`` (once) and declare every prototype. Raises if the extension was not built."""
    global _lib
    if _lib is None:
        if not _SO.exists():
            raise ImportError(
                f"dxtb_b200 CUDA extension not built ({_SO} missing). Run `python -m dxtb_b200.build` "
                "(needs nvcc, sm_100a). There is no CPU fallback."
            )
        handle = C.CDLL(str(_SO))
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.xtb_sizeof_batch() != C.sizeof(XtbBatch) or handle.xtb_sizeof_scf_opts() != C.sizeof(XtbScfOpts):
            raise ImportError("dxtb_b200: struct layout mismatch between _abi.py and include/xtb_b200.h")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed with status {rc}" + (" (CUDA launch error)" if rc > 0 else " (bad argument)"))
