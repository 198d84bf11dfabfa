"""Counting functions of the coordination numbers and their derivatives w.r.t. the distance."""
import math

import torch

KCN_D3 = 16.0
KCN_EEQ = 7.5
KA, KB, R_SHIFT = 10.0, 20.0, 2.0


def exp_count(r, r0, kcn: float = KCN_D3):
    return 1.0 / (1.0 + torch.exp(-kcn * (r0 / r - 1.0)))


def dexp_count(r, r0, kcn: float = KCN_D3):
    e = torch.exp(-kcn * (r0 / r - 1.0))
    return (-kcn * r0 * e) / (r**2 * ((e + 1.0) ** 2))


def erf_count(r, r0, kcn: float = KCN_EEQ):
    return 0.5 * (1.0 + torch.erf(-kcn * (r / r0 - 1.0)))


def derf_count(r, r0, kcn: float = KCN_EEQ):
    return -kcn / math.sqrt(math.pi) / r0 * torch.exp(-(kcn**2) * (r - r0) ** 2 / r0**2)


def gfn2_count(r, r0, ka: float = KA, kb: float = KB, r_shift: float = R_SHIFT):
    return exp_count(r, r0, ka) * exp_count(r, r0 + r_shift, kb)


def dgfn2_count(r, r0, ka: float = KA, kb: float = KB, r_shift: float = R_SHIFT):
    return dexp_count(r, r0, ka) * exp_count(r, r0 + r_shift, kb) + exp_count(r, r0, ka) * dexp_count(r, r0 + r_shift, kb)
