"""CODATA-2018 constants (same numbers as dxtb_b200/data/gfn1_param.json "codata2018")."""
from ..data._blob import third_party


class CODATA:
    h = 6.62607015e-34
    c = 299792458.0
    kb = 1.380649e-23
    na = 6.02214076e23
    e = 1.602176634e-19
    alpha = 7.2973525693e-3
    me = 9.1093837015e-31
    bohr = third_party()["codata2018"]["bohr_m"]
    hartree = third_party()["codata2018"]["hartree_j"]


def get_constant(name):
    return getattr(CODATA, name)
