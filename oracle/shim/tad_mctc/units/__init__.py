from . import codata, energy, length  # noqa: F401
from .codata import CODATA, get_constant  # noqa: F401
from .energy import *  # noqa: F401,F403
from .energy import (AU2COULOMB, AU2EV, AU2JOULE, AU2KCALMOL, AU2KELVIN, AU2KJMOL, AU2RCM, COULOMB2AU, EV2AU, JOULE2AU,  # noqa: F401
                     KCALMOL2AU, KELVIN2AU, KJMOL2AU, RCM2AU)
from .length import AA2AU, AU2AA, AU2METER, AU2NM, METER2AU, NM2AU  # noqa: F401

# atomic unit of time and derived conversion factors used by dxtb's spectroscopy modules (outside the hot path)
AU2SECOND = CODATA.h / (2.0 * 3.141592653589793 * CODATA.hartree)
SECOND2AU = 1.0 / AU2SECOND
AU2VAA = AU2JOULE / (CODATA.e * AU2AA)  # electric field: Hartree/(e Bohr) -> V/Angstrom
VAA2AU = 1.0 / AU2VAA
AMU2AU = 1.66053906660e-27 / CODATA.me
AU2AMU = 1.0 / AMU2AU
DEBYE2AU = 1e-21 / CODATA.c / (CODATA.e * AU2METER)
AU2DEBYE = 1.0 / DEBYE2AU
AU2KMMOL = (DEBYE2AU / AA2AU) ** -2 * AU2AMU * 42.256  # IR intensities, km/mol
AU2AA4AMU = AU2AA**4 / AMU2AU  # Raman activities
